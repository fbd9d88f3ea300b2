/* TEST INFRASTRUCTURE ONLY.  Empty stand-in for <png.h>: libpng is not installed in the build image, and the only
 * thing the reference's draw.c needs from png_util.h are the prototypes that header declares itself.  With this
 * directory on the include path the UNMODIFIED /root/reference/src/draw.c compiles; png_stub.c provides
 * write_gray_png() (see oracle/Makefile, target _ref/draw_ref). */
#ifndef FIZZ_ORACLE_PNG_STUB_H
#define FIZZ_ORACLE_PNG_STUB_H
#endif
