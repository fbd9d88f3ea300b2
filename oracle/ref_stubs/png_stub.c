/* TEST INFRASTRUCTURE ONLY.  Stand-in for the one function of /root/reference/src/png_util.c that draw.c calls
 * (write_gray_png, png_util.c:238-262): the same float -> [0,255] scaling, written as a binary PGM (P5) instead of
 * a PNG, because libpng is not available here.  The reference stores the intensity in all three RGB channels of an
 * 8-bit PNG; one channel of it is what this file holds. */
#include <stdio.h>
#include <stdlib.h>

int write_gray_png(FILE *outfile, int width, int height, float *img, float minI, float maxI) {
    unsigned char *row = (unsigned char *)malloc((size_t)width);
    fprintf(outfile, "P5\n%d %d\n255\n", width, height);
    for (int n = 0; n < height; ++n) {
        for (int m = 0; m < width; ++m) {
            int id = m + n * width;
            unsigned int I = (unsigned int)(255 * (img[id] - minI) / (maxI - minI)); /* png_util.c:249 */
            row[m] = (unsigned char)I;
        }
        fwrite(row, 1, (size_t)width, outfile);
    }
    free(row);
    fclose(outfile);
    return 1;
}
