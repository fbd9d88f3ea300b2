/* fizz_b200.h -- C-ABI of the B200-native batched stepping engine for Fizz-2D's hot path.
 *
 * The reference (BKaperick/Fizz-2D) has NO plugin / FFI layer: the hot path is the Python method
 * `World.update()` (src/physics.py:76-259) called from `main.py:98`.  This header is therefore the boundary a
 * maintainer would bind with ctypes from a batched driver (see INTEGRATION.md for the stub); every entry point
 * cites the reference interface it stands in for.  Plain pointers and sizes only; no torch / numpy types.
 *
 * Model: a *group* is W independent worlds that share one scene topology (F free polygons with fixed vertex
 * counts, X fixed polygons, one parameter set).  World state lives in ONE caller-owned device slab of 8-byte
 * words, structure-of-arrays with the world index fastest (row stride Wp = W rounded up to 128), so a warp's 32
 * worlds read 256 contiguous bytes per row.  The caller allocates the slab (e.g. a torch tensor), the library
 * only borrows the pointer.  All calls return 0 on success or a negative FIZZ_E_* code; fizz_last_error()
 * gives the message.  There is no CPU fallback: without a CUDA device every device call fails.
 */
#ifndef FIZZ_B200_H
#define FIZZ_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FIZZ_ABI_VERSION 2

#define FIZZ_MAX_FREE 8          /* free polygons per world (kernel instantiations F = 1..8)        */
#define FIZZ_MAX_FIXED 16        /* fixed polygons per scene                                         */
#define FIZZ_MAX_POLY_VERTS 16   /* vertices per polygon                                             */
#define FIZZ_MAX_FIXED_VERTS 96  /* total fixed vertices per scene (staged in shared memory)         */
#define FIZZ_MAX_HITS 32         /* contact tuples one check_collisions() may return per world       */

/* error codes */
#define FIZZ_OK 0
#define FIZZ_E_ARG (-1)     /* bad argument / unsupported shape */
#define FIZZ_E_CUDA (-2)    /* CUDA runtime error (no device, launch failure, ...) */
#define FIZZ_E_STATE (-3)   /* call sequence error (e.g. step before upload) */

/* execution modes (fizz_set_tuning) */
#define FIZZ_MODE_HYBRID 1  /* ballistic thread-per-world kernel + latency-tuned warp-per-world contact kernel per chunk
                               of timesteps; worlds found trapped leave the chunk pipeline and are stepped to the end
                               of the fizz_step() call by "runner" launches on library-owned streams, so that the
                               longest sequential chain never waits at a chunk boundary (default)             */
#define FIZZ_MODE_HYBRID_SYNC 3 /* the same without runners: every world stops at every chunk boundary      */
#define FIZZ_MODE_THREAD 0  /* one thread-per-world kernel with the full semantics                        */
#define FIZZ_MODE_THREAD_FP32 4 /* OPTIONAL single-precision mode: the thread-per-world kernel with every operation of
                               update() carried out in FP32 (state still stored as FP64 in the slab, converted on load
                               and store).  Not bit-compatible with the reference: positions follow the FP64 path to
                               ~1e-6 relative until the first contact decision differs (measured horizon) */
#define FIZZ_MODE_HYBRID3 2 /* throughput-tuned: adds a thread-per-world "separation prover" tier and splits the contact
                               kernel into a compact high-occupancy general instance and a trapped instance that run
                               side by side: best for busy scenes (use small chunks, e.g. 32)                   */
/* kernel ids (fizz_kernel_info) */
#define FIZZ_KERNEL_THREAD 0
#define FIZZ_KERNEL_BALLISTIC 1
#define FIZZ_KERNEL_CONTACT 2
#define FIZZ_KERNEL_PROVER 3
#define FIZZ_KERNEL_TRAPPED 4

/* per-world status word (slab section `status`) */
#define FIZZ_ST_RUNNING 0
#define FIZZ_ST_CAPPED 1    /* resolve loop wanted more than max_calls check_collisions in one update() */
#define FIZZ_ST_NONFINITE 2 /* a centre of mass became NaN/inf; world frozen at end of launch   */
#define FIZZ_ST_HIT_OVERFLOW 3 /* more than FIZZ_MAX_HITS simultaneous contact tuples; world frozen */

/* Tunables of the reference read at call time inside update() (physics.py:19-25; set by `PARAM` lines,
 * main.py:39-41) plus the two constructor arguments of World that matter (physics.py:32-33).  The three
 * derived fields are evaluated by the HOST in Python so that they are bit-identical to the Python float
 * expressions of the reference (physics.py:188,191,516). */
typedef struct FizzParams {
    double elasticity;     /* ELASTICITY  */
    double epsilon;        /* EPSILON     */
    double time_tol;       /* TIME_TOL    */
    double collision_tol;  /* COLLISION_TOL */
    double vibrate_tol;    /* VIBRATE_TOL */
    double time_disc;      /* World.time_disc (0.05) */
    double gx, gy;         /* World.global_accel / GRAVITY; (0,0) for every .in scene */
    double one_plus_e;     /* (1+ELASTICITY)                */
    double one_minus_e2;   /* (1-ELASTICITY**2)             */
    double acc_x, acc_y;   /* sum([1*global_accel])/1  -> the acceleration Point.move() installs (physics.py:516) */
    double height;         /* World.height, for p_energy (physics.py:373) */
    /* default-off extensions (absent from the reference, defined in csrc/fizz_ext.cuh): */
    double friction_mu;    /* Coulomb coefficient of every contact; 0 = off (the reference has no friction) */
    double ff_restitution; /* restitution between two FREE bodies (notes/collisions.tex:24-35); < 0 = off: the reference's
                              perfectly elastic exchange that ignores ELASTICITY (physics.py:201-223) */
    int32_t max_calls;     /* watchdog: state-relevant check_collisions() calls per update(); >= 32 */
    int32_t reserved;
} FizzParams;

/* Scene topology + fixed geometry of one group (host pointers, read during fizz_group_create only).
 * Replaces the object graph main.read_input() builds (main.py:23-83): `world.objs`, `world.fixed_objs`. */
typedef struct FizzGroupDesc {
    int32_t n_free;              /* F: len(world.objs)       */
    int32_t n_fixed;             /* X: len(world.fixed_objs) */
    const int32_t *nverts;       /* [F+X] vertices per polygon, free bodies first (physics.py:270 order) */
    const double *fixed_xy;      /* fixed vertices x,y interleaved, polygons concatenated (FixedPolygon.points) */
    const double *fixed_com;     /* [X][2] centroids (physics.py:314)  */
    const double *fixed_thr;     /* [X] broad-phase thresholds T(radius): largest s with sqrt(s) <= radius */
    FizzParams params;
} FizzGroupDesc;

/* Word offsets (units of 8 bytes) of every section inside the device slab.  Row r of a section starts at
 * offset + r*Wp; element w of the row is world w.  Doubles unless noted. */
typedef struct FizzLayout {
    int64_t n_worlds;   /* W  */
    int64_t stride;     /* Wp */
    int32_t n_free, n_fixed, n_free_verts, reserved;
    /* dynamic state (contiguous: [pos .. calls]) */
    int64_t pos;        /* [F*2] com.pos  x,y per free body        (physics.py:315)  */
    int64_t vel;        /* [F*2] com.vel                                              */
    int64_t heat;       /* [1]   world.heat                        (physics.py:53,195) */
    int64_t status;     /* [1]   int64 FIZZ_ST_*                                       */
    int64_t steps;      /* [1]   int64 completed update() calls                        */
    int64_t calls;      /* [1]   int64 state-relevant check_collisions() calls, lifetime */
    int64_t vflag;      /* [1]   int64 1 = `verts` holds explicit point.pos (pushed by a resolve pass), 0 = derived */
    int64_t tier;       /* [1]   int64 scheduling hint, bits 0..7: 0 quiescent, 1 candidates/no contact, 2 contact phase,
                                 3 trapped, 4 owned by a runner launch; bits 8..31: check_collisions() of the last step */
    int64_t state_words;/* words from `pos` to the end of `tier` */
    /* per-world constants */
    int64_t off;        /* [V*2] |r|cos(phi), |r|sin(phi) per free vertex (physics.py:358-359,632-635) */
    int64_t thr;        /* [F]   T(obj.radius) per free body        (physics.py:322,277-278) */
    int64_t mass;       /* [F]   obj.mass = area*density            (physics.py:624) */
    /* outputs */
    int64_t verts;      /* [V*2] point.pos of every free vertex after the last completed launch */
    int64_t energy;     /* [4]   World.energy(): total,kinetic,potential,heat (physics.py:66-74) */
    /* library scratch (never touched by the caller) */
    int64_t sched;      /* [2]+16 contact-kernel work lists, counts, cursors, counters */
    int64_t scene;      /* fixed-scene constants + vertex->body table     */
    int64_t total_words;
} FizzLayout;

typedef struct FizzGroup FizzGroup;

/* one recorded contact tuple (tests / debugging): the tuples check_collisions() returns (physics.py:288) */
typedef struct FizzContact {
    int64_t world;
    int32_t step;   /* 0-based update() index                                  */
    int32_t call;   /* 1-based state-relevant check_collisions() call in step  */
    int32_t i, j;   /* indices into objs+fixed_objs                            */
    double nx, ny, mag;
} FizzContact;

int fizz_abi_version(void);
/* The numpy 2-vector dot/norm contract this build of the library reproduces (csrc/fizz_common.cuh):
 * 1 = fma(a1, b1, a0*b0) (default: the golden host's OpenBLAS), 2 = fma(a0, b0, a1*b1), 0 = unfused.  One library per
 * mode (libfizz_b200.so, libfizz_b200_dot0.so, libfizz_b200_dot2.so); the Python loader picks by probing numpy. */
int fizz_dot_mode(void);
const char *fizz_last_error(void);

/* Host helpers implementing numpy's 2-vector contract for the Python loader's init-time math:
 *   norm2: sqrt(fma(y,y,x*x))   == np.linalg.norm([x,y])      (physics.py:322,633,358)
 *   dot2 : fma(a1,b1,a0*b0)     == np.dot([a0,a1],[b0,b1])    */
void fizz_host_norm2(int64_t n, const double *x, const double *y, double *out);
void fizz_host_dot2(int64_t n, const double *a0, const double *a1, const double *b0, const double *b1, double *out);
/* T(r): largest double s such that sqrt(s) <= r, so that `norm(d) > r` (physics.py:278) == `fma(dy,dy,dx*dx) > T(r)` */
void fizz_host_sqrt_threshold(int64_t n, const double *r, double *out);

/* Slab layout for a group of n_worlds worlds.  Pure host arithmetic. */
int fizz_layout(const FizzGroupDesc *desc, int64_t n_worlds, FizzLayout *out);

/* Bind a group to a caller-owned device slab of at least layout.total_words 8-byte words on `device`.
 * Stands in for constructing `physics.World(width,height)` + its objects (main.py:28,68-77). */
int fizz_group_create(const FizzGroupDesc *desc, int64_t n_worlds, int device, void *dev_slab, int64_t slab_words,
                      FizzGroup **out);
int fizz_group_destroy(FizzGroup *g);

/* Host -> device initial conditions and per-world constants, SoA rows of length W (not Wp) in host memory:
 * pos[F*2][W], vel[F*2][W], off[V*2][W], thr[F][W], mass[F][W].  Also zeroes heat/status/steps/calls.
 * Asynchronous on `stream` (a cudaStream_t; 0 = default stream); host buffers must stay valid until the
 * stream is synchronised (use pinned memory for true overlap). */
int fizz_upload(FizzGroup *g, const double *pos, const double *vel, const double *off, const double *thr,
                const double *mass, void *stream);

/* Advance every running world by n_steps World.update() calls (physics.py:76-259); stands in for the loop
 * `for t in range(N): plane.update()` of main.py:97-98.  Asynchronous on `stream`, no host synchronisation.
 * Hybrid mode: per chunk of timesteps one ballistic launch (thread per world, registers only), one work-list
 * launch and one persistent contact launch (warp per world).  Thread mode: one launch for all n_steps.
 * FIZZ_MODE_HYBRID, n_steps >= 128: worlds found trapped are additionally handed to "runner" launches on streams the
 * library owns; each of them is forked from `stream` by an event and joined back into `stream` before the call
 * returns, so whatever the caller enqueues on `stream` afterwards sees every world at its target.
 *
 * STREAM CONTRACT.  A group is stepped, uploaded, restored and read on ONE stream at a time: the library-owned runner
 * streams, runner work lists and events of a group are reused by the next fizz_step() call on the strength of the
 * previous call's join INTO ITS `stream`.  Consecutive calls for one group must therefore be issued on the same stream,
 * or on streams the caller has ordered with an event (the second call's stream waits for the first call's work);
 * calls for DIFFERENT groups may run on different streams concurrently.  The calls are not thread-safe per group. */
int fizz_step(FizzGroup *g, int64_t n_steps, void *stream);

/* Streaming variant: like fizz_step, but before every timestep t the World.energy() row of EVERY world
 * -- what main.py appends to energy.txt (physics.py:81-82) -- is written to dev_energy_log[t][4][Wp] (device memory,
 * n_steps * 4 * layout.stride doubles).  One launch set per timestep; HBM-bound output of 32 B per world-step. */
int fizz_step_logged(FizzGroup *g, int64_t n_steps, double *dev_energy_log, void *stream);

/* Execution mode and timesteps per chunk (hybrid mode).  Results never depend on either. */
int fizz_set_tuning(FizzGroup *g, int32_t mode, int64_t chunk_steps);

/* Save / restore the dynamic state block (layout.state_words words starting at layout.pos) to / from a
 * caller-owned device buffer, e.g. to replay a batch from its initial conditions without a host round trip.
 * `steps_done` = the number of completed update() calls the restored state corresponds to. */
int fizz_snapshot(FizzGroup *g, void *dev_dst, void *stream);
int fizz_restore(FizzGroup *g, const void *dev_src, int64_t steps_done, void *stream);

/* Make the slab's `verts` section current for every world (derives off+com where vflag == 0). Asynchronous. */
int fizz_vertices(FizzGroup *g, void *stream);

/* World.energy() for every world into the slab's `energy` section (physics.py:66-74). Asynchronous. */
int fizz_energy(FizzGroup *g, void *stream);

/* Device -> host: any pointer may be NULL.  pos/vel [F*2][W], heat [W], energy [4][W] (runs fizz_energy first),
 * verts [V*2][W], status/steps/calls int64 [W].  Asynchronous on `stream`. */
int fizz_download(FizzGroup *g, double *pos, double *vel, double *heat, double *energy, double *verts,
                  int64_t *status, int64_t *steps, int64_t *calls, void *stream);

/* Optional contact trace for parity tests: device buffer of `capacity` FizzContact records + a device int64
 * counter (both caller-owned); pass NULL to disable.  The counter may exceed capacity (overflow is dropped). */
int fizz_set_trace(FizzGroup *g, void *dev_records, int64_t capacity, void *dev_counter);

/* Optional per-world introspection: dev_general_calls = device int64 [n_worlds] (caller-owned, zeroed by the caller) that
 * accumulates, per world, the check_collisions() calls the contact kernel made OUTSIDE its single-contact fast path -- the
 * expensive ones; NULL switches it off.  (Which world bounds a pass?  tools/shard_probe.py) */
int fizz_set_world_stats(FizzGroup *g, void *dev_general_calls);

/* Introspection for bench/roofline reporting: `which` = FIZZ_KERNEL_*; launches = kernels launched so far. */
int fizz_kernel_info(FizzGroup *g, int32_t which, int32_t *regs_per_thread, int32_t *smem_bytes_per_block,
                     int32_t *block_threads, int32_t *blocks_per_sm, int64_t *launches);

/* Optional per-kernel device timing (CUDA events on the launching stream, recorded inside fizz_step).
 * fizz_profile_read synchronises on the recorded events and returns accumulated milliseconds and launch counts
 * indexed by FIZZ_KERNEL_* (the work-list kernel is counted with the contact kernel). */
int fizz_set_profiling(FizzGroup *g, int32_t enabled);
int fizz_profile_read(FizzGroup *g, double *ms3, int64_t *launches3, int32_t reset);

/* Work counters accumulated on the device since the last reset (synchronises `stream`):
 * out4[0] world-steps advanced by the ballistic kernel, [1] by the contact kernel, [2] check_collisions() calls made
 * by the contact kernel, [3] of those inside its single-contact fast path.  Thread mode does not count. */
int fizz_counters(FizzGroup *g, int64_t *out4, int32_t reset, void *stream);

/* Spec rounds (mode FIZZ_MODE_HYBRID, calls of >= 128 timesteps, contact trace off): expensive worlds -- a trapped body
 * makes hundreds of resolve passes per timestep, forever -- are stepped TIME-PARALLEL: K warps evaluate K consecutive
 * pieces of one world's future from predicted start states, and a piece is accepted only if the state it started from
 * equals, bit for bit, the state the piece before it ended in.  Results never depend on any of these settings (a wrong
 * prediction is thrown away and evaluated again); they trade latency of the longest chains against wasted evaluations.
 * enabled: 0 off, 1 on, < 0 keep (default on; $FIZZ_SPEC=0 at group creation turns it off).  The others, <= 0 keep:
 * window0 = timesteps a round of a world without history tries to cover (32); slot_cost = cost a slot aims at, in units
 * of 256 clock cycles of one warp (200); min_cost = cost per timestep from which a world takes part (48). */
int fizz_set_spec(FizzGroup *g, int32_t enabled, int32_t window0, int32_t slot_cost, int32_t min_cost);

/* What the spec rounds did since the last reset (synchronises `stream`): out4[0] timesteps accepted, [1] timesteps
 * evaluated (accepted + thrown away), [2] rounds committed, [3] rounds that threw at least one slot away. */
int fizz_spec_stats(FizzGroup *g, int64_t *out4, int32_t reset, void *stream);

/* FP64 FMA throughput microbenchmark (the roofline denominator MEASURED_PEAKS.json lacks): runs `iters`
 * dependent-chain DFMA rounds on every SM of `device` and returns achieved TFLOP/s (FMA = 2 flops). */
int fizz_fp64_peak(int device, int32_t iters, double *tflops, double *ms);

/* ---- batched rasteriser ------------------------------------------------------------------
 * Replaces the reference's only native program, /root/reference/src/draw.c (`./draw START END`: spec_to_image
 * :189-221, draw_polygon :116-157, draw_line :87-114, fill_shape :159-187) and the grey scaling of
 * png_util.c:238-262, for a BATCH of frames that share one topology (the frames of one world over time, or of many
 * worlds of a group at one time).  One CTA per frame; the canvas lives in shared memory as a bitmap.
 *
 * points_dev: device, int32 [n_frames][n_points][2] = (x, y) of every polygon vertex in file order, n_points =
 * sum(nverts[0..n_polygons-1]) (host array).  gray_dev: device, uint8 [n_frames][height][width], 255 = FILLED, 0 = EMPTY
 * (the value the reference writes to all three channels of its 8-bit RGB PNG).  Fails (FIZZ_E_ARG) when width > 1024
 * (a row of the bitmap is one warp of 32-bit words) or when the two bitmaps of a frame do not fit the shared memory of
 * an SM: 2 * height * ceil(width/32) * 4 bytes <= 227 KB (the reference's largest scene is 1010 x 710: 182 KB). */
#define FIZZ_RASTER_MAX_POLYGONS 64
int fizz_raster_frames(int device, int32_t width, int32_t height, int64_t n_frames, int32_t n_polygons,
                       const int32_t *nverts, const int32_t *points_dev, uint8_t *gray_dev, void *stream);

/* Point.__str__ (physics.py:600-606) on the device: verts_dev = double [n_frames][n_free_points][2] (point.pos of the
 * free polygons) -> int() truncation, clamp to [0,width-1] x [0,height-1]; the n_fixed_points integer points of the
 * fixed polygons (fixed_points_dev, int32 [n_fixed_points][2], may be NULL when 0) are appended to every frame, the
 * order World.__str__ (physics.py:291-301) prints.  points_dev: int32 [n_frames][n_free_points + n_fixed_points][2]. */
int fizz_points_from_vertices(int device, int32_t width, int32_t height, int64_t n_frames, int32_t n_free_points,
                              const double *verts_dev, int32_t n_fixed_points, const int32_t *fixed_points_dev,
                              int32_t *points_dev, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* FIZZ_B200_H */
