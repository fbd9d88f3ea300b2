/* TEST INFRASTRUCTURE ONLY -- CPU oracle for the Fizz-2D stepping hot path.
 *
 * A literal, scalar, one-world-at-a-time restatement in plain C of `World.update()` and its callees in
 * the reference (/root/reference/src/physics.py:66-289, :384-464, :502-529, :765-840).  Only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load it.  The shipped
 * engine (fizz2d_b200/csrc) never links or calls it.
 *
 * Parity status: PINNED -- checked bit-for-bit against traces of the reference itself (run in the build
 * container through oracle/ref_harness.py) that are committed under tests/golden/.  The reference has no
 * tests or golden vectors of its own.
 * The two default-off EXTENSIONS (friction_mu, ff_restitution) have no reference at all: for them this file is the
 * definition the CUDA kernels are checked against ("parity unpinned" by construction; with both off, nothing changes).
 */
#ifndef FIZZ_ORACLE_H
#define FIZZ_ORACLE_H

#ifdef __cplusplus
extern "C" {
#endif

#define FZO_MAXV 16 /* max vertices per polygon */

typedef struct {
    double elasticity;    /* physics.ELASTICITY   (physics.py:22)  */
    double epsilon;       /* physics.EPSILON      (physics.py:19)  */
    double time_tol;      /* physics.TIME_TOL     (physics.py:23)  */
    double collision_tol; /* physics.COLLISION_TOL(physics.py:24)  */
    double vibrate_tol;   /* physics.VIBRATE_TOL  (physics.py:25)  */
    double time_disc;     /* World.time_disc      (physics.py:16,33) */
    double gx, gy;        /* World.global_accel / GRAVITY (physics.py:15,32); zero from .in files */
    double one_plus_e;    /* (1+ELASTICITY)      evaluated by Python on the host (physics.py:188) */
    double one_minus_e2;  /* (1-ELASTICITY**2)   evaluated by Python on the host (physics.py:191) */
    double height;        /* World.height, for p_energy (physics.py:373) */
    /* default-off extensions -- NOT in the reference, defined by this project (see fizz_oracle.c "extensions"): */
    double friction_mu;   /* Coulomb coefficient, 0 = off */
    double ff_restitution;/* restitution between two free bodies, < 0 = off (reference: elastic exchange) */
    int max_calls;        /* watchdog: state-relevant check_collisions calls per update()       */
} fzo_params;

typedef struct {
    int step, call, i, j; /* 0-based step, 1-based state-relevant call index within the step, pair */
    double nx, ny, mag;
} fzo_contact;

typedef struct fzo_world fzo_world;

/* Bodies: `nfree` free polygons (file order) then `nfixed` fixed ones.
 * nverts[nfree+nfixed]; com[2*(nfree+nfixed)] centroids (physics.py:314); vel[2*nfree];
 * radius[nfree+nfixed] (physics.py:322); mass[nfree] (physics.py:624);
 * offx/offy: per free vertex |r|cos(phi), |r|sin(phi) (physics.py:358-359,632-635), concatenated;
 * fixedxy: fixed vertices x,y interleaved, concatenated. */
fzo_world *fzo_create(int nfree, int nfixed, const int *nverts, const double *com, const double *vel,
                      const double *radius, const double *mass, const double *offx, const double *offy,
                      const double *fixedxy, const fzo_params *prm);
void fzo_destroy(fzo_world *w);

/* One World.update() (physics.py:76-259). Returns 0 ok, 1 if the world is/became capped (frozen). */
int fzo_update(fzo_world *w);
/* n updates; stops early when capped. Returns number of completed steps (world lifetime total). */
long fzo_run(fzo_world *w, long nsteps);

/* World.energy() (physics.py:66-74): out[4] = total, kinetic, potential, heat. */
void fzo_energy(const fzo_world *w, double *out);
/* out[4*nfree] = com.pos.x, com.pos.y, com.vel.x, com.vel.y per free body. */
void fzo_com(const fzo_world *w, double *out);
/* current vertex positions of all free bodies, x,y interleaved, concatenated. */
void fzo_verts(const fzo_world *w, double *out);
double fzo_heat(const fzo_world *w);
long fzo_steps_done(const fzo_world *w);
int fzo_status(const fzo_world *w);       /* 0 running, 1 capped */
long fzo_calls_last_step(const fzo_world *w);
long fzo_total_calls(const fzo_world *w);  /* state-relevant check_collisions calls, lifetime */
double fzo_total_ops(const fzo_world *w);  /* algorithmic FP64 ops (SURVEY 8d rules), lifetime */

/* contact trace: attach a caller-owned buffer; returns through fzo_contact_count how many were produced
 * (may exceed cap; only the first `cap` are stored). */
void fzo_trace(fzo_world *w, fzo_contact *buf, long cap);
long fzo_contact_count(const fzo_world *w);

/* The numpy 2-vector dot/norm contract this build follows: 1 = fma(a1,b1,a0*b0) (default), 2, 0. */
int fzo_dot_mode(void);

/* Batch helper for the CPU baseline: step `n` worlds `nsteps` each on `nthreads` pthreads. */
void fzo_run_batch(fzo_world **worlds, long n, long nsteps, int nthreads);

#ifdef __cplusplus
}
#endif
#endif
