// fizz_engine.cu -- host side of the C-ABI (include/fizz_b200.h) + small utility kernels.
//
// Execution modes per group (fizz_set_tuning):
//   FIZZ_MODE_HYBRID (default): per chunk of timesteps, (1) ballistic_kernel<F> -- thread per world, registers
//     only, advances quiescent steps and stops a world at the first non-empty broad phase; (2) classify_kernel
//     builds the list of worlds short of the chunk target; (3) contact_kernel -- persistent, warp per world,
//     cooperative SAT / resolution -- brings those worlds to the chunk target.  No host synchronisation.
//     Calls of >= 128 timesteps: in every chunk whose work list is short (the scene has calmed down), the worlds whose
//     timesteps are expensive -- trapped bodies: hundreds of resolve passes per timestep, forever -- are stepped
//     TIME-PARALLEL by the "spec rounds" (fizz_common.cuh: SpecParams; $FIZZ_SPEC=0 turns them off): K warps evaluate K
//     consecutive pieces of a world's future from predicted start states, a prediction is accepted only if it equals the
//     real state bit for bit.  The few expensive worlds no predictor fits leave the pipeline at a chunk boundary: "runner"
//     launches of the contact kernel on library-owned streams step them to the end of the fizz_step() call.
//   FIZZ_MODE_HYBRID_SYNC: the plain chunk pipeline (no spec rounds, no runner launches).
//   FIZZ_MODE_HYBRID3: adds the separation-prover tier and runs two contact-kernel instances side by side.
//   FIZZ_MODE_THREAD: step_kernel<F, double> -- thread per world, full semantics in one launch.
//   FIZZ_MODE_THREAD_FP32: step_kernel<F, float> -- the same operation sequence in single precision (optional mode).
// Also here: the batched rasteriser's entry points (fizz_raster.cuh).
#include "fizz_b200.h"

#include "fizz_ballistic_kernel.cuh"
#include "fizz_common.cuh"
#include "fizz_contact_kernel.cuh"
#include "fizz_raster.cuh"
#include "fizz_thread_kernel.cuh"

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

using namespace fizz;

namespace {

// World.energy (physics.py:66-74): K = sum (.5*m)*norm(v)**2 ; P = sum (GRAVITY[1]*m)*(height - y) ; H = world.heat
__global__ void energy_kernel(const double *pos, const double *vel, const double *heat, const double *mass,
                              double *energy, long long W, long long Wp, int F, double gy, double height) {
    const long long w = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (w >= W) return;
    double K = 0.0, P = 0.0;
    for (int b = 0; b < F; b++) {
        const double m = mass[b * Wp + w];
        const double vx = vel[(2 * b) * Wp + w], vy = vel[(2 * b + 1) * Wp + w];
        const double nv = sqrt(sq2(vx, vy));
        K += (.5 * m) * (nv * nv);
        P += (gy * m) * (height - pos[(2 * b + 1) * Wp + w]);
    }
    const double H = heat[w];
    energy[0 * Wp + w] = ((0.0 + K) + P) + H;
    energy[1 * Wp + w] = K;
    energy[2 * Wp + w] = P;
    energy[3 * Wp + w] = H;
}

// FP64 FMA throughput probe: 8 independent chains per thread, fully unrolled.
__global__ void fp64_peak_kernel(double *out, int iters, double a, double b) {
    double x0 = threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int u = 0; u < 8; u++) {
            x0 = fma(x0, a, b); x1 = fma(x1, a, b); x2 = fma(x2, a, b); x3 = fma(x3, a, b);
            x4 = fma(x4, a, b); x5 = fma(x5, a, b); x6 = fma(x6, a, b); x7 = fma(x7, a, b);
        }
    }
    const double s = ((x0 + x1) + (x2 + x3)) + ((x4 + x5) + (x6 + x7));
    if (s == 12345.678) out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}


// ------------------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------------------
thread_local std::string g_err;

int fail(int code, const std::string &msg) {
    g_err = msg;
    return code;
}
int cuda_fail(cudaError_t e, const char *what) {
    return fail(FIZZ_E_CUDA, std::string(what) + ": " + cudaGetErrorString(e));
}
#define CU(call)                                            \
    do {                                                    \
        cudaError_t e_ = (call);                            \
        if (e_ != cudaSuccess) return cuda_fail(e_, #call); \
    } while (0)

inline long long round_up(long long x, long long m) { return (x + m - 1) / m * m; }

// dynamic shared memory of a runner launch: enough that a runner CTA has its SM to itself (see fizz_step)
constexpr size_t RUNNER_SMEM = 200 * 1024;

struct KernelInfo {
    cudaFuncAttributes attr;
    int blocks_per_sm;
};

// cudaFuncAttributeMaxDynamicSharedMemorySize is per function and process-wide: it must only ever grow, or a group
// with a smaller footprint created later would make the launches of an earlier, larger group invalid.
// (`current` is an array of MAX_DEVICES entries: the attribute belongs to the function ON A DEVICE, and one process may
//  drive several)
constexpr int MAX_DEVICES = 64;
template <typename K>
cudaError_t raise_dynamic_smem(K kernel, size_t smem, size_t *current) {
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    size_t *cur = current + (dev >= 0 && dev < MAX_DEVICES ? dev : 0);
    const size_t want = smem > 48 * 1024 ? smem : 48 * 1024;
    if (want <= *cur && dev < MAX_DEVICES) return cudaSuccess;
    e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)want);
    if (e == cudaSuccess) *cur = want;
    return e;
}

template <int F>
cudaError_t prep_thread(size_t smem, KernelInfo *ki) {
    static size_t current[MAX_DEVICES] = {}, current32[MAX_DEVICES] = {};
    cudaError_t e = raise_dynamic_smem(step_kernel<F, double>, smem, current);
    if (e == cudaSuccess) e = raise_dynamic_smem(step_kernel<F, float>, smem, current32);   // (same size: simpler)
    if (e != cudaSuccess) return e;
    e = cudaFuncGetAttributes(&ki->attr, step_kernel<F, double>);
    if (e != cudaSuccess) return e;
    return cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ki->blocks_per_sm, step_kernel<F, double>, BLOCK, smem);
}
// The lean variant must work; the prover variant (FIZZ_MODE_HYBRID3 only) needs shared memory that a vertex-rich scene
// may not leave: then only that mode is unavailable (*prover_ok = false), the group is still created.
template <int F>
cudaError_t prep_ballistic(size_t smem, KernelInfo *ki, KernelInfo *kp_prover, bool *prover_ok) {
    static size_t current[MAX_DEVICES] = {};
    cudaError_t e = cudaFuncGetAttributes(&ki->attr, ballistic_kernel<F, false>);
    if (e != cudaSuccess) return e;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ki->blocks_per_sm, ballistic_kernel<F, false>, BALLISTIC_BLOCK, 0);
    if (e != cudaSuccess) return e;
    *prover_ok = false;
    std::memset(kp_prover, 0, sizeof(*kp_prover));
    if (raise_dynamic_smem(ballistic_kernel<F, true>, smem, current) == cudaSuccess &&
        cudaFuncGetAttributes(&kp_prover->attr, ballistic_kernel<F, true>) == cudaSuccess &&
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&kp_prover->blocks_per_sm, ballistic_kernel<F, true>, BALLISTIC_BLOCK,
                                                      smem) == cudaSuccess &&
        kp_prover->blocks_per_sm >= 1)
        *prover_ok = true;
    else
        (void)cudaGetLastError();
    return cudaSuccess;
}

#define FIZZ_DISPATCH_F(F_, CALL)          \
    switch (F_) {                          \
        case 1: { constexpr int FF = 1; CALL; } break; \
        case 2: { constexpr int FF = 2; CALL; } break; \
        case 3: { constexpr int FF = 3; CALL; } break; \
        case 4: { constexpr int FF = 4; CALL; } break; \
        case 5: { constexpr int FF = 5; CALL; } break; \
        case 6: { constexpr int FF = 6; CALL; } break; \
        case 7: { constexpr int FF = 7; CALL; } break; \
        default: { constexpr int FF = 8; CALL; } break; \
    }

}  // namespace

struct ProfSpan {
    int kind;  // FIZZ_KERNEL_*
    cudaEvent_t a, b;
};

struct FizzGroup {
    FizzLayout lay;
    FizzParams prm;
    KParams kp;
    int device;
    int sm_count;
    double *slab;
    bool uploaded;
    long long target;
    long long launches;
    int mode;
    long long chunk;
    // thread mode
    size_t thread_smem;
    KernelInfo thread_ki;
    bool thread_ok;
    bool prover_ok;   // ballistic_kernel<F, true> fits (FIZZ_MODE_HYBRID3)
    // hybrid mode
    KernelInfo ballistic_ki, prover_ki, contact_ki, trapped_ki;
    size_t contact_smem, ballistic_smem;
    long long *sched_list, *sched_list2;
    cudaStream_t side;          // library-owned stream: the trapped instance runs beside the general one
    cudaEvent_t ev_fork, ev_join;
    // runner mode: trapped worlds leave the chunk pipeline and are stepped to the end of the fizz_step() call by their
    // own launch of the latency-tuned instance, one library-owned stream per launch (no launch waits for another one)
    static const int NR = 16;   // runner launches per fizz_step() call (one stream each; further boundaries launch none)
    cudaStream_t rstream[NR];
    cudaEvent_t rdone[NR];
    long long *rlist;            // [NR][rcap] work lists (engine-owned, outside the caller's slab)
    long long rcap;              // entries per list: one warp per world, one CTA per SM, a third of the SMs at most
                                 // (runner CTAs keep their SMs to themselves: the chunk pipeline needs the rest)
    unsigned long long *rcount;  // [NR][2] count, cursor
    // spec rounds (fizz_common.cuh: SpecParams): time-parallel stepping of the heavy trapped worlds, chunk by chunk
    bool spec_on;
    int spec_S, spec_D, spec_cost, spec_budget;
    long long scap;              // entries of the spec list
    long long *spec_list;        // [scap]                       (all of it allocated on first use)
    double *spec_slots;          // [scap][SPEC_K][slot words]
    long long *spec_ctrl;        // [W][SPEC_CTRL_WORDS]
    unsigned long long *spec_state;   // [0] list length, [1] [2] [3] the rounds' queue: head, tail, worlds in flight, [4] unused cursor
    unsigned long long *spec_queue;   // [scap * SPEC_K] tickets
    int spec_ctas;                    // CTAs per SM of the rounds' launch
    cudaStream_t spec_stream;
    cudaEvent_t spec_fork, spec_join;
    static const int SPEC_RMAX = 64;     // (debug statistics: chunks)
    unsigned long long *sched_count;  // [0] = count, [1] = cursor
    unsigned char *vbody_dev;
    // optional per-kernel timing (fizz_set_profiling)
    bool profile;
    std::vector<cudaEvent_t> ev_pool;
    size_t ev_used;
    std::vector<ProfSpan> spans;
    double prof_ms[3];   // accumulated: thread, ballistic, classify+contact
    long long prof_launches[3];
    cudaError_t prof_error;  // first failure of a timing event (fizz_step reports it)
};

// Next timing event of the pool, recorded on `st`; nullptr (and the error latched for fizz_step to report) when the
// runtime refuses to create or record one.
static cudaEvent_t next_event(FizzGroup *g, cudaStream_t st) {
    if (g->ev_used == g->ev_pool.size()) {
        cudaEvent_t e = nullptr;
        const cudaError_t rc = cudaEventCreate(&e);
        if (rc != cudaSuccess) { g->prof_error = rc; return nullptr; }
        g->ev_pool.push_back(e);
    }
    cudaEvent_t e = g->ev_pool[g->ev_used++];
    const cudaError_t rc = cudaEventRecord(e, st);
    if (rc != cudaSuccess) { g->prof_error = rc; return nullptr; }
    return e;
}

extern "C" {

int fizz_abi_version(void) { return FIZZ_ABI_VERSION; }
int fizz_dot_mode(void) { return FIZZ_DOT_MODE; }
const char *fizz_last_error(void) { return g_err.c_str(); }

void fizz_host_norm2(int64_t n, const double *x, const double *y, double *out) {
    for (int64_t i = 0; i < n; i++) out[i] = std::sqrt(sq2(x[i], y[i]));
}
void fizz_host_dot2(int64_t n, const double *a0, const double *a1, const double *b0, const double *b1, double *out) {
    for (int64_t i = 0; i < n; i++) out[i] = dot2(a0[i], a1[i], b0[i], b1[i]);
}
void fizz_host_sqrt_threshold(int64_t n, const double *r, double *out) {
    for (int64_t i = 0; i < n; i++) {
        const double ri = r[i];
        if (!(ri >= 0.0) || std::isinf(ri)) { out[i] = ri; continue; }  // NaN / negative / inf: pass through
        double s = ri * ri;
        while (std::sqrt(s) > ri) s = std::nextafter(s, 0.0);
        for (;;) {
            const double up = std::nextafter(s, INFINITY);
            if (std::isinf(up) || std::sqrt(up) > ri) break;
            s = up;
        }
        out[i] = s;
    }
}

int fizz_layout(const FizzGroupDesc *d, int64_t n_worlds, FizzLayout *out) {
    if (!d || !out) return fail(FIZZ_E_ARG, "fizz_layout: null argument");
    if (d->n_free < 1 || d->n_free > FIZZ_MAX_FREE) return fail(FIZZ_E_ARG, "fizz_layout: n_free must be 1..8");
    if (d->n_fixed < 0 || d->n_fixed > FIZZ_MAX_FIXED) return fail(FIZZ_E_ARG, "fizz_layout: n_fixed must be 0..16");
    if (n_worlds < 1) return fail(FIZZ_E_ARG, "fizz_layout: n_worlds must be >= 1");
    int V = 0, FV = 0;
    for (int b = 0; b < d->n_free + d->n_fixed; b++) {
        const int n = d->nverts[b];
        if (n < 3 || n > FIZZ_MAX_POLY_VERTS) return fail(FIZZ_E_ARG, "fizz_layout: polygons need 3..16 vertices");
        if (b < d->n_free) V += n; else FV += n;
    }
    if (FV > FIZZ_MAX_FIXED_VERTS) return fail(FIZZ_E_ARG, "fizz_layout: too many fixed vertices (max 96)");
    FizzLayout L;
    std::memset(&L, 0, sizeof(L));
    const int64_t Wp = round_up(n_worlds, 128);
    const int F = d->n_free;
    L.n_worlds = n_worlds; L.stride = Wp; L.n_free = F; L.n_fixed = d->n_fixed; L.n_free_verts = V;
    int64_t o = 0;
    L.pos = o; o += 2 * F * Wp;
    L.vel = o; o += 2 * F * Wp;
    L.heat = o; o += Wp;
    L.status = o; o += Wp;
    L.steps = o; o += Wp;
    L.calls = o; o += Wp;
    L.vflag = o; o += Wp;
    L.tier = o; o += Wp;
    L.state_words = o;
    L.off = o; o += 2 * (int64_t)V * Wp;
    L.thr = o; o += F * Wp;
    L.mass = o; o += F * Wp;
    L.verts = o; o += 2 * (int64_t)V * Wp;
    L.energy = o; o += 4 * Wp;
    L.sched = o; o += 2 * Wp + 16;                   // two work lists + counts/cursors/counters (library scratch)
    L.scene = o;
    o += round_up((int64_t)FV * 6, 16);              // fixed scene constants (library-written)
    o += round_up(((int64_t)V + 7) / 8, 16);         // vertex -> body table
    L.total_words = o;
    *out = L;
    return FIZZ_OK;
}

int fizz_group_create(const FizzGroupDesc *d, int64_t n_worlds, int device, void *dev_slab, int64_t slab_words,
                      FizzGroup **out) {
    if (!out) return fail(FIZZ_E_ARG, "fizz_group_create: null out");
    *out = nullptr;
    FizzLayout L;
    int rc = fizz_layout(d, n_worlds, &L);
    if (rc) return rc;
    if (!dev_slab || slab_words < L.total_words) return fail(FIZZ_E_ARG, "fizz_group_create: slab too small");
    if (d->params.max_calls < 32) return fail(FIZZ_E_ARG, "fizz_group_create: max_calls must be >= 32");
    if (!(d->params.friction_mu >= 0.0)) return fail(FIZZ_E_ARG, "fizz_group_create: friction_mu must be >= 0 (0 = off)");
    if (!(d->params.ff_restitution <= 1.0)) return fail(FIZZ_E_ARG, "fizz_group_create: ff_restitution must be <= 1 (< 0 = off)");
    int ndev = 0;
    CU(cudaGetDeviceCount(&ndev));
    if (device < 0 || device >= ndev) return fail(FIZZ_E_CUDA, "fizz_group_create: no such CUDA device");
    CU(cudaSetDevice(device));
    cudaDeviceProp prop;
    CU(cudaGetDeviceProperties(&prop, device));

    FizzGroup *g = new FizzGroup();
    std::memset(&g->kp, 0, sizeof(g->kp));
    g->lay = L; g->prm = d->params; g->device = device; g->slab = (double *)dev_slab;
    g->sm_count = prop.multiProcessorCount;
    g->uploaded = false; g->target = 0; g->launches = 0;
    g->mode = FIZZ_MODE_HYBRID; g->chunk = 256;
    g->profile = false; g->ev_used = 0; g->prof_error = cudaSuccess;
    g->side = nullptr; g->ev_fork = nullptr; g->ev_join = nullptr;
    g->rlist = nullptr; g->rcount = nullptr;
    g->spec_list = nullptr; g->spec_slots = nullptr; g->spec_ctrl = nullptr; g->spec_state = nullptr; g->spec_queue = nullptr;
    g->spec_stream = nullptr; g->spec_fork = nullptr; g->spec_join = nullptr;
    {
        auto env_int = [](const char *name, long long dflt, long long lo, long long hi) {
            const char *v = std::getenv(name);
            if (!v || !*v) return dflt;
            const long long x = std::atoll(v);
            return x < lo ? lo : (x > hi ? hi : x);
        };
        g->spec_on = env_int("FIZZ_SPEC", 1, 0, 1) != 0;
        g->spec_S = (int)env_int("FIZZ_SPEC_S", 32, 2, 512);          // window of a world without history (timesteps)
        g->spec_D = (int)env_int("FIZZ_SPEC_D", 200, 1, 1 << 20);     // cost per item (~0.13 us per unit)
        g->spec_cost = (int)env_int("FIZZ_SPEC_COST", 48, 1, 1 << 20);  // worlds at least this costly per timestep
        g->spec_budget = (int)env_int("FIZZ_SPEC_BUDGET", 0, 0, 1 << 30);   // 0: D
        g->spec_ctas = (int)env_int("FIZZ_SPEC_CTAS", 2, 1, 3);       // CTAs per SM of the rounds' launch
        g->scap = env_int("FIZZ_SPEC_CAP", 4096, 1, 1 << 20);
        if (g->scap > n_worlds) g->scap = n_worlds;
    }
    for (int i = 0; i < FizzGroup::NR; i++) { g->rstream[i] = nullptr; g->rdone[i] = nullptr; }
    for (int i = 0; i < 3; i++) { g->prof_ms[i] = 0.0; g->prof_launches[i] = 0; }
    KParams &kp = g->kp;
    double *s = g->slab;
    kp.pos = s + L.pos; kp.vel = s + L.vel; kp.heat = s + L.heat;
    kp.status = (long long *)(s + L.status); kp.steps = (long long *)(s + L.steps);
    kp.calls = (long long *)(s + L.calls); kp.vflag = (long long *)(s + L.vflag);
    kp.tier = (long long *)(s + L.tier);
    kp.off = s + L.off; kp.thr = s + L.thr; kp.mass = s + L.mass; kp.verts = s + L.verts;
    g->sched_list = (long long *)(s + L.sched);
    g->sched_list2 = (long long *)(s + L.sched + L.stride);
    g->sched_count = (unsigned long long *)(s + L.sched + 2 * L.stride);
    kp.counters = g->sched_count + 4;
    kp.W = L.n_worlds; kp.Wp = L.stride;
    kp.F = L.n_free; kp.X = L.n_fixed; kp.V = L.n_free_verts;
    const int F = kp.F, X = kp.X;
    int vfree = 0, vfix = 0;
    for (int b = 0; b < F + X; b++) {
        kp.nv[b] = d->nverts[b];
        if (b < F) { kp.vs[b] = vfree; vfree += d->nverts[b]; }
        else { kp.vs[b] = vfix; vfix += d->nverts[b]; }
    }
    kp.FV = vfix;
    {
        int max_na = 0;
        for (int i = 0; i < F; i++)
            for (int j = i + 1; j < F + X; j++) max_na = std::max(max_na, kp.nv[i] + kp.nv[j]);
        kp.swl = max_na <= 8 ? 3 : (max_na <= 16 ? 4 : 5);
    }
    double *fixed_dev = s + L.scene;
    kp.fixed_geom = fixed_dev;
    g->vbody_dev = (unsigned char *)(fixed_dev + round_up((long long)vfix * 6, 16));
    // fixed-scene constants, computed exactly as polypoly_collision would (:775-811) every call
    double *fg = new double[(size_t)(vfix > 0 ? vfix : 1) * 6];
    for (int j = 0; j < X; j++) {
        const int n = kp.nv[F + j], v0 = kp.vs[F + j];
        const double *xy = d->fixed_xy + 2 * v0;
        kp.fcx[j] = d->fixed_com[2 * j]; kp.fcy[j] = d->fixed_com[2 * j + 1]; kp.fthr[j] = d->fixed_thr[j];
        for (int k = 0; k < n; k++) {
            const int k2 = (k + 1 == n) ? 0 : k + 1;
            const double ax = xy[2 * k2 + 1] - xy[2 * k + 1];
            const double ay = xy[2 * k] - xy[2 * k2];
            const double nrm = std::sqrt(sq2(ax, ay));
            const double nx = ax / nrm, ny = ay / nrm;
            double lo, hi;
            lo = hi = dot2(nx, ny, xy[0], xy[1]);
            for (int v = 1; v < n; v++) {
                const double cur = dot2(nx, ny, xy[2 * v], xy[2 * v + 1]);
                if (cur < lo) lo = cur;
                if (cur > hi) hi = cur;
            }
            double *o6 = fg + (size_t)(v0 + k) * 6;
            o6[0] = xy[2 * k]; o6[1] = xy[2 * k + 1]; o6[2] = nx; o6[3] = ny; o6[4] = lo; o6[5] = hi;
        }
    }
    {
        double xlo = INFINITY, xhi = -INFINITY, ylo = INFINITY, yhi = -INFINITY, rmax = 0.0;
        for (int j = 0; j < X; j++) {
            xlo = std::min(xlo, kp.fcx[j]); xhi = std::max(xhi, kp.fcx[j]);
            ylo = std::min(ylo, kp.fcy[j]); yhi = std::max(yhi, kp.fcy[j]);
            const double r = std::sqrt(kp.fthr[j]);     // T(r) <= r^2 (1 + 2^-52): sqrt(T) is r to an ulp; NaN stays NaN
            rmax = (r > rmax || r != r) ? r : rmax;
        }
        kp.hbx = X ? .5 * (xlo + xhi) : 0.0; kp.hby = X ? .5 * (ylo + yhi) : 0.0;
        kp.hhx = X ? .5 * (xhi - xlo) : 0.0; kp.hhy = X ? .5 * (yhi - ylo) : 0.0;
        kp.frmax = rmax;
    }
    cudaError_t e = cudaMemcpy(fixed_dev, fg, sizeof(double) * (size_t)vfix * 6, cudaMemcpyHostToDevice);
    delete[] fg;
    if (e != cudaSuccess) { delete g; return cuda_fail(e, "cudaMemcpy(fixed scene)"); }
    unsigned char vb[FIZZ_MAX_FREE * FIZZ_MAX_POLY_VERTS];
    for (int b = 0; b < F; b++)
        for (int v = 0; v < kp.nv[b]; v++) vb[kp.vs[b] + v] = (unsigned char)b;
    e = cudaMemcpy(g->vbody_dev, vb, (size_t)kp.V, cudaMemcpyHostToDevice);
    if (e != cudaSuccess) { delete g; return cuda_fail(e, "cudaMemcpy(vertex table)"); }
    const FizzParams &p = d->params;
    kp.coll_tol = p.collision_tol; kp.time_tol = p.time_tol; kp.eps = p.epsilon; kp.vib_tol = p.vibrate_tol;
    kp.time_disc = p.time_disc; kp.ope = p.one_plus_e; kp.ome2 = p.one_minus_e2; kp.accx = p.acc_x; kp.accy = p.acc_y;
    kp.max_calls = p.max_calls;
    kp.mu = p.friction_mu; kp.ffr = p.ff_restitution;
    fizz_host_sqrt_threshold(1, &p.vibrate_tol, &kp.vib_thr);

    // thread mode (may not fit in shared memory for vertex-rich worlds: then only hybrid mode is available)
    g->thread_smem = sizeof(double) * ((size_t)kp.FV * 6 + (size_t)kp.V * 2 * BLOCK) + sizeof(int) * 64;
    FIZZ_DISPATCH_F(F, e = prep_thread<FF>(g->thread_smem, &g->thread_ki));
    g->thread_ok = (e == cudaSuccess);
    if (!g->thread_ok) (void)cudaGetLastError();
    // hybrid mode
    g->ballistic_smem = ballistic_smem_bytes(F, kp.V, kp.FV);
    FIZZ_DISPATCH_F(F, e = prep_ballistic<FF>(g->ballistic_smem, &g->ballistic_ki, &g->prover_ki, &g->prover_ok));
    if (e != cudaSuccess) { delete g; return cuda_fail(e, "ballistic kernel setup"); }
    g->contact_smem = contact_smem_bytes(kp.V, kp.FV);
    {
        static size_t current_general[MAX_DEVICES] = {}, current_trapped[MAX_DEVICES] = {};
        e = raise_dynamic_smem(contact_kernel<false>, g->contact_smem, current_general);
        // (runner launches ask for RUNNER_SMEM so that nothing else becomes resident on the SM that hosts them)
        if (e == cudaSuccess)
            e = raise_dynamic_smem(contact_kernel<true>, g->contact_smem > RUNNER_SMEM ? g->contact_smem : RUNNER_SMEM,
                                   current_trapped);
    }
    if (e == cudaSuccess) e = cudaFuncGetAttributes(&g->contact_ki.attr, contact_kernel<false>);
    if (e == cudaSuccess)
        e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&g->contact_ki.blocks_per_sm, contact_kernel<false>, CONTACT_BLOCK,
                                                          g->contact_smem);
    if (e == cudaSuccess) e = cudaFuncGetAttributes(&g->trapped_ki.attr, contact_kernel<true>);
    if (e == cudaSuccess)
        e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&g->trapped_ki.blocks_per_sm, contact_kernel<true>, CONTACT_BLOCK,
                                                          g->contact_smem);
    if (e != cudaSuccess) { delete g; return cuda_fail(e, "contact kernel setup"); }
    if (g->contact_ki.blocks_per_sm < 1 || g->trapped_ki.blocks_per_sm < 1) {
        delete g;
        return fail(FIZZ_E_ARG, "fizz_group_create: the scene does not fit in the shared memory of an SM (contact kernel: "
                                "too many free/fixed vertices)");
    }
    e = cudaStreamCreateWithFlags(&g->side, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&g->ev_fork, cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&g->ev_join, cudaEventDisableTiming);
    for (int i = 0; i < FizzGroup::NR && e == cudaSuccess; i++) {
        e = cudaStreamCreateWithFlags(&g->rstream[i], cudaStreamNonBlocking);
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&g->rdone[i], cudaEventDisableTiming);
    }
    g->rcap = (long long)(g->sm_count / 3 > 0 ? g->sm_count / 3 : 1) * CONTACT_WARPS;  // a third of the SMs at most
    if (e == cudaSuccess) e = cudaMalloc(&g->rlist, sizeof(long long) * (size_t)FizzGroup::NR * (size_t)g->rcap);
    if (e == cudaSuccess) e = cudaMalloc(&g->rcount, sizeof(unsigned long long) * 2 * FizzGroup::NR);
    if (e != cudaSuccess) { fizz_group_destroy(g); return cuda_fail(e, "side stream setup"); }
    *out = g;
    return FIZZ_OK;
}

int fizz_group_destroy(FizzGroup *g) {
    if (g) {
        for (cudaEvent_t e : g->ev_pool) cudaEventDestroy(e);
        if (g->ev_fork) cudaEventDestroy(g->ev_fork);
        if (g->ev_join) cudaEventDestroy(g->ev_join);
        if (g->side) cudaStreamDestroy(g->side);
        for (int i = 0; i < FizzGroup::NR; i++) {
            if (g->rdone[i]) cudaEventDestroy(g->rdone[i]);
            if (g->rstream[i]) cudaStreamDestroy(g->rstream[i]);
        }
        if (g->rlist) cudaFree(g->rlist);
        if (g->spec_list) cudaFree(g->spec_list);
        if (g->spec_slots) cudaFree(g->spec_slots);
        if (g->spec_ctrl) cudaFree(g->spec_ctrl);
        if (g->spec_state) cudaFree(g->spec_state);
        if (g->spec_queue) cudaFree(g->spec_queue);
        if (g->spec_fork) cudaEventDestroy(g->spec_fork);
        if (g->spec_join) cudaEventDestroy(g->spec_join);
        if (g->spec_stream) cudaStreamDestroy(g->spec_stream);
        if (g->rcount) cudaFree(g->rcount);
    }
    delete g;
    return FIZZ_OK;
}

int fizz_set_profiling(FizzGroup *g, int32_t enabled) {
    if (!g) return fail(FIZZ_E_ARG, "fizz_set_profiling: null group");
    g->profile = enabled != 0;
    return FIZZ_OK;
}

int fizz_profile_read(FizzGroup *g, double *ms, int64_t *launches, int32_t reset) {
    if (!g || !ms || !launches) return fail(FIZZ_E_ARG, "fizz_profile_read: null argument");
    CU(cudaSetDevice(g->device));
    for (const ProfSpan &sp : g->spans) {
        CU(cudaEventSynchronize(sp.b));
        float t = 0.f;
        CU(cudaEventElapsedTime(&t, sp.a, sp.b));
        g->prof_ms[sp.kind] += t;
        g->prof_launches[sp.kind] += 1;
    }
    g->spans.clear();
    g->ev_used = 0;
    for (int i = 0; i < 3; i++) { ms[i] = g->prof_ms[i]; launches[i] = g->prof_launches[i]; }
    if (reset)
        for (int i = 0; i < 3; i++) { g->prof_ms[i] = 0.0; g->prof_launches[i] = 0; }
    return FIZZ_OK;
}

int fizz_set_tuning(FizzGroup *g, int32_t mode, int64_t chunk_steps) {
    if (!g) return fail(FIZZ_E_ARG, "fizz_set_tuning: null group");
    if (mode != FIZZ_MODE_HYBRID && mode != FIZZ_MODE_THREAD && mode != FIZZ_MODE_HYBRID3 && mode != FIZZ_MODE_HYBRID_SYNC &&
        mode != FIZZ_MODE_THREAD_FP32)
        return fail(FIZZ_E_ARG, "fizz_set_tuning: unknown mode");
    if ((mode == FIZZ_MODE_THREAD || mode == FIZZ_MODE_THREAD_FP32) && !g->thread_ok)
        return fail(FIZZ_E_ARG, "fizz_set_tuning: thread mode does not fit in shared memory for this topology");
    if (mode == FIZZ_MODE_HYBRID3 && !g->prover_ok)
        return fail(FIZZ_E_ARG, "fizz_set_tuning: hybrid3 mode: the prover tier does not fit in shared memory for this topology");
    if (chunk_steps < 1) return fail(FIZZ_E_ARG, "fizz_set_tuning: chunk_steps must be >= 1");
    g->mode = mode;
    g->chunk = chunk_steps;
    return FIZZ_OK;
}

static int copy_rows(double *dst, const double *src, int64_t rows, int64_t W, int64_t Wp, cudaMemcpyKind kind,
                     cudaStream_t st) {
    if (rows == 0) return FIZZ_OK;
    if (kind == cudaMemcpyHostToDevice)
        CU(cudaMemcpy2DAsync(dst, sizeof(double) * Wp, src, sizeof(double) * W, sizeof(double) * W, rows, kind, st));
    else
        CU(cudaMemcpy2DAsync(dst, sizeof(double) * W, src, sizeof(double) * Wp, sizeof(double) * W, rows, kind, st));
    return FIZZ_OK;
}

int fizz_upload(FizzGroup *g, const double *pos, const double *vel, const double *off, const double *thr,
                const double *mass, void *stream) {
    if (!g || !pos || !vel || !off || !thr || !mass) return fail(FIZZ_E_ARG, "fizz_upload: null argument");
    cudaStream_t st = (cudaStream_t)stream;
    CU(cudaSetDevice(g->device));
    const FizzLayout &L = g->lay;
    const int64_t W = L.n_worlds, Wp = L.stride;
    const int F = L.n_free, V = L.n_free_verts;
    double *s = g->slab;
    // heat/status/steps/calls/vflag := 0 (also clears the padding columns of pos/vel so they stay finite)
    CU(cudaMemsetAsync(s + L.pos, 0, sizeof(double) * L.state_words, st));
    int rc;
    if ((rc = copy_rows(s + L.pos, pos, 2 * F, W, Wp, cudaMemcpyHostToDevice, st))) return rc;
    if ((rc = copy_rows(s + L.vel, vel, 2 * F, W, Wp, cudaMemcpyHostToDevice, st))) return rc;
    if ((rc = copy_rows(s + L.off, off, 2 * V, W, Wp, cudaMemcpyHostToDevice, st))) return rc;
    if ((rc = copy_rows(s + L.thr, thr, F, W, Wp, cudaMemcpyHostToDevice, st))) return rc;
    if ((rc = copy_rows(s + L.mass, mass, F, W, Wp, cudaMemcpyHostToDevice, st))) return rc;
    g->uploaded = true;
    g->target = 0;
    return FIZZ_OK;
}

int fizz_snapshot(FizzGroup *g, void *dev_dst, void *stream) {
    if (!g || !dev_dst) return fail(FIZZ_E_ARG, "fizz_snapshot: null argument");
    if (!g->uploaded) return fail(FIZZ_E_STATE, "fizz_snapshot: call fizz_upload first");
    CU(cudaSetDevice(g->device));
    CU(cudaMemcpyAsync(dev_dst, g->slab + g->lay.pos, sizeof(double) * g->lay.state_words, cudaMemcpyDeviceToDevice,
                       (cudaStream_t)stream));
    return FIZZ_OK;
}

int fizz_restore(FizzGroup *g, const void *dev_src, int64_t steps_done, void *stream) {
    if (!g || !dev_src || steps_done < 0) return fail(FIZZ_E_ARG, "fizz_restore: bad argument");
    if (!g->uploaded) return fail(FIZZ_E_STATE, "fizz_restore: call fizz_upload first");
    CU(cudaSetDevice(g->device));
    CU(cudaMemcpyAsync(g->slab + g->lay.pos, dev_src, sizeof(double) * g->lay.state_words, cudaMemcpyDeviceToDevice,
                       (cudaStream_t)stream));
    g->target = steps_done;
    return FIZZ_OK;
}

int fizz_step(FizzGroup *g, int64_t n_steps, void *stream) {
    if (!g) return fail(FIZZ_E_ARG, "fizz_step: null group");
    if (!g->uploaded) return fail(FIZZ_E_STATE, "fizz_step: call fizz_upload first");
    if (n_steps < 0) return fail(FIZZ_E_ARG, "fizz_step: n_steps < 0");
    if (n_steps == 0) return FIZZ_OK;
    cudaStream_t st = (cudaStream_t)stream;
    CU(cudaSetDevice(g->device));
    KParams kp = g->kp;
    const int F = kp.F;
    if (g->mode == FIZZ_MODE_THREAD || g->mode == FIZZ_MODE_THREAD_FP32) {
        g->target += n_steps;
        kp.target_steps = g->target;
        dim3 grid((unsigned)(g->lay.stride / BLOCK));
        cudaEvent_t ea = g->profile ? next_event(g, st) : nullptr;
        if (g->mode == FIZZ_MODE_THREAD) {
            FIZZ_DISPATCH_F(F, (step_kernel<FF, double><<<grid, BLOCK, g->thread_smem, st>>>(kp)));
        } else {
            FIZZ_DISPATCH_F(F, (step_kernel<FF, float><<<grid, BLOCK, g->thread_smem, st>>>(kp)));
        }
        if (g->profile) {
            cudaEvent_t eb = next_event(g, st);
            if (ea && eb) g->spans.push_back({FIZZ_KERNEL_THREAD, ea, eb});
            if (g->prof_error != cudaSuccess) return cuda_fail(g->prof_error, "fizz_step: timing event");
        }
        CU(cudaGetLastError());
        g->launches++;
        return FIZZ_OK;
    }
    const long long W = g->lay.n_worlds;
    const dim3 bgrid((unsigned)((W + BALLISTIC_BLOCK - 1) / BALLISTIC_BLOCK));
    const dim3 cgrid((unsigned)((W + 255) / 256));
    // persistent grid: one CTA slot per resident block, but never more warps than worlds
    const long long kneed = (W + CONTACT_WARPS - 1) / CONTACT_WARPS;
    const long long kfull = (long long)g->sm_count * g->contact_ki.blocks_per_sm;
    const long long tfull = (long long)g->sm_count * g->trapped_ki.blocks_per_sm;
    const dim3 kgrid((unsigned)(kneed < kfull ? kneed : kfull)), tgrid((unsigned)(kneed < tfull ? kneed : tfull));
    // runner mode (FIZZ_MODE_HYBRID): at a chunk boundary the worlds flagged "trapped" (tier 3) are handed to their own
    // launch of the latency-tuned instance, on a library-owned stream that is idle, with the END of this call as target;
    // they are marked tier 4 on the way, which keeps the chunk pipeline away from them.  Nothing waits for a runner
    // before the join below, and a runner waits for nothing.  Short calls (per-timestep stepping) do not bother.
    const long long final_target = g->target + n_steps;
#ifdef FIZZ_RUNNER_DEBUG
    cudaEvent_t d0 = nullptr, d_chunk[64] = {}, d_spec[64] = {}, d_ball[64] = {}, d_main[64] = {};
    static unsigned long long d_cnt[64][9];
    std::memset(d_cnt, 0, sizeof(d_cnt));
    static unsigned long long *d_sdbg = nullptr;
    long long d_round0[66] = {};
    if (!d_sdbg) cudaMalloc(&d_sdbg, sizeof(unsigned long long) * 4 * FizzGroup::SPEC_RMAX);
    cudaMemsetAsync(d_sdbg, 0, sizeof(unsigned long long) * 4 * FizzGroup::SPEC_RMAX, st);
    long long d_len[64] = {};
    cudaEventCreate(&d0);
    cudaEventRecord(d0, st);
#endif
    // spec rounds (FIZZ_MODE_HYBRID, calls of >= 128 timesteps): in every chunk after the first, the worlds whose
    // timesteps are expensive are brought to the chunk target time-parallel (fizz_common.cuh: SpecParams) on a library-owned
    // stream, beside the launch that steps everything else; whatever the rounds leave is finished by a launch after the
    // join.  Not with the contact trace on (its records would come from discarded slots too).  With spec rounds the
    // runner launches only take the expensive worlds the rounds cannot predict.
    const bool spec = (g->mode == FIZZ_MODE_HYBRID) && n_steps >= 128 && g->spec_on && kp.trace == nullptr;
    const bool runners = (g->mode == FIZZ_MODE_HYBRID) && n_steps >= 128;
    const SpecParams no_spec = {};
    SpecParams sp = {};
    long long spec_round = 0;
    if (spec) {
        if (!g->spec_slots) {
            const size_t slot_words = (size_t)g->scap * SPEC_K * (size_t)spec_slot_words(F);
            const size_t ctrl_words = (size_t)g->lay.n_worlds * SPEC_CTRL_WORDS;
            CU(cudaMalloc(&g->spec_list, sizeof(long long) * (size_t)g->scap));
            CU(cudaMalloc(&g->spec_slots, sizeof(double) * slot_words));
            CU(cudaMalloc(&g->spec_ctrl, sizeof(long long) * ctrl_words));
            CU(cudaMalloc(&g->spec_state, sizeof(unsigned long long) * 8));
            CU(cudaMalloc(&g->spec_queue, sizeof(unsigned long long) * (size_t)g->scap * SPEC_K));
            // (on the caller's stream: everything that touches these buffers is ordered after it through the fork events;
            //  a plain cudaMemset runs on the legacy default stream, which the non-blocking streams do not wait for)
            CU(cudaMemsetAsync(g->spec_state, 0, sizeof(unsigned long long) * 8, st));
            CU(cudaMemsetAsync(g->spec_queue, 0, sizeof(unsigned long long) * (size_t)g->scap * SPEC_K, st));
            CU(cudaMemsetAsync(g->spec_slots, 0, sizeof(double) * slot_words, st));
            CU(cudaMemsetAsync(g->spec_ctrl, 0, sizeof(long long) * ctrl_words, st));
            CU(cudaStreamCreateWithFlags(&g->spec_stream, cudaStreamNonBlocking));
            CU(cudaEventCreateWithFlags(&g->spec_fork, cudaEventDisableTiming));
            CU(cudaEventCreateWithFlags(&g->spec_join, cudaEventDisableTiming));
        }
        sp.slots = g->spec_slots; sp.ctrl = g->spec_ctrl;
        sp.S = g->spec_S; sp.D = g->spec_D; sp.budget_cost = g->spec_budget > 0 ? g->spec_budget : g->spec_D; sp.budget_steps = 32; sp.enabled = 1;
        { const char *v = std::getenv("FIZZ_SPEC_ABORT_X4"); sp.abort_x4 = (v && *v) ? std::atoi(v) : 12; if (sp.abort_x4 < 4) sp.abort_x4 = 4; }
        { const char *v = std::getenv("FIZZ_SPEC_NOPRED"); sp.nopred = (v && *v == '1') ? 1 : 0; }
        sp.queue = g->spec_queue; sp.state = g->spec_state; sp.qcap = (unsigned long long)g->scap * SPEC_K;
    }
    bool rused[FizzGroup::NR] = {};
    int rnext = 0;
    const unsigned long long runner_cap = (unsigned long long)g->rcap;  // all resident at once, see FizzGroup::rcap
    const dim3 rgrid((unsigned)(kneed < g->sm_count ? kneed : g->sm_count));
    const long long runner_rate = 48;   // check_collisions() in the world's last completed timestep
    const size_t runner_smem = g->contact_smem > RUNNER_SMEM ? g->contact_smem : RUNNER_SMEM;
    int64_t nchunk = 0;
    for (int64_t done = 0; done < n_steps; nchunk++) {
        // chunk length: the tuned value for the first chunks (contact episodes are short and early), then doubling every
        // second chunk up to 16x: a chunk boundary costs ~0.8 ms of launches and world reloads, and a world the
        // ballistic kernel hands over late in a long chunk only costs ~0.15 us per quiescent step in the contact kernel
        const int64_t len = g->chunk << ((nchunk / 2 < 4) ? nchunk / 2 : 4);
        const int64_t c = (n_steps - done < len) ? (n_steps - done) : len;
        g->target += c;
        kp.target_steps = g->target;
        cudaEvent_t ea = g->profile ? next_event(g, st) : nullptr;
        // tier 0 (quiescent worlds): lean registers-only variant; then tiers 0..1 (what is left + worlds with broad-phase
        // candidates): prover variant.  The tier is a performance hint only; every kernel re-derives what it needs.
        CU(cudaMemsetAsync(g->sched_count, 0, 2 * sizeof(unsigned long long), st));
        live_kernel<<<cgrid, 256, 0, st>>>(kp.status, kp.steps, kp.tier, W, g->target, 0, g->sched_list, g->sched_count);
        FIZZ_DISPATCH_F(F, (ballistic_kernel<FF, false><<<bgrid, BALLISTIC_BLOCK, 0, st>>>(kp, g->sched_list, g->sched_count)));
        if (g->mode == FIZZ_MODE_HYBRID3) {
            CU(cudaMemsetAsync(g->sched_count, 0, 2 * sizeof(unsigned long long), st));
            live_kernel<<<cgrid, 256, 0, st>>>(kp.status, kp.steps, kp.tier, W, g->target, 1, g->sched_list, g->sched_count);
            FIZZ_DISPATCH_F(F, (ballistic_kernel<FF, true><<<bgrid, BALLISTIC_BLOCK, g->ballistic_smem, st>>>(kp, g->sched_list, g->sched_count)));
            g->launches += 2;
        }
        cudaEvent_t eb = g->profile ? next_event(g, st) : nullptr;
#ifdef FIZZ_RUNNER_DEBUG
        if (nchunk < 64) { cudaEventCreate(&d_ball[nchunk]); cudaEventRecord(d_ball[nchunk], st); }
#endif
        CU(cudaMemsetAsync(g->sched_count, 0, 4 * sizeof(unsigned long long), st));
        if (g->mode == FIZZ_MODE_HYBRID || g->mode == FIZZ_MODE_HYBRID_SYNC) {
            // (nothing is flagged before the first chunk has run; a stream is used once per call: a second launch on it
            //  would wait for the first, which runs to the end of the call.  Every call ends with a join, so the streams
            //  are free again when the next call's launches reach the device -- no host-side query, the host may be
            //  many calls ahead of the GPU)
            int slot = -1;
            if (runners && g->target > c && rnext < FizzGroup::NR) slot = rnext++;
            if (slot >= 0) {
                unsigned long long *rc = g->rcount + 2 * slot;
                long long *rl = g->rlist + (size_t)slot * (size_t)g->rcap;
                CU(cudaMemsetAsync(rc, 0, 2 * sizeof(unsigned long long), st));
                // candidates: flagged trapped AND heavy right now (>= 48 check_collisions() in their last timestep: the
                // heaviest ~0.3 % of config 3) -- half of all worlds see one long resolve loop early on, and a runner
                // list longer than the grid would make the worlds at its end wait instead of run (hence also the cap).
                // A runner CTA asks for RUNNER_SMEM of shared memory, i.e. for an SM of its own: measured on config 3,
                // sharing SMs with the chunk pipeline makes the longest chain 0-20 % slower depending on where it lands
                // (540-640 ms per pass), alone it is 545-560 ms every time (no runners: 580-620)
                // (two passes so that the very heaviest get their slots first when the cap bites)
                if (spec) {
                    // spec mode: the runner launches take the expensive worlds the spec rounds cannot predict (bit 56 of the
                    // tier word: e.g. two free bodies stuck together far from the scene) -- inherently sequential chains,
                    // which should neither stop at chunk boundaries nor share an SM with anything
                    classify_kernel<<<cgrid, 256, 0, st>>>(kp.status, kp.steps, kp.tier, W, final_target, 0, 3, 2,
                                                           g->spec_cost, 1LL << 40, rl, rc, 4, runner_cap, 3);
                } else {
                classify_kernel<<<cgrid, 256, 0, st>>>(kp.status, kp.steps, kp.tier, W, final_target, 3, 3, 1,
                                                       3 * runner_rate, 1LL << 40, rl, rc, 4, runner_cap);
                classify_kernel<<<cgrid, 256, 0, st>>>(kp.status, kp.steps, kp.tier, W, final_target, 3, 3, 1,
                                                       runner_rate, 3 * runner_rate, rl, rc, 4, runner_cap);
                }
                CU(cudaEventRecord(g->ev_fork, st));
                CU(cudaStreamWaitEvent(g->rstream[slot], g->ev_fork, 0));
                KParams kr = kp;
                kr.target_steps = final_target;
                contact_kernel<true><<<rgrid, CONTACT_BLOCK, runner_smem, g->rstream[slot]>>>(kr, rl, rc, rc + 1, runner_cap, 1, no_spec);
                rused[slot] = true;
                g->launches += 2;
            }
            if (spec && g->target > c) {
                // ---- a chunk with spec rounds ----------------------------------------------------------------------------
                // (1) everything that lags; its length decides, ON THE DEVICE, which of the launches below do the work:
                //     a long list (>= BULK_LIST: the scene stays busy for the first few hundred timesteps) is a throughput
                //     phase -- the compact high-occupancy instance steps all of it, there are no rounds;
                unsigned long long *all_count = g->sched_count + 2;
                classify_kernel<<<cgrid, 256, 0, st>>>(kp.status, kp.steps, kp.tier, W, g->target, 0, 3, 0, 0, 1,
                                                       g->sched_list2, all_count);
                contact_kernel<false><<<kgrid, CONTACT_BLOCK, g->contact_smem, st>>>(kp, g->sched_list2, all_count,
                                                                                     g->sched_count + 3, ~0ull, 2, no_spec);
                // (2) otherwise: the worlds whose timesteps cost >= spec_cost and that the rounds can predict (list in world
                //     order, rebuilt every chunk) go through the rounds, on the library's spec stream; everything else that
                //     lags is stepped beside them by the latency-tuned instance.  Both lists are made BEFORE the rounds start:
                //     their commit steps rewrite the tier words the selection reads
                CU(cudaMemsetAsync(g->spec_state, 0, sizeof(unsigned long long), st));
                classify_kernel<<<cgrid, 256, 0, st>>>(kp.status, kp.steps, kp.tier, W, g->target, 0, 3, 2, g->spec_cost,
                                                       1LL << 40, g->spec_list, g->spec_state, -1, (unsigned long long)g->scap, 1,
                                                       all_count);
                classify_kernel<<<cgrid, 256, 0, st>>>(kp.status, kp.steps, kp.tier, W, g->target, 0, 3, 2, 0, g->spec_cost,
                                                       g->sched_list, g->sched_count, -1, ~0ull, 2, all_count);
                CU(cudaEventRecord(g->spec_fork, st));
                CU(cudaStreamWaitEvent(g->spec_stream, g->spec_fork, 0));
#ifdef FIZZ_RUNNER_DEBUG
                sp.round = (int)(nchunk < 64 ? nchunk : 63); sp.dbg = d_sdbg;
                if (nchunk < 64) d_round0[nchunk + 1] = nchunk + 1;
#endif
                spec_reset_kernel<<<1, 1, 0, g->spec_stream>>>(sp);
                spec_seed_kernel<<<(unsigned)((g->scap + 127) / 128), 128, 0, g->spec_stream>>>(kp, sp, g->spec_list,
                                                                                                 (unsigned long long)g->scap);
                // one launch per chunk: the warps take (world, slot) items from the rounds' queue until every listed world
                // is at the chunk target; it leaves room on every SM for the launch beside it
                const dim3 sgrid((unsigned)(g->sm_count * g->spec_ctas));
                contact_kernel<true><<<sgrid, CONTACT_BLOCK, g->contact_smem, g->spec_stream>>>(
                    kp, g->spec_list, g->spec_state, g->spec_state + 4, (unsigned long long)g->scap, 0, sp);
                CU(cudaEventRecord(g->spec_join, g->spec_stream));
#ifdef FIZZ_RUNNER_DEBUG
                if (nchunk < 64) { cudaEventCreate(&d_spec[nchunk]); cudaEventRecord(d_spec[nchunk], g->spec_stream); }
#endif
                spec_round++;
                contact_kernel<true><<<tgrid, CONTACT_BLOCK, g->contact_smem, st>>>(kp, g->sched_list, g->sched_count,
                                                                                    g->sched_count + 1, ~0ull, 0, no_spec);
#ifdef FIZZ_RUNNER_DEBUG
                if (nchunk < 64) { cudaEventCreate(&d_main[nchunk]); cudaEventRecord(d_main[nchunk], st); }
#endif
                // (3) join, then whatever is still short of the chunk target: worlds that turned out unpredictable during
                //     the rounds, worlds beyond the spec list's capacity
                CU(cudaStreamWaitEvent(st, g->spec_join, 0));
                CU(cudaMemsetAsync(all_count, 0, 2 * sizeof(unsigned long long), st));
                classify_kernel<<<cgrid, 256, 0, st>>>(kp.status, kp.steps, kp.tier, W, g->target, 0, 3, 0, 0, 1,
                                                       g->sched_list2, all_count);
                contact_kernel<true><<<tgrid, CONTACT_BLOCK, g->contact_smem, st>>>(kp, g->sched_list2, all_count,
                                                                                    g->sched_count + 3, ~0ull, 0, no_spec);
                g->launches += 11;
            } else {
                // latency-tuned: ONE instance (two narrow-phase items per lane, 3 CTAs/SM) steps every lagging world, so the
                // world with the longest chain of resolve passes never waits for anything
                // (work list in world order: measured -- grouping the trapped worlds at the head of the list, or sorting them
                //  by call rate, puts the longest chains on the same SMs and costs 2-5 % of a pass)
                classify_kernel<<<cgrid, 256, 0, st>>>(kp.status, kp.steps, kp.tier, W, g->target, 0, 3, 0, 0, 1,
                                                       g->sched_list, g->sched_count);   // (tier 4: owned by a runner)
                // the chunk that starts at step 0 has every world in the scene and none trapped yet: a pure throughput
                // phase, which the compact high-occupancy instance handles ~20 % faster; afterwards latency rules
                if (g->target == c)
                    contact_kernel<false><<<kgrid, CONTACT_BLOCK, g->contact_smem, st>>>(kp, g->sched_list, g->sched_count,
                                                                                         g->sched_count + 1, ~0ull, 0, no_spec);
                else
                    contact_kernel<true><<<tgrid, CONTACT_BLOCK, g->contact_smem, st>>>(kp, g->sched_list, g->sched_count,
                                                                                        g->sched_count + 1, ~0ull, 0, no_spec);
#ifdef FIZZ_RUNNER_DEBUG
                if (nchunk < 64) { cudaEventCreate(&d_main[nchunk]); cudaEventRecord(d_main[nchunk], st); }
#endif
            }
            g->launches -= 2;
        } else {
            // throughput-tuned (hybrid3): two instances run SIDE BY SIDE on disjoint worlds -- the trapped instance (tier
            // 3 at the start of the chunk) on the library's side stream, the compact high-occupancy general instance
            // (everything else that lags) on the caller's stream.  Both bring their worlds to the chunk target; the
            // general instance only FLAGS the worlds it finds trapped, the hand-over happens at the next chunk.
            classify_kernel<<<cgrid, 256, 0, st>>>(kp.status, kp.steps, kp.tier, W, g->target, 3, 3, 0, 0, 1, g->sched_list2,
                                                   g->sched_count + 2);
            classify_kernel<<<cgrid, 256, 0, st>>>(kp.status, kp.steps, kp.tier, W, g->target, 0, 2, 0, 0, 1, g->sched_list,
                                                   g->sched_count);
            CU(cudaEventRecord(g->ev_fork, st));
            CU(cudaStreamWaitEvent(g->side, g->ev_fork, 0));
            contact_kernel<true><<<tgrid, CONTACT_BLOCK, g->contact_smem, g->side>>>(kp, g->sched_list2, g->sched_count + 2,
                                                                                     g->sched_count + 3, ~0ull, 0, no_spec);
            CU(cudaEventRecord(g->ev_join, g->side));
            contact_kernel<false><<<kgrid, CONTACT_BLOCK, g->contact_smem, st>>>(kp, g->sched_list, g->sched_count,
                                                                                 g->sched_count + 1, ~0ull, 0, no_spec);
            CU(cudaStreamWaitEvent(st, g->ev_join, 0));
        }
        if (g->profile) {
            cudaEvent_t ec = next_event(g, st);
            if (ea && eb) g->spans.push_back({FIZZ_KERNEL_BALLISTIC, ea, eb});
            if (eb && ec) g->spans.push_back({FIZZ_KERNEL_CONTACT, eb, ec});
            if (g->prof_error != cudaSuccess) return cuda_fail(g->prof_error, "fizz_step: timing event");
        }
        CU(cudaGetLastError());
        g->launches += 6;
        done += c;
#ifdef FIZZ_RUNNER_DEBUG
        if (nchunk < 64) {
            cudaEventCreate(&d_chunk[nchunk]); cudaEventRecord(d_chunk[nchunk], st); d_len[nchunk] = c;
            cudaMemcpyAsync(d_cnt[nchunk], kp.counters, 4 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, st);
            cudaMemcpyAsync(&d_cnt[nchunk][4], g->sched_count, 4 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, st);
            if (spec) cudaMemcpyAsync(&d_cnt[nchunk][8], g->spec_state, sizeof(unsigned long long), cudaMemcpyDeviceToHost, st);
        }
#endif
    }
    bool any = false;
#ifdef FIZZ_RUNNER_DEBUG
    cudaEvent_t d_mainend = nullptr, d_run[FizzGroup::NR] = {};
    cudaEventCreate(&d_mainend);
    cudaEventRecord(d_mainend, st);
    for (int i = 0; i < FizzGroup::NR; i++)
        if (rused[i]) { cudaEventCreate(&d_run[i]); cudaEventRecord(d_run[i], g->rstream[i]); }
#endif
    for (int i = 0; i < FizzGroup::NR; i++)
        if (rused[i]) {
            CU(cudaEventRecord(g->rdone[i], g->rstream[i]));
            CU(cudaStreamWaitEvent(st, g->rdone[i], 0));
            any = true;
        }
    if (any) {
        // the runners' worlds go back to the pipeline (the only place tier 4 is released: see the contact kernel)
        release_kernel<<<cgrid, 256, 0, st>>>(kp.tier, W);
        // whatever is still short of the target (a world beyond a runner list's cap)
        CU(cudaMemsetAsync(g->sched_count, 0, 2 * sizeof(unsigned long long), st));
        classify_kernel<<<cgrid, 256, 0, st>>>(kp.status, kp.steps, kp.tier, W, final_target, 0, 4, 0, 0, 1,
                                               g->sched_list, g->sched_count);
        kp.target_steps = final_target;
        contact_kernel<true><<<tgrid, CONTACT_BLOCK, g->contact_smem, st>>>(kp, g->sched_list, g->sched_count,
                                                                            g->sched_count + 1, ~0ull, 0, no_spec);
        CU(cudaGetLastError());
        g->launches += 2;
    }
#ifdef FIZZ_RUNNER_DEBUG
    {
        cudaStreamSynchronize(st);
        float t = 0;
        cudaEventElapsedTime(&t, d0, d_mainend);
        fprintf(stderr, "[runner debug] main pipeline done at %.1f ms;", t);
        for (int i = 0; i < FizzGroup::NR; i++)
            if (d_run[i]) { cudaEventElapsedTime(&t, d0, d_run[i]); fprintf(stderr, " runner%d %.1f", i, t); cudaEventDestroy(d_run[i]); }
        unsigned long long rc[2 * FizzGroup::NR];
        cudaMemcpy(rc, g->rcount, sizeof(rc), cudaMemcpyDeviceToHost);
        fprintf(stderr, " | chunks (len@ms):");
        for (int i = 0; i < 64; i++)
            if (d_chunk[i]) { cudaEventElapsedTime(&t, d0, d_chunk[i]); fprintf(stderr, " %lld@%.1f", d_len[i], t); }
        fprintf(stderr, "\n  per chunk: [ballistic done, main list done, chunk done] ms | cumulative: ballistic world-steps, contact world-steps, calls, fast calls | main list, straggler list, spec list");
        for (int i = 0; i < 64; i++)
            if (d_chunk[i]) {
                float tb = 0, tm = 0;
                cudaEventElapsedTime(&t, d0, d_chunk[i]);
                if (d_ball[i]) { cudaEventElapsedTime(&tb, d0, d_ball[i]); cudaEventDestroy(d_ball[i]); }
                if (d_main[i]) { cudaEventElapsedTime(&tm, d0, d_main[i]); cudaEventDestroy(d_main[i]); }
                fprintf(stderr, "\n   chunk %2d: [%.1f %.1f %.1f] | %llu %llu %llu %llu | %llu %llu %llu", i, tb, tm, t, d_cnt[i][0], d_cnt[i][1],
                        d_cnt[i][2], d_cnt[i][3], d_cnt[i][4], d_cnt[i][6], d_cnt[i][8]);
                cudaEventDestroy(d_chunk[i]);
            }
        fprintf(stderr, " | list sizes:");
        for (int i = 0; i < FizzGroup::NR; i++) fprintf(stderr, " %llu", rc[2 * i]);
        if (spec) {
            fprintf(stderr, "\n  spec rounds end at (ms):");
            for (int i = 0; i < 64; i++)
                if (d_spec[i]) { cudaEventElapsedTime(&t, d0, d_spec[i]); fprintf(stderr, " %.1f", t); cudaEventDestroy(d_spec[i]); }
            std::vector<long long> c((size_t)W * SPEC_CTRL_WORDS);
            cudaMemcpy(c.data(), g->spec_ctrl, sizeof(long long) * c.size(), cudaMemcpyDeviceToHost);
            long long n = 0, acc = 0, ev = 0, miss = 0, blind = 0, rounds = 0;
            for (long long w = 0; w < W; w++) {
                const long long *x = c.data() + w * SPEC_CTRL_WORDS;
                if (x[9] == 0) continue;
                n++; acc += x[5]; ev += x[6]; miss += x[7]; blind += x[8]; rounds += x[9];
            }
            {
                std::vector<unsigned long long> sd(4 * FizzGroup::SPEC_RMAX);
                cudaMemcpy(sd.data(), d_sdbg, sizeof(unsigned long long) * sd.size(), cudaMemcpyDeviceToHost);
                fprintf(stderr, "\n  spec rounds per chunk: rounds with work / launched | longest item mean, max (us at 1.965 GHz) | item time total (ms) | items, given up");
                long long prev = 0;
                for (int i = 1; i < 65; i++) {
                    if (!d_round0[i]) continue;
                    long long used = 0, items = 0, ab = 0;
                    double sum_max = 0, mx = 0, tot = 0;
                    for (long long r = prev; r < d_round0[i]; r++) {
                        if (sd[4 * r + 2] == 0) continue;
                        used++; sum_max += sd[4 * r] / 1965.0; mx = std::max(mx, sd[4 * r] / 1965.0); tot += sd[4 * r + 1] / 1965.0e3;
                        items += sd[4 * r + 2]; ab += sd[4 * r + 3];
                    }
                    fprintf(stderr, "\n   chunk %2d: %lld / %lld | %.0f %.0f | %.1f | %lld %lld", i - 1, used, d_round0[i] - prev, used ? sum_max / used : 0.0, mx,
                            tot, items, ab);
                    prev = d_round0[i];
                }
            }
            {
                std::vector<long long> tw((size_t)W);
                cudaMemcpy(tw.data(), kp.tier, sizeof(long long) * (size_t)W, cudaMemcpyDeviceToHost);
                std::vector<long long> idx;
                for (long long w = 0; w < W; w++) if (c[w * SPEC_CTRL_WORDS + 9]) idx.push_back(w);
                std::sort(idx.begin(), idx.end(), [&](long long a, long long b) {
                    return c[a * SPEC_CTRL_WORDS + 9] > c[b * SPEC_CTRL_WORDS + 9]; });
                fprintf(stderr, "\n  worst worlds (rounds - accepted/32): world rounds accepted evaluated mispredicted blind | calls cost tier | classes");
                for (size_t i = 0; i < idx.size() && i < 16; i++) {
                    const long long *x = c.data() + idx[i] * SPEC_CTRL_WORDS;
                    fprintf(stderr, "\n   %lld: (window estimate %lld) %lld %lld %lld %lld %lld | %lld %lld %lld | %016llx %016llx", idx[i], x[10] / 16, x[9], x[5], x[6], x[7], x[8],
                            (tw[idx[i]] >> 8) & 0xffffff, (tw[idx[i]] >> 32) & 0xffffff, tw[idx[i]] & 255, (unsigned long long)x[1], (unsigned long long)x[2]);
                }
            }
            {
                std::vector<unsigned long long> cyc(65536);
                cudaMemcpyFromSymbol(cyc.data(), g_dbg_cycles, sizeof(unsigned long long) * 65536);
                std::vector<long long> tw((size_t)W);
                cudaMemcpy(tw.data(), kp.tier, sizeof(long long) * (size_t)W, cudaMemcpyDeviceToHost);
                std::vector<int> idx(std::min<long long>(W, 65536));
                for (size_t i = 0; i < idx.size(); i++) idx[i] = (int)i;
                std::sort(idx.begin(), idx.end(), [&](int a, int b) { return cyc[a] > cyc[b]; });
                fprintf(stderr, "\n  ordinary-launch time from timestep 2688 on (cumulative over calls): world ms | calls cost tier flag | spec rounds accepted");
                for (int i = 0; i < 12 && i < (int)idx.size(); i++) {
                    const int w = idx[i];
                    fprintf(stderr, "\n   %d: %.1f | %lld %lld %lld %lld | %lld %lld", w, cyc[w] / 1.965e6, (tw[w] >> 8) & 0xffffff, (tw[w] >> 32) & 0xffffff,
                            tw[w] & 255, (tw[w] >> 56) & 1, c[(size_t)w * SPEC_CTRL_WORDS + 9], c[(size_t)w * SPEC_CTRL_WORDS + 5]);
                }
            }
            fprintf(stderr, "\n  spec (cumulative): %lld worlds, %lld launches this call; timesteps accepted %lld of %lld evaluated; world-rounds %lld, "
                    "%lld mispredicted, %lld without a predictor", n, spec_round, acc, ev, rounds, miss, blind);
        }
        fprintf(stderr, "\n");
        cudaEventDestroy(d_mainend); cudaEventDestroy(d0);
    }
#endif
    return FIZZ_OK;
}

int fizz_energy(FizzGroup *g, void *stream) {
    if (!g) return fail(FIZZ_E_ARG, "fizz_energy: null group");
    if (!g->uploaded) return fail(FIZZ_E_STATE, "fizz_energy: call fizz_upload first");
    cudaStream_t st = (cudaStream_t)stream;
    CU(cudaSetDevice(g->device));
    const FizzLayout &L = g->lay;
    double *s = g->slab;
    const int threads = 256;
    const unsigned blocks = (unsigned)((L.n_worlds + threads - 1) / threads);
    energy_kernel<<<blocks, threads, 0, st>>>(s + L.pos, s + L.vel, s + L.heat, s + L.mass, s + L.energy, L.n_worlds,
                                              L.stride, L.n_free, g->prm.gy, g->prm.height);
    CU(cudaGetLastError());
    g->launches++;
    return FIZZ_OK;
}

int fizz_step_logged(FizzGroup *g, int64_t n_steps, double *dev_energy_log, void *stream) {
    if (!g || !dev_energy_log) return fail(FIZZ_E_ARG, "fizz_step_logged: null argument");
    if (!g->uploaded) return fail(FIZZ_E_STATE, "fizz_step_logged: call fizz_upload first");
    if (n_steps < 0) return fail(FIZZ_E_ARG, "fizz_step_logged: n_steps < 0");
    cudaStream_t st = (cudaStream_t)stream;
    CU(cudaSetDevice(g->device));
    const FizzLayout &L = g->lay;
    double *s = g->slab;
    const int threads = 256;
    const unsigned blocks = (unsigned)((L.n_worlds + threads - 1) / threads);
    const long long saved_chunk = g->chunk;
    g->chunk = 1;
    int rc = FIZZ_OK;
    for (int64_t t = 0; t < n_steps && rc == FIZZ_OK; t++) {
        // the row main.py would append to energy.txt BEFORE update() number t (physics.py:81-82), for every world
        energy_kernel<<<blocks, threads, 0, st>>>(s + L.pos, s + L.vel, s + L.heat, s + L.mass,
                                                  dev_energy_log + (size_t)t * 4 * L.stride, L.n_worlds, L.stride, L.n_free,
                                                  g->prm.gy, g->prm.height);
        g->launches++;
        rc = fizz_step(g, 1, stream);
    }
    g->chunk = saved_chunk;
    if (rc != FIZZ_OK) return rc;
    CU(cudaGetLastError());
    return FIZZ_OK;
}

int fizz_vertices(FizzGroup *g, void *stream) {
    if (!g) return fail(FIZZ_E_ARG, "fizz_vertices: null group");
    if (!g->uploaded) return fail(FIZZ_E_STATE, "fizz_vertices: call fizz_upload first");
    cudaStream_t st = (cudaStream_t)stream;
    CU(cudaSetDevice(g->device));
    const FizzLayout &L = g->lay;
    double *s = g->slab;
    const int threads = 256;
    const unsigned blocks = (unsigned)((L.n_worlds + threads - 1) / threads);
    derive_verts_kernel<<<blocks, threads, 0, st>>>(s + L.pos, s + L.off, (const long long *)(s + L.vflag), s + L.verts,
                                                    L.n_worlds, L.stride, L.n_free_verts, g->vbody_dev);
    CU(cudaGetLastError());
    g->launches++;
    return FIZZ_OK;
}

int fizz_download(FizzGroup *g, double *pos, double *vel, double *heat, double *energy, double *verts,
                  int64_t *status, int64_t *steps, int64_t *calls, void *stream) {
    if (!g) return fail(FIZZ_E_ARG, "fizz_download: null group");
    if (!g->uploaded) return fail(FIZZ_E_STATE, "fizz_download: call fizz_upload first");
    cudaStream_t st = (cudaStream_t)stream;
    CU(cudaSetDevice(g->device));
    const FizzLayout &L = g->lay;
    const int64_t W = L.n_worlds, Wp = L.stride;
    const int F = L.n_free, V = L.n_free_verts;
    double *s = g->slab;
    int rc;
    if (energy && (rc = fizz_energy(g, stream))) return rc;
    if (verts && (rc = fizz_vertices(g, stream))) return rc;
    if (pos && (rc = copy_rows(pos, s + L.pos, 2 * F, W, Wp, cudaMemcpyDeviceToHost, st))) return rc;
    if (vel && (rc = copy_rows(vel, s + L.vel, 2 * F, W, Wp, cudaMemcpyDeviceToHost, st))) return rc;
    if (heat && (rc = copy_rows(heat, s + L.heat, 1, W, Wp, cudaMemcpyDeviceToHost, st))) return rc;
    if (energy && (rc = copy_rows(energy, s + L.energy, 4, W, Wp, cudaMemcpyDeviceToHost, st))) return rc;
    if (verts && (rc = copy_rows(verts, s + L.verts, 2 * V, W, Wp, cudaMemcpyDeviceToHost, st))) return rc;
    if (status && (rc = copy_rows((double *)status, s + L.status, 1, W, Wp, cudaMemcpyDeviceToHost, st))) return rc;
    if (steps && (rc = copy_rows((double *)steps, s + L.steps, 1, W, Wp, cudaMemcpyDeviceToHost, st))) return rc;
    if (calls && (rc = copy_rows((double *)calls, s + L.calls, 1, W, Wp, cudaMemcpyDeviceToHost, st))) return rc;
    return FIZZ_OK;
}

int fizz_counters(FizzGroup *g, int64_t *out4, int32_t reset, void *stream) {
    if (!g || !out4) return fail(FIZZ_E_ARG, "fizz_counters: null argument");
    cudaStream_t st = (cudaStream_t)stream;
    CU(cudaSetDevice(g->device));
    CU(cudaMemcpyAsync(out4, g->kp.counters, 4 * sizeof(int64_t), cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    if (reset) CU(cudaMemsetAsync(g->kp.counters, 0, 4 * sizeof(int64_t), st));
    return FIZZ_OK;
}

int fizz_set_spec(FizzGroup *g, int32_t enabled, int32_t window0, int32_t slot_cost, int32_t min_cost) {
    if (!g) return fail(FIZZ_E_ARG, "fizz_set_spec: null group");
    if (window0 > 512) return fail(FIZZ_E_ARG, "fizz_set_spec: window0 must be <= 512");
    if (enabled >= 0) g->spec_on = enabled != 0;
    if (window0 > 0) g->spec_S = window0 < 2 ? 2 : window0;
    if (slot_cost > 0) g->spec_D = slot_cost;
    if (min_cost > 0) g->spec_cost = min_cost;
    return FIZZ_OK;
}

int fizz_spec_stats(FizzGroup *g, int64_t *out4, int32_t reset, void *stream) {
    if (!g || !out4) return fail(FIZZ_E_ARG, "fizz_spec_stats: null argument");
    cudaStream_t st = (cudaStream_t)stream;
    CU(cudaSetDevice(g->device));
    CU(cudaMemcpyAsync(out4, g->kp.counters + 4, 4 * sizeof(int64_t), cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    if (reset) CU(cudaMemsetAsync(g->kp.counters + 4, 0, 4 * sizeof(int64_t), st));
    return FIZZ_OK;
}

int fizz_set_trace(FizzGroup *g, void *dev_records, int64_t capacity, void *dev_counter) {
    if (!g) return fail(FIZZ_E_ARG, "fizz_set_trace: null group");
    if (dev_records && (!dev_counter || capacity < 1)) return fail(FIZZ_E_ARG, "fizz_set_trace: need counter and capacity");
    g->kp.trace = (FizzContact *)dev_records;
    g->kp.trace_count = (unsigned long long *)dev_counter;
    g->kp.trace_cap = dev_records ? capacity : 0;
    return FIZZ_OK;
}

int fizz_set_world_stats(FizzGroup *g, void *dev_general_calls) {
    if (!g) return fail(FIZZ_E_ARG, "fizz_set_world_stats: null group");
    g->kp.general_calls = (long long *)dev_general_calls;
    return FIZZ_OK;
}

int fizz_kernel_info(FizzGroup *g, int32_t which, int32_t *regs, int32_t *smem, int32_t *threads, int32_t *blocks_per_sm,
                     int64_t *launches) {
    if (!g) return fail(FIZZ_E_ARG, "fizz_kernel_info: null group");
    const KernelInfo *ki;
    int th;
    size_t sm;
    switch (which) {
        case FIZZ_KERNEL_THREAD: ki = &g->thread_ki; th = BLOCK; sm = g->thread_smem; break;
        case FIZZ_KERNEL_BALLISTIC: ki = &g->ballistic_ki; th = BALLISTIC_BLOCK; sm = 0; break;
        case FIZZ_KERNEL_PROVER: ki = &g->prover_ki; th = BALLISTIC_BLOCK; sm = g->ballistic_smem; break;
        case FIZZ_KERNEL_CONTACT: ki = &g->contact_ki; th = CONTACT_BLOCK; sm = g->contact_smem; break;
        case FIZZ_KERNEL_TRAPPED: ki = &g->trapped_ki; th = CONTACT_BLOCK; sm = g->contact_smem; break;
        default: return fail(FIZZ_E_ARG, "fizz_kernel_info: unknown kernel id");
    }
    if (regs) *regs = ki->attr.numRegs;
    if (smem) *smem = (int32_t)sm;
    if (threads) *threads = th;
    if (blocks_per_sm) *blocks_per_sm = ki->blocks_per_sm;
    if (launches) *launches = g->launches;
    return FIZZ_OK;
}

int fizz_fp64_peak(int device, int32_t iters, double *tflops, double *ms_out) {
    if (!tflops || iters < 1) return fail(FIZZ_E_ARG, "fizz_fp64_peak: bad argument");
    CU(cudaSetDevice(device));
    cudaDeviceProp prop;
    CU(cudaGetDeviceProperties(&prop, device));
    const int threads = 256, blocks = prop.multiProcessorCount * 8;
    double *out = nullptr;
    CU(cudaMalloc(&out, sizeof(double) * (size_t)threads * blocks));
    cudaEvent_t e0, e1;
    CU(cudaEventCreate(&e0));
    CU(cudaEventCreate(&e1));
    fp64_peak_kernel<<<blocks, threads>>>(out, iters / 8 + 1, 1.0000001, 1e-9);  // warm-up
    CU(cudaEventRecord(e0));
    fp64_peak_kernel<<<blocks, threads>>>(out, iters, 1.0000001, 1e-9);
    CU(cudaEventRecord(e1));
    CU(cudaEventSynchronize(e1));
    float ms = 0.f;
    CU(cudaEventElapsedTime(&ms, e0, e1));
    CU(cudaGetLastError());
    const double flops = 2.0 * 64.0 * (double)iters * threads * blocks;
    *tflops = flops / (ms * 1e-3) / 1e12;
    if (ms_out) *ms_out = ms;
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(out);
    return FIZZ_OK;
}

int fizz_raster_frames(int device, int32_t width, int32_t height, int64_t n_frames, int32_t n_polygons,
                       const int32_t *nverts, const int32_t *points_dev, uint8_t *gray_dev, void *stream) {
    if (width < 1 || height < 1 || n_frames < 0 || n_polygons < 0 || n_polygons > FIZZ_RASTER_MAX_POLYGONS)
        return fail(FIZZ_E_ARG, "fizz_raster_frames: bad size (1 <= width, height; 0 <= polygons <= 64)");
    if (n_frames == 0) return FIZZ_OK;
    if ((n_polygons && (!nverts || !points_dev)) || !gray_dev) return fail(FIZZ_E_ARG, "fizz_raster_frames: null argument");
    if (n_frames > 0x7fffffffLL) return fail(FIZZ_E_ARG, "fizz_raster_frames: more than 2^31-1 frames in one call");
    RasterParams rp;
    rp.W = width; rp.H = height; rp.RW = (width + 31) / 32;
    rp.n_poly = n_polygons; rp.n_points = 0;
    for (int p = 0; p < n_polygons; p++) {
        if (nverts[p] < 1) return fail(FIZZ_E_ARG, "fizz_raster_frames: polygon without vertices");
        rp.nverts[p] = nverts[p];
        rp.n_points += nverts[p];
    }
    rp.points = points_dev; rp.gray = gray_dev;
    if (rp.RW > 32) return fail(FIZZ_E_ARG, "fizz_raster_frames: canvas wider than 1024 pixels (a row is one warp of words)");
    const size_t smem = 2 * sizeof(unsigned) * (size_t)height * rp.RW;
    if (smem > 227 * 1024) return fail(FIZZ_E_ARG, "fizz_raster_frames: frame too large for the shared memory of an SM");
    CU(cudaSetDevice(device));
    static size_t raised[MAX_DEVICES] = {};
    CU(raise_dynamic_smem(raster_kernel, smem, raised));
    raster_kernel<<<(unsigned)n_frames, RASTER_BLOCK, smem, (cudaStream_t)stream>>>(rp);
    CU(cudaGetLastError());
    return FIZZ_OK;
}

int fizz_points_from_vertices(int device, int32_t width, int32_t height, int64_t n_frames, int32_t n_free_points,
                              const double *verts_dev, int32_t n_fixed_points, const int32_t *fixed_points_dev,
                              int32_t *points_dev, void *stream) {
    if (width < 1 || height < 1 || n_frames < 0 || n_free_points < 0 || n_fixed_points < 0)
        return fail(FIZZ_E_ARG, "fizz_points_from_vertices: bad size");
    const long long total = (long long)n_frames * (n_free_points + n_fixed_points);
    if (total == 0) return FIZZ_OK;
    if ((n_free_points && !verts_dev) || (n_fixed_points && !fixed_points_dev) || !points_dev)
        return fail(FIZZ_E_ARG, "fizz_points_from_vertices: null argument");
    CU(cudaSetDevice(device));
    points_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
        verts_dev, fixed_points_dev, points_dev, n_frames, n_free_points, n_fixed_points, width, height);
    CU(cudaGetLastError());
    return FIZZ_OK;
}

}  // extern "C"
