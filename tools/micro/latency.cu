// Dependent-issue latency of the instructions on the single-contact fast path (one warp, nothing else on the SM):
// the resolve chain of a trapped world is a serial dependency chain, so these cycles -- not throughput -- set its speed.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -fmad=false -o latency latency.cu && ./latency
#include <cstdio>
#include <cuda_runtime.h>

#define REP 256
#define CHAIN(NAME, DECL, BODY, SINK)                                                  \
    __global__ void NAME(long long *out, double a, double b, unsigned u, int iters) {  \
        const int lane = threadIdx.x & 31;                                             \
        (void)lane;                                                                    \
        DECL;                                                                          \
        long long t0 = clock64();                                                      \
        for (int it = 0; it < iters; it++) {                                           \
            _Pragma("unroll") for (int r = 0; r < REP; r++) { BODY; }                  \
        }                                                                              \
        long long t1 = clock64();                                                      \
        if (threadIdx.x == 0) { out[0] = t1 - t0; out[1] = (long long)(SINK); }        \
    }

CHAIN(k_dadd, double x = a, asm volatile("add.f64 %0, %0, %1;" : "+d"(x) : "d"(b)), x)
CHAIN(k_dmul, double x = a, asm volatile("mul.f64 %0, %0, %1;" : "+d"(x) : "d"(b)), x)
CHAIN(k_dfma, double x = a, asm volatile("fma.rn.f64 %0, %0, %1, %1;" : "+d"(x) : "d"(b)), x)
// min via compare + select (DSETP + 2 FSEL), the building block of the projection min/max
CHAIN(k_dmin, double x = a + lane; double y = b + lane,
      asm volatile("{.reg .pred p; setp.lt.f64 p, %1, %0; selp.f64 %0, %1, %0, p;}" : "+d"(x) : "d"(y)), x)
// DSETP -> predicate -> 32-bit SEL
CHAIN(k_dsetp_sel, double x = a; unsigned v = u,
      asm volatile("{.reg .pred p; .reg .f64 t; cvt.rn.f64.u32 t, %0; setp.lt.f64 p, t, %1; selp.u32 %0, %0, 7, p;}"
                   : "+r"(v) : "d"(x)), v)
CHAIN(k_iadd, unsigned v = u + lane, asm volatile("add.u32 %0, %0, %1;" : "+r"(v) : "r"(lane)), v)
CHAIN(k_lop, unsigned v = u + lane, asm volatile("xor.b32 %0, %0, %1;" : "+r"(v) : "r"(lane * 3 + 1)), v)
CHAIN(k_isetp_sel, unsigned v = u + lane,
      asm volatile("{.reg .pred p; setp.ne.u32 p, %0, %1; selp.u32 %0, %0, %2, p;}" : "+r"(v) : "r"(lane), "r"(lane + 9)), v)
// ballot result -> lane-dependent predicate -> ballot: IADD + ISETP + VOTE per link
CHAIN(k_vote, unsigned v = u,
      asm volatile("{.reg .pred p; .reg .u32 t; add.u32 t, %0, %1; setp.gt.u32 p, t, 12345; "
                   "vote.sync.ballot.b32 %0, p, 0xffffffff;}" : "+r"(v) : "r"(lane)), v)
// CREDUX + (uniform -> vector move) + IADD per link
CHAIN(k_redux, unsigned v = u + lane,
      asm volatile("{redux.sync.min.u32 %0, %0, 0xffffffff; add.u32 %0, %0, %1;}" : "+r"(v) : "r"(lane)), v)
// the two-stage pattern of the fast path: CREDUX -> move -> ISETP -> SEL per link
CHAIN(k_redux_cmp, unsigned v = u + lane,
      asm volatile("{.reg .u32 m; .reg .pred p; redux.sync.min.u32 m, %0, 0xffffffff; setp.eq.u32 p, m, %0; "
                   "selp.u32 %0, %1, %0, p;}" : "+r"(v) : "r"(lane + 17)), v)
CHAIN(k_shfl, unsigned v = u + lane,
      asm volatile("shfl.sync.idx.b32 %0, %0, %0, 0x1f, 0xffffffff;" : "+r"(v)), v)
// BREV + FLO + IADD(lane) per link
CHAIN(k_ffs, unsigned v = u | 1u,
      asm volatile("{.reg .u32 t; brev.b32 t, %0; bfind.shiftamt.u32 t, t; add.u32 %0, t, %1;}" : "+r"(v) : "r"(lane + 1)), v)
CHAIN(k_popc, unsigned v = u + lane, asm volatile("{popc.b32 %0, %0; add.u32 %0, %0, %1;}" : "+r"(v) : "r"(lane + 256)), v)
// a never-taken, data-dependent forward branch per link
CHAIN(k_branch, unsigned v = u + lane,
      if (v == 0xdeadbeefu) { asm volatile("trap;"); } asm volatile("add.u32 %0, %0, %1;" : "+r"(v) : "r"(lane)), v)
// issue rate of INDEPENDENT instructions from one warp (4 chains)
CHAIN(k_dadd4, double x0 = a; double x1 = a + 1; double x2 = a + 2; double x3 = a + 3,
      asm volatile("add.f64 %0, %0, %4; add.f64 %1, %1, %4; add.f64 %2, %2, %4; add.f64 %3, %3, %4;"
                   : "+d"(x0), "+d"(x1), "+d"(x2), "+d"(x3) : "d"(b)), x0 + x1 + x2 + x3)
CHAIN(k_iadd4, unsigned x0 = u; unsigned x1 = u + 1; unsigned x2 = u + 2; unsigned x3 = u + 3,
      asm volatile("add.u32 %0, %0, %4; add.u32 %1, %1, %4; add.u32 %2, %2, %4; add.u32 %3, %3, %4;"
                   : "+r"(x0), "+r"(x1), "+r"(x2), "+r"(x3) : "r"(u)), x0 + x1 + x2 + x3)
CHAIN(k_mix4, double x0 = a; double x1 = a + 1; unsigned x2 = u + 2; unsigned x3 = u + 3,
      asm volatile("add.f64 %0, %0, %4; add.u32 %2, %2, %5; add.f64 %1, %1, %4; add.u32 %3, %3, %5;"
                   : "+d"(x0), "+d"(x1), "+r"(x2), "+r"(x3) : "d"(b), "r"(u)), x0 + x1 + x2 + x3)
CHAIN(k_fsel4, unsigned x0 = u; unsigned x1 = u + 1; unsigned x2 = u + 2; unsigned x3 = u + 3,
      asm volatile("{.reg .pred p; setp.ne.u32 p, %4, 77; selp.u32 %0, %0, %4, p; selp.u32 %1, %1, %4, p; "
                   "selp.u32 %2, %2, %4, p; selp.u32 %3, %3, %4, p;}"
                   : "+r"(x0), "+r"(x1), "+r"(x2), "+r"(x3) : "r"(u)), x0 + x1 + x2 + x3)

__global__ void k_lds(long long *out, int iters) {
    __shared__ int next[64];
    if (threadIdx.x < 64) next[threadIdx.x] = (threadIdx.x + 1) & 63;
    __syncthreads();
    int p = threadIdx.x & 31;
    long long t0 = clock64();
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int r = 0; r < REP; r++) p = next[p];
    }
    long long t1 = clock64();
    if (threadIdx.x == 0) { out[0] = t1 - t0; out[1] = p; }
}


// ---- control flow the compiler cannot prove warp-uniform (trip count loaded per lane): *.sync ops get a BRA.DIV guard
__global__ void k_vote_div(long long *out, const int *trip, int zero, unsigned u) {
    const int lane = threadIdx.x & 31;
    const int n = trip[lane * zero];
    unsigned v = u;
    long long t0 = clock64();
    for (int it = 0; it < n; it++) {
#pragma unroll
        for (int r = 0; r < 16; r++) v = __ballot_sync(0xffffffffu, v + lane > 12345u);
        if (v == 0x12345u + lane) break;
    }
    long long t1 = clock64();
    if (threadIdx.x == 0) { out[0] = t1 - t0; out[1] = v; }
}
__global__ void k_vote_div_sw(long long *out, const int *trip, int zero, unsigned u) {
    const int lane = threadIdx.x & 31;
    const int n = trip[lane * zero];
    unsigned v = u;
    long long t0 = clock64();
    for (int it = 0; it < n; it++) {
#pragma unroll
        for (int r = 0; r < 16; r++) { __syncwarp(); v = __ballot_sync(0xffffffffu, v + lane > 12345u); }
        if (v == 0x12345u + lane) break;
    }
    long long t1 = clock64();
    if (threadIdx.x == 0) { out[0] = t1 - t0; out[1] = v; }
}
// rolled loop: one IADD + compare + taken back-edge per link
__global__ void k_backedge(long long *out, const int *trip, int zero, unsigned u) {
    const int lane = threadIdx.x & 31;
    const int n = trip[lane * zero] * 16;
    unsigned v = u + lane;
    long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < n; it++) { asm volatile("add.u32 %0, %0, %1;" : "+r"(v) : "r"(lane)); if (v == 0xdeadbeefu) break; }
    long long t1 = clock64();
    if (threadIdx.x == 0) { out[0] = t1 - t0; out[1] = v; }
}
// a TAKEN forward branch over a block per link
__global__ void k_skip(long long *out, const int *trip, int zero, unsigned u) {
    const int lane = threadIdx.x & 31;
    const int n = trip[lane * zero];
    unsigned v = u + lane;
    long long t0 = clock64();
    for (int it = 0; it < n; it++) {
#pragma unroll
        for (int r = 0; r < 16; r++) {
            if (v & 0x80000000u) {
#pragma unroll
                for (int q = 0; q < 24; q++) asm volatile("xor.b32 %0, %0, %1; shl.b32 %0, %0, 1;" : "+r"(v) : "r"(lane + q));
            }
            asm volatile("add.u32 %0, %0, %1;" : "+r"(v) : "r"(lane));
        }
    }
    long long t1 = clock64();
    if (threadIdx.x == 0) { out[0] = t1 - t0; out[1] = v; }
}
// 4 independent CREDUX / 4 independent SHFL per link: pipelining of the warp-wide units
__global__ void k_redux4(long long *out, const int *trip, int zero, unsigned u) {
    const int lane = threadIdx.x & 31;
    unsigned a = u + lane, b = u + 2 * lane, c = u + 3 * lane, d = u ^ lane;
    long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < 64; it++) {
#pragma unroll
        for (int r = 0; r < 16; r++) {
            unsigned ra = __reduce_min_sync(0xffffffffu, a), rb = __reduce_min_sync(0xffffffffu, b);
            unsigned rc = __reduce_min_sync(0xffffffffu, c), rd = __reduce_min_sync(0xffffffffu, d);
            a = ra + lane; b = rb + lane; c = rc + lane; d = rd + lane;
        }
    }
    long long t1 = clock64();
    if (threadIdx.x == 0) { out[0] = t1 - t0; out[1] = a + b + c + d; }
}
__global__ void k_shfl4(long long *out, const int *trip, int zero, unsigned u) {
    const int lane = threadIdx.x & 31;
    unsigned a = u + lane, b = u + 2 * lane, c = u + 3 * lane, d = u ^ lane;
    long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < 64; it++) {
#pragma unroll
        for (int r = 0; r < 16; r++) {
            const int src = a & 31;
            unsigned ra = __shfl_sync(0xffffffffu, a, src), rb = __shfl_sync(0xffffffffu, b, src);
            unsigned rc = __shfl_sync(0xffffffffu, c, src), rd = __shfl_sync(0xffffffffu, d, src);
            a = ra + 1; b = rb; c = rc; d = rd;
        }
    }
    long long t1 = clock64();
    if (threadIdx.x == 0) { out[0] = t1 - t0; out[1] = a + b + c + d; }
}
// winner lane three ways: ballot+ffs | redux over lane ids | ballot + fns-free clz of reversed layout
__global__ void k_win_ffs(long long *out, const int *trip, int zero, unsigned u) {
    const int lane = threadIdx.x & 31;
    unsigned v = u;
    long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < 64; it++) {
#pragma unroll
        for (int r = 0; r < 16; r++) { unsigned b = __ballot_sync(0xffffffffu, ((v + lane) & 7u) == 3u); v = __ffs(b) + v; }
    }
    long long t1 = clock64();
    if (threadIdx.x == 0) { out[0] = t1 - t0; out[1] = v; }
}
__global__ void k_win_redux(long long *out, const int *trip, int zero, unsigned u) {
    const int lane = threadIdx.x & 31;
    unsigned v = u;
    long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < 64; it++) {
#pragma unroll
        for (int r = 0; r < 16; r++) { unsigned w = __reduce_min_sync(0xffffffffu, (((v + lane) & 7u) == 3u) ? lane : 32); v = w + v; }
    }
    long long t1 = clock64();
    if (threadIdx.x == 0) { out[0] = t1 - t0; out[1] = v; }
}
__global__ void k_lds128(long long *out, int iters) {
    __shared__ int4 next[64];
    if (threadIdx.x < 64) next[threadIdx.x] = make_int4((threadIdx.x + 1) & 63, 1, 2, 3);
    __syncthreads();
    int p = threadIdx.x & 31;
    long long t0 = clock64();
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int r = 0; r < REP; r++) { int4 q = next[p]; p = q.x + (q.w - 3); }
    }
    long long t1 = clock64();
    if (threadIdx.x == 0) { out[0] = t1 - t0; out[1] = p; }
}

int main() {
    long long *d, h[2];
    cudaMalloc(&d, 64);
    const int iters = 64;
#define RUN(K, OPS, ...)                                                                          \
    for (int rep = 0; rep < 2; rep++) {                                                           \
        K<<<1, 32>>>(d, __VA_ARGS__);                                                             \
        cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);                                             \
    }                                                                                             \
    printf("%-14s %7.2f cycles per link (%d instr-equivalents per link)  [%s]\n", #K,             \
           h[0] / (double)(iters * REP), OPS, cudaGetErrorString(cudaGetLastError()));
    RUN(k_dadd, 1, 1.0, 1e-9, 3u, iters)
    RUN(k_dmul, 1, 1.0, 1.0000001, 3u, iters)
    RUN(k_dfma, 1, 1.0, 1e-9, 3u, iters)
    RUN(k_dmin, 3, 1.0, 2.0, 3u, iters)
    RUN(k_dsetp_sel, 3, 1.0, 2.0, 3u, iters)
    RUN(k_iadd, 1, 1.0, 2.0, 3u, iters)
    RUN(k_lop, 1, 1.0, 2.0, 3u, iters)
    RUN(k_isetp_sel, 2, 1.0, 2.0, 3u, iters)
    RUN(k_vote, 2, 1.0, 2.0, 3u, iters)
    RUN(k_redux, 2, 1.0, 2.0, 3u, iters)
    RUN(k_redux_cmp, 4, 1.0, 2.0, 3u, iters)
    RUN(k_shfl, 1, 1.0, 2.0, 3u, iters)
    RUN(k_ffs, 3, 1.0, 2.0, 3u, iters)
    RUN(k_popc, 2, 1.0, 2.0, 3u, iters)
    RUN(k_branch, 3, 1.0, 2.0, 3u, iters)
    RUN(k_dadd4, 4, 1.0, 1e-9, 3u, iters)
    RUN(k_iadd4, 4, 1.0, 2.0, 3u, iters)
    RUN(k_mix4, 4, 1.0, 1e-9, 3u, iters)
    RUN(k_fsel4, 5, 1.0, 2.0, 3u, iters)
    for (int rep = 0; rep < 2; rep++) { k_lds<<<1, 64>>>(d, iters); cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost); }
    printf("%-14s %7.2f cycles per link\n", "k_lds", h[0] / (double)(iters * REP));
    for (int rep = 0; rep < 2; rep++) { k_lds128<<<1, 64>>>(d, iters); cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost); }
    printf("%-14s %7.2f cycles per link (LDS.128 + IADD3)\n", "k_lds128", h[0] / (double)(iters * REP));
    int *trip, htrip = 64;
    cudaMalloc(&trip, 4);
    cudaMemcpy(trip, &htrip, 4, cudaMemcpyHostToDevice);
#define RUN2(K, LINKS, NOTE)                                                                      \
    for (int rep = 0; rep < 2; rep++) { K<<<1, 32>>>(d, trip, 0, 3u); cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost); } \
    printf("%-14s %7.2f cycles per link  %s [%s]\n", #K, h[0] / (double)(LINKS), NOTE, cudaGetErrorString(cudaGetLastError()));
    RUN2(k_vote_div, 64 * 16, "IADD+ISETP+VOTE behind a BRA.DIV guard")
    RUN2(k_vote_div_sw, 64 * 16, "same with __syncwarp() in front")
    RUN2(k_backedge, 64 * 16, "IADD+ISETP+taken back-edge")
    RUN2(k_skip, 64 * 16, "ISETP + taken forward branch + IADD")
    RUN2(k_redux4, 64 * 16, "4 independent CREDUX + IADD")
    RUN2(k_shfl4, 64 * 16, "4 independent SHFL.IDX + IADD")
    RUN2(k_win_ffs, 64 * 16, "ISETP + VOTE + ffs + IADD")
    RUN2(k_win_redux, 64 * 16, "ISETP + SEL + CREDUX + IADD")
    return 0;
}
