// Shared-memory flag ping-pong between two warps of one CTA: the hand-off latency a multi-warp resolve pipeline would
// pay per pass.  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o pingpong pingpong.cu && ./pingpong
#include <cstdio>
#include <cuda_runtime.h>
template <bool FENCE>
__global__ void pingpong(long long *out, int iters) {
    __shared__ volatile int flagA, flagB;
    __shared__ volatile double payload[8];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) { flagA = 0; flagB = 0; }
    __syncthreads();
    long long t0 = clock64();
    double acc = 0.0;
    for (int i = 1; i <= iters; i++) {
        if (warp == 0) {
            if (lane < 8) payload[lane] = acc + lane;
            if (FENCE) __threadfence_block(); else __syncwarp();
            if (lane == 0) flagA = i;
            while (flagB != i) {}
            acc += payload[lane & 7];
        } else if (warp == 1) {
            while (flagA != i) {}
            double v = payload[lane & 7];
            if (lane < 8) payload[lane] = v * 1.0000001;
            if (FENCE) __threadfence_block(); else __syncwarp();
            if (lane == 0) flagB = i;
        }
    }
    long long t1 = clock64();
    if (threadIdx.x == 0) { out[0] = t1 - t0; out[1] = (long long)acc; }
}
int main() {
    long long *d, h[2];
    cudaMalloc(&d, 16);
    for (int rep = 0; rep < 2; rep++) {
        pingpong<true><<<1, 128>>>(d, 100000);
        cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
    }
    printf("round trip with __threadfence_block: %.1f cycles (%s)\n", h[0] / 100000.0, cudaGetErrorString(cudaGetLastError()));
    for (int rep = 0; rep < 2; rep++) {
        pingpong<false><<<1, 128>>>(d, 100000);
        cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
    }
    printf("round trip with __syncwarp only:     %.1f cycles (%s)\n", h[0] / 100000.0, cudaGetErrorString(cudaGetLastError()));
    return 0;
}
