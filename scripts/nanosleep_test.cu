// How long does __nanosleep(ns) really sleep on this GPU?  (one thread per CTA polls, like queue_pop's wait loop)
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k(unsigned ns, int iters, long long *out) {
    if (threadIdx.x == 0) {
        long long t0 = clock64();
        for (int i = 0; i < iters; i++) __nanosleep(ns);
        out[blockIdx.x] = (clock64() - t0) / iters;
    }
    __syncthreads();
}
int main() {
    long long *d; cudaMalloc(&d, 8 * 1184);
    long long h[4];
    for (unsigned ns : {0u, 100u, 400u, 1600u, 6400u, 20000u, 100000u}) {
        k<<<1184, 128>>>(ns, 200, d);
        cudaDeviceSynchronize();
        cudaMemcpy(h, d, 32, cudaMemcpyDeviceToHost);
        printf("nanosleep(%u): %lld cycles per call (~%.0f ns at 1.9 GHz)\n", ns, h[0], h[0] / 1.9);
    }
    return 0;
}
