// What does compute-sanitizer synccheck accept?  Named barriers with a subset of the CTA's warps, reached (a) from one program
// counter, (b) from two program counters, (c) through one noinline call, while the other warps wait on an mbarrier / flag.
// build: nvcc -arch=sm_100a -o synccheck_probe synccheck_probe.cu ; run: compute-sanitizer --tool synccheck ./synccheck_probe <variant>
#include <cstdio>
#include <cuda_runtime.h>
__device__ __noinline__ void nbar(unsigned id, unsigned n) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n) : "memory"); }
template <int V>
__global__ void k(int* out) {
    __shared__ volatile int flag;
    const int warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) flag = 0;
    __syncthreads();
    if (V == 0) {           // one PC, subset of warps
        if (warp < 2) asm volatile("bar.sync 1, 64;" ::: "memory");
    } else if (V == 1) {    // two PCs
        if (warp == 0) { out[1] = 1; asm volatile("bar.sync 1, 64;" ::: "memory"); }
        else if (warp == 1) { out[2] = 2; asm volatile("bar.sync 1, 64;" ::: "memory"); }
    } else if (V == 2) {    // one noinline call from two paths
        if (warp == 0) { out[1] = 1; nbar(1, 64); }
        else if (warp == 1) { out[2] = 2; nbar(1, 64); }
    } else if (V == 3) {    // two PCs + setmaxnreg-free role split where the other warps spin on a flag
        if (warp == 0) { out[1] = 1; asm volatile("bar.sync 1, 64;" ::: "memory"); if (threadIdx.x == 0) flag = 1; }
        else if (warp == 1) { out[2] = 2; asm volatile("bar.sync 1, 64;" ::: "memory"); }
        else { while (flag == 0) {} }
    }
    __syncthreads();
    if (threadIdx.x == 0) out[0] = V + 100;
}
int main(int argc, char** argv) {
    int v = argc > 1 ? atoi(argv[1]) : 0;
    int* d; cudaMalloc(&d, 64);
    if (v == 0) k<0><<<2, 128>>>(d); else if (v == 1) k<1><<<2, 128>>>(d); else if (v == 2) k<2><<<2, 128>>>(d); else k<3><<<2, 128>>>(d);
    cudaError_t e = cudaDeviceSynchronize();
    int h[4]; cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
    printf("variant %d: %s out0=%d\n", v, cudaGetErrorString(e), h[0]);
    return 0;
}
