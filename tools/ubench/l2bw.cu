// Measured L2 read bandwidth: every SM streams an L2-resident buffer many times inside one kernel (float4 loads, no
// reuse in L1: ld.global.cg).  The denominator of bench.py's roofline_l2.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o l2bw l2bw.cu && ./l2bw
#include <cstdio>
#include <cuda_runtime.h>
__global__ void __launch_bounds__(256) rd(const float4* __restrict__ x, long long n, int reps, float* sink) {
    float acc = 0.f;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (int r = 0; r < reps; ++r)
        for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
            float4 v;
            asm volatile("ld.global.cg.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(x + i));
            acc += v.x + v.y + v.z + v.w;
        }
    if (acc == 123.456f) *sink = acc;
}
int main() {
    float* sink; cudaMalloc(&sink, 4);
    for (int mb : {16, 32, 48, 64, 96, 512}) {
        const long long n = (long long)mb * 1024 * 1024 / 16;
        float4* x; cudaMalloc(&x, n * 16); cudaMemset(x, 0, n * 16);
        const int reps = mb <= 96 ? 200 : 20;
        cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
        rd<<<148 * 8, 256>>>(x, n, 2, sink); cudaDeviceSynchronize();
        cudaEventRecord(e0); rd<<<148 * 8, 256>>>(x, n, reps, sink); cudaEventRecord(e1); cudaDeviceSynchronize();
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        printf("%4d MB buffer x %3d passes: %8.3f ms  %8.1f GB/s\n", mb, reps, ms, (double)n * 16 * reps / ms / 1e6);
        cudaFree(x);
    }
    return 0;
}
