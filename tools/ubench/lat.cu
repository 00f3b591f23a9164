// Single-warp dependent-chain latencies on sm_100a (cycles per op): what bounds the stepping kernel's tail, where one
// lane of one warp walks a long track alone.  nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o lat lat.cu
#include <cstdio>
#include <cuda_runtime.h>
#define N 4096
__global__ void k_dmul(double x, double y, long long* out, double* sink) {
    long long t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; ++i) x = __dmul_rn(x, y);
    long long t1 = clock64(); out[0] = t1 - t0; *sink = x;
}
__global__ void k_dadd(double x, double y, long long* out, double* sink) {
    long long t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; ++i) x = __dadd_rn(x, y);
    long long t1 = clock64(); out[0] = t1 - t0; *sink = x;
}
__global__ void k_f2f(float x, long long* out, double* sink) {   // f32->f64->f32 round trips: 2 conversions per trip
    long long t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; ++i) { double d; asm volatile("cvt.f64.f32 %0, %1;" : "=d"(d) : "f"(x)); asm volatile("cvt.rn.f32.f64 %0, %1;" : "=f"(x) : "d"(d)); }
    long long t1 = clock64(); out[0] = t1 - t0; *sink = x;
}
__global__ void k_dsetp(double x, double y, long long* out, double* sink) {  // DSETP -> FSEL pair (2 regs) chain
    long long t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; ++i) { x = (x > y) ? y : x + 0.0 * 0; asm volatile("" : "+d"(x)); double t = x; x = y; y = t; }
    long long t1 = clock64(); out[0] = t1 - t0; *sink = x + y;
}
__global__ void k_fmul(float x, float y, long long* out, double* sink) {
    long long t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; ++i) x = __fmul_rn(x, y);
    long long t1 = clock64(); out[0] = t1 - t0; *sink = x;
}
__global__ void k_imad(unsigned x, unsigned y, long long* out, double* sink) {
    long long t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; ++i) x = x * y + 12345u;
    long long t1 = clock64(); out[0] = t1 - t0; *sink = x;
}
__global__ void k_imadwide(unsigned x, long long* out, double* sink) {   // the Philox round: mul.hi/lo then xor
    unsigned c0 = x, c2 = x + 1;
    long long t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; ++i) {
        unsigned hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
        unsigned hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
        c0 = hi1 ^ lo0 ^ i; c2 = hi0 ^ lo1 ^ i;
    }
    long long t1 = clock64(); out[0] = t1 - t0; *sink = c0 + c2;
}
__global__ void k_lds(long long* out, double* sink) {
    __shared__ int s[1024];
    for (int i = threadIdx.x; i < 1024; i += 32) s[i] = (i * 37 + 11) & 1023;
    __syncwarp();
    int p = threadIdx.x;
    long long t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; ++i) p = s[p];
    long long t1 = clock64(); out[0] = t1 - t0; *sink = p;
}
// pointer chase through global memory with ld.global.nc; stride chosen by the host
__global__ void k_ldg(const unsigned* __restrict__ buf, int n, long long* out, double* sink) {
    unsigned p = threadIdx.x == 0 ? 0 : 0;
    long long t0 = clock64();
    for (int i = 0; i < n; ++i) p = __ldg(buf + p);
    long long t1 = clock64(); out[0] = t1 - t0; *sink = p;
}
// the step's f64 tail as in fast_step: q = (d*u)*(s*s), c1, c2, target, compare, select, feeding the next trip
__global__ void k_tail(double u0, double u1, double u2, double uu, long long* out, double* sink) {
    double uc = u0; float d = 1.5f;
    long long t0 = clock64();
#pragma unroll 4
    for (int i = 0; i < N; ++i) {
        double s0 = uc + u0, s1 = uc + u1, s2 = uc + u2;
        double q0 = ((double)d * u0) * (s1 * s2), q1 = ((double)(d + 1.f) * u1) * (s0 * s2), q2 = ((double)(d + 2.f) * u2) * (s0 * s1);
        double c1 = q0 + q1, c2 = c1 + q2, tg = uu * c2;
        bool a = q0 > tg, b = c1 > tg;
        uc = a ? u0 : (b ? u1 : u2);
        d = a ? 1.25f : (b ? 1.5f : 1.75f);
    }
    long long t1 = clock64(); out[0] = t1 - t0; *sink = uc + d;
}
int main() {
    long long* out; double* sink; cudaMalloc(&out, 8); cudaMalloc(&sink, 8);
    long long h;
#define RUN(name, call) call; call; cudaDeviceSynchronize(); cudaMemcpy(&h, out, 8, cudaMemcpyDeviceToHost); printf("%-28s %7.2f cycles/op\n", name, (double)h / N);
    RUN("DMUL dependent", (k_dmul<<<1, 32>>>(1.0000001, 1.0000001, out, sink)));
    RUN("DADD dependent", (k_dadd<<<1, 32>>>(1.0, 1e-9, out, sink)));
    RUN("F2F f32->f64->f32 (2 cvt)", (k_f2f<<<1, 32>>>(1.5f, out, sink)));
    RUN("DSETP+select swap", (k_dsetp<<<1, 32>>>(1.0, 2.0, out, sink)));
    RUN("FMUL dependent", (k_fmul<<<1, 32>>>(1.0000001f, 1.0000001f, out, sink)));
    RUN("IMAD dependent", (k_imad<<<1, 32>>>(3u, 5u, out, sink)));
    RUN("Philox-like round", (k_imadwide<<<1, 32>>>(3u, out, sink)));
    RUN("LDS pointer chase", (k_lds<<<1, 32>>>(out, sink)));
    RUN("step f64 tail (1 trip)", (k_tail<<<1, 32>>>(1.0, 2.0, 3.0, 0.6, out, sink)));
    // global pointer chase: 64 MB buffer, stride 4 KB+ -> L1 miss; second pass hits L2
    const int words = 16 << 20; unsigned* hbuf = new unsigned[words]; unsigned* dbuf; cudaMalloc(&dbuf, words * 4ull);
    const int stride = 1031 * 32;  // words
    for (int i = 0; i < words; ++i) hbuf[i] = 0;
    { unsigned p = 0; for (int i = 0; i < 2048; ++i) { unsigned nx = (p + stride) % words; hbuf[p] = nx; p = nx; } }
    cudaMemcpy(dbuf, hbuf, words * 4ull, cudaMemcpyHostToDevice);
    for (int pass = 0; pass < 3; ++pass) {
        k_ldg<<<1, 32>>>(dbuf, 2000, out, sink); cudaDeviceSynchronize(); cudaMemcpy(&h, out, 8, cudaMemcpyDeviceToHost);
        printf("LDG.nc chase pass %d (stride 132 KB) %7.2f cycles/load\n", pass, (double)h / 2000);
    }
    // small footprint: L1 hits
    for (int i = 0; i < 64; ++i) hbuf[i * 32] = ((i + 1) % 64) * 32;
    cudaMemcpy(dbuf, hbuf, 64 * 32 * 4, cudaMemcpyHostToDevice);
    for (int pass = 0; pass < 2; ++pass) {
        k_ldg<<<1, 32>>>(dbuf, 2000, out, sink); cudaDeviceSynchronize(); cudaMemcpy(&h, out, 8, cudaMemcpyDeviceToHost);
        printf("LDG.nc chase pass %d (8 KB footprint) %7.2f cycles/load\n", pass, (double)h / 2000);
    }
    return 0;
}
