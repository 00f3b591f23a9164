// Ceiling of a table-driven random walk on sm_100a: every lane walks its own chain of dependent gathers over a raster
// (5000 x 6000 cells), one record per (cell, previous move), one presence increment per step.  Measures what the L1TEX
// wavefront path, L2 and HBM deliver for this access pattern, independent of the stepping kernel's arithmetic.
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o walk walk.cu && ./walk
// Modes: T = one LDG.64 of a {T1, T2} threshold record per step + RED (the transition-table walk)
//        N = T without the RED          F = three LDG.64 gathers of {updraft, potential} + RED (the field-gather walk)
//        P = T + L2 prefetch of the three records the next step can read (latency regime)
//        Q = T + L1 (ld.ca dummy) look-ahead of the three next records
//        B = T with records in 8 x 8-cell blocks (DRAM page / TLB locality)
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

__device__ __forceinline__ void philox(unsigned c0, unsigned c1, unsigned c2, unsigned c3, unsigned k0, unsigned k1,
                                       unsigned& o0, unsigned& o1, unsigned& o2, unsigned& o3) {
#pragma unroll
    for (int i = 0; i < 10; ++i) {
        unsigned hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
        unsigned hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
        unsigned n0 = hi1 ^ c1 ^ k0, n1 = lo1, n2 = hi0 ^ c3 ^ k1, n3 = lo0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3; k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    o0 = c0; o1 = c1; o2 = c2; o3 = c3;
}
__device__ __forceinline__ unsigned mix(unsigned x) { x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16; return x; }

struct Lut { int off[8][3]; int slot[8][3]; };
// slot = flat move index 3(dr+1)+(dc+1) with the centre squeezed out; candidates = moves within 45 degrees
static void make_lut(Lut& L, int nc) {
    const int c3[9][3] = {{0,1,3},{0,1,2},{1,2,5},{0,3,6},{0,0,0},{2,5,8},{3,6,7},{6,7,8},{5,7,8}};
    for (int f = 0; f < 9; ++f) { if (f == 4) continue; int s = f - (f > 4);
        for (int j = 0; j < 3; ++j) { int i = c3[f][j]; L.off[s][j] = (i / 3 - 1) * nc + (i % 3 - 1); L.slot[s][j] = i - (i > 4); } }
}

__global__ void fill_table(uint2* tab, long long n) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x, st = (long long)gridDim.x * blockDim.x;
    for (; i < n; i += st) { unsigned h = mix((unsigned)i * 2654435761u + 17u); unsigned a = (h & 0x3FFFFFFF) + 0x10000000u;  // T1 in [0.125, 0.625)
        tab[i] = make_uint2(a, a + ((mix(h) & 0x1FFFFFFF))); }
}
__global__ void fill_fields(float2* f, long long n) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x, st = (long long)gridDim.x * blockDim.x;
    for (; i < n; i += st) f[i] = make_float2(1.0f + (mix((unsigned)i) & 1023) * 1e-3f, 1000.f - (float)(i / 6000) * 0.2f);
}

template <int MODE>   // 0 T, 1 N, 2 F, 3 P, 4 Q, 5 B
__global__ void __launch_bounds__(128) walk(const uint2* __restrict__ tab, const float2* __restrict__ fld, unsigned* pres, int nr, int nc,
                                            int steps, Lut lut, unsigned long long* sink, int northbias) {
    __shared__ int s_off[8][4];
    __shared__ int s_slot[8][4];
    if (threadIdx.x < 24) { int s = threadIdx.x / 3, j = threadIdx.x % 3; s_off[s][j] = lut.off[s][j]; s_slot[s][j] = lut.slot[s][j]; }
    __syncthreads();
    const unsigned gid = blockIdx.x * blockDim.x + threadIdx.x;
    int row = 100 + (mix(gid) % 100), col = 500 + (mix(gid ^ 0xabcdef) % 5000);
    int lin = row * nc + col;
    unsigned s = 6;    // heading north
    unsigned long long acc = 0;
    const int lo = 3 * nc, hi = (nr - 3) * nc;
    for (int k = 0; k < steps; k += 4) {
        unsigned w[4];
        philox(gid, 0, (unsigned)(k >> 2), 0, 0x1234567u, 0x89abcdeu, w[0], w[1], w[2], w[3]);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            int pick;
            if (MODE == 2) {
                const float2 f0 = __ldg(fld + lin + s_off[s][0]), f1 = __ldg(fld + lin + s_off[s][1]), f2 = __ldg(fld + lin + s_off[s][2]);
                const float a = f0.x * (1000.f - f0.y), b = f1.x * (1000.f - f1.y), c = f2.x * (1000.f - f2.y);
                const float t = (float)(w[j] >> 8) * (1.0f / 16777216.0f) * (a + b + c);
                pick = (a > t) ? 0 : ((a + b > t) ? 1 : 2);
            } else {
                long long idx;
                if (MODE == 5) { const int r = lin / nc, c = lin - r * nc; idx = ((long long)((r >> 3) * ((nc + 7) >> 3) + (c >> 3)) * 64 + ((r & 7) << 3) + (c & 7)) * 8 + s; }
                else idx = (long long)lin * 8 + s;
                uint2 rec;
                if (MODE == 4) asm volatile("ld.global.ca.v2.u32 {%0, %1}, [%2];" : "=r"(rec.x), "=r"(rec.y) : "l"(tab + idx));
                else rec = __ldg(tab + idx);
                if (MODE == 3 || MODE == 4) {
#pragma unroll
                    for (int q = 0; q < 3; ++q) {
                        const uint2* p = tab + ((long long)(lin + s_off[s][q]) * 8 + s_slot[s][q]);
                        if (MODE == 3) asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
                        else { unsigned d0, d1; asm volatile("ld.global.ca.v2.u32 {%0, %1}, [%2];" : "=r"(d0), "=r"(d1) : "l"(p)); acc += (d0 == 0x12345u); }
                    }
                }
                const unsigned r = w[j] >> 1;
                unsigned t1 = rec.x, t2 = rec.y;
                if (northbias) { t1 = 0x2AAAAAAAu; t2 = 0x55555555u; }
                pick = (r < t1) ? 0 : ((r < t2) ? 1 : 2);
                acc += rec.x & 1;
            }
            lin += s_off[s][pick];
            s = s_slot[s][pick];
            if (lin < lo || lin >= hi) { lin = (100 + (w[j] % 100)) * nc + 500 + (w[(j + 1) & 3] % 5000); s = 6; }
            if (MODE != 1) asm volatile("red.relaxed.gpu.global.add.u32 [%0], 1;" ::"l"(pres + lin) : "memory");
        }
    }
    if (acc == 0xFFFFFFFFFFFFULL) *sink = acc;
}

template <int MODE>
static void run(const char* name, const uint2* tab, const float2* fld, unsigned* pres, int nr, int nc, int blocks, int threads,
                int steps, const Lut& lut, unsigned long long* sink, int northbias = 0) {
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    walk<MODE><<<blocks, threads>>>(tab, fld, pres, nr, nc, steps / 4, lut, sink, northbias);   // warm-up
    CK(cudaDeviceSynchronize());
    CK(cudaEventRecord(e0));
    walk<MODE><<<blocks, threads>>>(tab, fld, pres, nr, nc, steps, lut, sink, northbias);
    CK(cudaEventRecord(e1)); CK(cudaDeviceSynchronize());
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
    const double total = (double)blocks * threads * steps;
    printf("%-34s blocks %5d x %3d steps %7d : %9.3f ms  %.3e steps/s  %.3f us/step/lane\n", name, blocks, threads, steps, ms,
           total / ms * 1e3, ms * 1e3 / steps);
}

int main() {
    const int nr = 5000, nc = 6000; const long long ncell = (long long)nr * nc;
    uint2* tab; float2* fld; unsigned* pres; unsigned long long* sink;
    const long long tab_n = (((long long)((nr + 7) >> 3) * ((nc + 7) >> 3)) * 64) * 8;
    CK(cudaMalloc(&tab, tab_n * sizeof(uint2))); CK(cudaMalloc(&fld, ncell * sizeof(float2)));
    CK(cudaMalloc(&pres, ncell * 4)); CK(cudaMalloc(&sink, 8)); CK(cudaMemset(pres, 0, ncell * 4));
    fill_table<<<148 * 8, 256>>>(tab, tab_n); fill_fields<<<148 * 8, 256>>>(fld, ncell); CK(cudaDeviceSynchronize());
    Lut lut; make_lut(lut, nc);
    int sm = 148;
    for (int per : {2, 4, 6, 8, 12, 16}) {
        char nm[64];
        snprintf(nm, 64, "T table+RED        %2d CTA/SM", per); run<0>(nm, tab, fld, pres, nr, nc, sm * per, 128, 4096, lut, sink);
    }
    for (int per : {6, 12, 16}) { char nm[64]; snprintf(nm, 64, "N table, no RED    %2d CTA/SM", per); run<1>(nm, tab, fld, pres, nr, nc, sm * per, 128, 4096, lut, sink); }
    for (int per : {6, 12, 16}) { char nm[64]; snprintf(nm, 64, "F 3 gathers+RED    %2d CTA/SM", per); run<2>(nm, tab, fld, pres, nr, nc, sm * per, 128, 4096, lut, sink); }
    for (int per : {6, 12, 16}) { char nm[64]; snprintf(nm, 64, "B blocked table+RED %2d CTA/SM", per); run<5>(nm, tab, fld, pres, nr, nc, sm * per, 128, 4096, lut, sink); }
    for (int per : {6, 12}) { char nm[64]; snprintf(nm, 64, "T north-coherent   %2d CTA/SM", per); run<0>(nm, tab, fld, pres, nr, nc, sm * per, 128, 4096, lut, sink, 1); }
    // latency regime: one warp, then one warp per SM
    run<0>("T lone warp", tab, fld, pres, nr, nc, 1, 32, 16384, lut, sink);
    run<2>("F lone warp", tab, fld, pres, nr, nc, 1, 32, 16384, lut, sink);
    run<3>("P lone warp (L2 prefetch)", tab, fld, pres, nr, nc, 1, 32, 16384, lut, sink);
    run<4>("Q lone warp (L1 look-ahead)", tab, fld, pres, nr, nc, 1, 32, 16384, lut, sink);
    run<5>("B lone warp (blocked)", tab, fld, pres, nr, nc, 1, 32, 16384, lut, sink);
    run<0>("T one warp per SM", tab, fld, pres, nr, nc, 148, 32, 16384, lut, sink);
    run<3>("P one warp per SM", tab, fld, pres, nr, nc, 148, 32, 16384, lut, sink);
    run<4>("Q one warp per SM", tab, fld, pres, nr, nc, 148, 32, 16384, lut, sink);
    run<3>("P bulk 12 CTA/SM", tab, fld, pres, nr, nc, 148 * 12, 128, 4096, lut, sink);
    return 0;
}
