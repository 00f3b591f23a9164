// Table walk ceiling, NON-coherent: walkers are spread uniformly over the whole 5000 x 6000 raster (the bulk regime of a
// queue-fed stepping kernel, where lanes hold tracks of all ages), so every record read is a random 32-byte sector of a
// 1.92 GB table and every presence increment a random sector of a 120 MB raster.  Variants: L2 cache-policy hints
// (table evict_first, presence evict_last), no presence, presence only, presence footprint.
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o walk2 walk2.cu && ./walk2
#include <cstdio>
#include <cstdlib>
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
static void make_lut(Lut& L, int nc) {
    const int c3[9][3] = {{0,1,3},{0,1,2},{1,2,5},{0,3,6},{0,0,0},{2,5,8},{3,6,7},{6,7,8},{5,7,8}};
    for (int f = 0; f < 9; ++f) { if (f == 4) continue; int s = f - (f > 4);
        for (int j = 0; j < 3; ++j) { int i = c3[f][j]; L.off[s][j] = (i / 3 - 1) * nc + (i % 3 - 1); L.slot[s][j] = i - (i > 4); } }
}
__global__ void fill_table(uint2* tab, long long n) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x, st = (long long)gridDim.x * blockDim.x;
    for (; i < n; i += st) { unsigned h = mix((unsigned)i * 2654435761u + 17u); unsigned a = (h & 0x3FFFFFFF) + 0x10000000u;
        tab[i] = make_uint2(a, a + ((mix(h) & 0x1FFFFFFF))); }
}

// TAB: 0 no table read (moves from the random word alone), 1 plain ld.global.nc, 2 + L2 evict_first policy, 3 + L1 no_allocate too
// RED: 0 none, 1 plain red, 2 red with L2 evict_last policy
// LAYOUT: 0 row-major records and counters; 1 both in 256-row x 128-column blocks (one 2 MB page of records per block)
// CLUSTER: 0 a CTA's walkers anywhere; 1 a CTA's walkers start in one 64-column band (rows anywhere)
template <int TAB, int RED, int LAYOUT = 0, int CLUSTER = 0>
__global__ void __launch_bounds__(128) walk(const uint2* __restrict__ tab, unsigned* pres, int nr, int nc, int steps, Lut lut,
                                            unsigned long long* sink, unsigned pres_mask) {
    __shared__ int s_off[8][4];
    __shared__ int s_slot[8][4];
    if (threadIdx.x < 24) { int s = threadIdx.x / 3, j = threadIdx.x % 3; s_off[s][j] = lut.off[s][j]; s_slot[s][j] = lut.slot[s][j]; }
    __syncthreads();
    unsigned long long pol_first, pol_last;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol_first));
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol_last));
    const unsigned gid = blockIdx.x * blockDim.x + threadIdx.x;
    const int span = nr - 8;
    int col0 = 4 + (int)(mix(gid ^ 0xabcdef) % (unsigned)(nc - 8));
    if (CLUSTER) col0 = 4 + (int)((blockIdx.x * 2654435761u) % (unsigned)(nc - 72)) + (int)(mix(gid ^ 0xabcdef) & 63);
    int lin = (4 + (int)(mix(gid) % (unsigned)span)) * nc + col0;
    const int ntc = (nc + 127) >> 7;
    unsigned s = 6;
    unsigned long long acc = 0;
    const int lo = 3 * nc, hi = (nr - 3) * nc;
    for (int k = 0; k < steps; k += 4) {
        unsigned w[4];
        philox(gid, 0, (unsigned)(k >> 2), 0, 0x1234567u, 0x89abcdeu, w[0], w[1], w[2], w[3]);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            unsigned t1 = 0x2AAAAAAAu, t2 = 0x55555555u;
            if (TAB) {
                long long cell = lin;
                if (LAYOUT) { const int r = lin / nc, c = lin - r * nc; cell = ((long long)((r >> 8) * ntc + (c >> 7)) << 15) + ((r & 255) << 7) + (c & 127); }
                const uint2* p = tab + (cell * 8 + s);
                uint2 rec;
                if (TAB == 1) rec = __ldg(p);
                else if (TAB == 2) asm volatile("ld.global.nc.L2::cache_hint.v2.u32 {%0, %1}, [%2], %3;" : "=r"(rec.x), "=r"(rec.y) : "l"(p), "l"(pol_first));
                else asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v2.u32 {%0, %1}, [%2], %3;" : "=r"(rec.x), "=r"(rec.y) : "l"(p), "l"(pol_first));
                t1 = rec.x; t2 = rec.y;
            }
            const unsigned r = w[j] >> 1;
            const int pick = (r < t1) ? 0 : ((r < t2) ? 1 : 2);
            lin += s_off[s][pick];
            s = s_slot[s][pick];
            if (lin < lo) lin += (nr - 8) * nc; else if (lin >= hi) lin -= (nr - 8) * nc;       // wrap: stay spread over all rows
            unsigned pidx = (unsigned)lin;
            if (LAYOUT) { const int r = lin / nc, c = lin - r * nc; pidx = (unsigned)((((r >> 8) * ntc + (c >> 7)) << 15) + ((r & 255) << 7) + (c & 127)); }
            if (RED == 1) asm volatile("red.relaxed.gpu.global.add.u32 [%0], 1;" ::"l"(pres + (pidx & pres_mask)) : "memory");
            if (RED == 2) asm volatile("red.relaxed.gpu.global.add.L2::cache_hint.u32 [%0], 1, %1;" ::"l"(pres + (pidx & pres_mask)), "l"(pol_last) : "memory");
            acc += t1 & 1;
        }
    }
    if (acc == 0xFFFFFFFFFFFFULL) *sink = acc;
}

template <int TAB, int RED, int LAYOUT = 0, int CLUSTER = 0>
static void run(const char* name, const uint2* tab, unsigned* pres, int nr, int nc, int blocks, int threads, int steps, const Lut& lut,
                unsigned long long* sink, unsigned pres_mask = 0xFFFFFFFFu) {
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    walk<TAB, RED, LAYOUT, CLUSTER><<<blocks, threads>>>(tab, pres, nr, nc, steps / 4, lut, sink, pres_mask);
    CK(cudaDeviceSynchronize());
    CK(cudaEventRecord(e0));
    walk<TAB, RED, LAYOUT, CLUSTER><<<blocks, threads>>>(tab, pres, nr, nc, steps, lut, sink, pres_mask);
    CK(cudaEventRecord(e1)); CK(cudaDeviceSynchronize());
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
    const double total = (double)blocks * threads * steps;
    printf("%-44s blocks %5d x %3d : %9.3f ms  %.3e steps/s  %.3f us/step/lane\n", name, blocks, threads, ms, total / ms * 1e3, ms * 1e3 / steps);
}

int main() {
    const int nr = 5000, nc = 6000; const long long ncell = (long long)nr * nc;
    uint2* tab; unsigned* pres; unsigned long long* sink;
    const long long nblk = (long long)((nr + 255) >> 8) * ((nc + 127) >> 7) << 15;      // cells incl. padding of the blocked layout
    CK(cudaMalloc(&tab, nblk * 8 * sizeof(uint2))); CK(cudaMalloc(&pres, nblk * 4)); CK(cudaMemset(pres, 0, nblk * 4)); CK(cudaMalloc(&sink, 8)); CK(cudaMemset(pres, 0, ncell * 4));
    fill_table<<<148 * 8, 256>>>(tab, nblk * 8); CK(cudaDeviceSynchronize());
    Lut lut; make_lut(lut, nc);
    const int sm = 148, steps = 2048;
    for (int per : {4, 8, 16}) {
        char nm[80];
        snprintf(nm, 80, "table only (plain)              %2d CTA/SM", per); run<1, 0>(nm, tab, pres, nr, nc, sm * per, 128, steps, lut, sink);
        snprintf(nm, 80, "RED only (plain)                %2d CTA/SM", per); run<0, 1>(nm, tab, pres, nr, nc, sm * per, 128, steps, lut, sink);
        snprintf(nm, 80, "table + RED (plain)             %2d CTA/SM", per); run<1, 1>(nm, tab, pres, nr, nc, sm * per, 128, steps, lut, sink);
        snprintf(nm, 80, "table evict_first + RED plain   %2d CTA/SM", per); run<2, 1>(nm, tab, pres, nr, nc, sm * per, 128, steps, lut, sink);
        snprintf(nm, 80, "table evict_first + RED evict_last %2d CTA/SM", per); run<2, 2>(nm, tab, pres, nr, nc, sm * per, 128, steps, lut, sink);
        snprintf(nm, 80, "table no_alloc/evict_first + RED evict_last %2d", per); run<3, 2>(nm, tab, pres, nr, nc, sm * per, 128, steps, lut, sink);
        snprintf(nm, 80, "table + RED on a 32 MB raster   %2d CTA/SM", per); run<1, 1>(nm, tab, pres, nr, nc, sm * per, 128, steps, lut, sink, (1u << 23) - 1);
        snprintf(nm, 80, "RED only, evict_last            %2d CTA/SM", per); run<0, 2>(nm, tab, pres, nr, nc, sm * per, 128, steps, lut, sink);
        snprintf(nm, 80, "table + RED, clustered CTAs     %2d CTA/SM", per); run<1, 1, 0, 1>(nm, tab, pres, nr, nc, sm * per, 128, steps, lut, sink);
        snprintf(nm, 80, "table + RED, blocked layout     %2d CTA/SM", per); run<1, 1, 1, 0>(nm, tab, pres, nr, nc, sm * per, 128, steps, lut, sink);
        snprintf(nm, 80, "table + RED, blocked + clustered %2d CTA/SM", per); run<1, 1, 1, 1>(nm, tab, pres, nr, nc, sm * per, 128, steps, lut, sink);
        snprintf(nm, 80, "table only, blocked + clustered %2d CTA/SM", per); run<1, 0, 1, 1>(nm, tab, pres, nr, nc, sm * per, 128, steps, lut, sink);
        snprintf(nm, 80, "RED only, blocked + clustered   %2d CTA/SM", per); run<0, 1, 1, 1>(nm, tab, pres, nr, nc, sm * per, 128, steps, lut, sink);
    }
    // persisting-L2 window over the presence raster
    {
        cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
        size_t want = prop.persistingL2CacheMaxSize;
        printf("L2 %d MB, persisting max %zu MB, window max %zu MB\n", prop.l2CacheSize >> 20, want >> 20, (size_t)prop.accessPolicyMaxWindowSize >> 20);
        CK(cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, want));
        cudaStream_t st; CK(cudaStreamCreate(&st));
        cudaStreamAttrValue attr = {};
        attr.accessPolicyWindow.base_ptr = pres;
        attr.accessPolicyWindow.num_bytes = (size_t)ncell * 4 < (size_t)prop.accessPolicyMaxWindowSize ? (size_t)ncell * 4 : (size_t)prop.accessPolicyMaxWindowSize;
        attr.accessPolicyWindow.hitRatio = (float)((double)want / (double)attr.accessPolicyWindow.num_bytes);
        if (attr.accessPolicyWindow.hitRatio > 1.f) attr.accessPolicyWindow.hitRatio = 1.f;
        attr.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
        attr.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
        CK(cudaStreamSetAttribute(st, cudaStreamAttributeAccessPolicyWindow, &attr));
        for (int per : {8, 16}) {
            cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
            walk<1, 1><<<sm * per, 128, 0, st>>>(tab, pres, nr, nc, steps / 4, lut, sink, 0xFFFFFFFFu);
            CK(cudaEventRecord(e0, st));
            walk<1, 1><<<sm * per, 128, 0, st>>>(tab, pres, nr, nc, steps, lut, sink, 0xFFFFFFFFu);
            CK(cudaEventRecord(e1, st)); CK(cudaStreamSynchronize(st));
            float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
            printf("table + RED, persisting window (ratio %.2f) %2d CTA/SM : %9.3f ms  %.3e steps/s\n", attr.accessPolicyWindow.hitRatio, per, ms,
                   (double)sm * per * 128 * steps / ms * 1e3);
        }
    }
    return 0;
}
