// Data-parallel primitives used by the potential solver (stage 2).
//
// Every solver kernel is written once as a `__host__ __device__` lambda over an index range and launched
// through pfor()/preduce().  The product build runs them as CUDA kernels on sm_100a.  Compiling the same
// translation unit with -DSSRS_HOST_EMU turns pfor() into a serial host loop: that build is TEST
// INFRASTRUCTURE ONLY (tests/hostemu.py builds it into tests/_build/, `-m "not gpu"` tests use it to check
// the solver's logic on CPU); nothing in ssrs_b200/ ever loads it, and the product library contains no
// host execution path.
#pragma once
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <vector>

#ifndef SSRS_HOST_EMU
#include <cuda_runtime.h>
#endif

namespace ssrs {
namespace par {

#ifdef SSRS_HOST_EMU
// ------------------------------------------------------------------------------------------------
#define SSRS_HD
typedef void* stream_t;

// workspace arenas (see the CUDA build below); on the host every block is its own malloc
struct Arena {
    std::vector<void*> blocks;
    size_t mark() const { return blocks.size(); }
    void* alloc(size_t bytes) { void* p = malloc(bytes ? bytes : 1); if (p) blocks.push_back(p); return p; }
    void release_to(size_t m) { while (blocks.size() > m) { free(blocks.back()); blocks.pop_back(); } }
};
inline Arena& arena(bool temp) { static thread_local Arena a[2]; return a[temp ? 1 : 0]; }
inline int dev_zero(void* p, size_t bytes, stream_t) { memset(p, 0, bytes); return 0; }
inline int dev_fill_byte(void* p, int v, size_t bytes, stream_t) { memset(p, v, bytes); return 0; }
inline int copy_d2d(void* d, const void* s, size_t bytes, stream_t) { memcpy(d, s, bytes); return 0; }
inline int copy_h2d(void* d, const void* s, size_t bytes, stream_t) { memcpy(d, s, bytes); return 0; }
inline int copy_d2h(void* d, const void* s, size_t bytes, stream_t) { memcpy(d, s, bytes); return 0; }
inline int sync(stream_t) { return 0; }

template <class F>
inline int pfor_range(int64_t i0, int64_t i1, stream_t, F f) {
    for (int64_t i = i0; i < i1; ++i) f(i);
    return 0;
}
template <class F>
inline int preduce_sum2_range(int64_t i0, int64_t i1, stream_t, double* out0, double* out1, F f) {
    double s0 = 0.0, s1 = 0.0;
    for (int64_t i = i0; i < i1; ++i) { double a, b; f(i, a, b); s0 += a; s1 += b; }
    *out0 = s0; *out1 = s1;
    return 0;
}
// one "warp" per row: partial(i, lane) for lane = 0..31 is summed and handed to finish(i, total)
template <class P, class Fin>
inline int pfor_warp_rows(int64_t i0, int64_t i1, stream_t, P partial, Fin finish) {
    for (int64_t i = i0; i < i1; ++i) {
        double part[32];
        for (int lane = 0; lane < 32; ++lane) part[lane] = partial(i, lane);
        for (int o = 16; o > 0; o >>= 1)
            for (int lane = 0; lane < o; ++lane) part[lane] += part[lane + o];      // the shuffle tree's order
        finish(i, part[0]);
    }
    return 0;
}
template <class F>
inline int pfor2d_rows(int r0, int r1, int cols, stream_t, F f) {
    for (int r = r0; r < r1; ++r)
        for (int c = 0; c < cols; ++c) f(r, c);
    return 0;
}
template <class F>
inline int preduce2d_sum2_rows(int r0, int r1, int cols, stream_t, double* out0, double* out1, F f) {
    double s0 = 0.0, s1 = 0.0;
    for (int r = r0; r < r1; ++r)
        for (int c = 0; c < cols; ++c) { double a, b; f(r, c, a, b); s0 += a; s1 += b; }
    *out0 = s0; *out1 = s1;
    return 0;
}
inline int exclusive_scan_i64(int64_t* data, int64_t n, int64_t* total, stream_t) {
    int64_t run = 0;
    for (int64_t i = 0; i < n; ++i) { int64_t v = data[i]; data[i] = run; run += v; }
    *total = run;
    return 0;
}
inline int atomic_add_int(int* p, int v) { int o = *p; *p = o + v; return o; }
inline void atomic_min_i64(int64_t* p, int64_t v) { if (v < *p) *p = v; }
inline void atomic_max_i64(int64_t* p, int64_t v) { if (v > *p) *p = v; }

#else
// ------------------------------------------------------------------------------------------------
#define SSRS_HD __host__ __device__
typedef cudaStream_t stream_t;

// Workspace arenas.  The solver needs ~300 B/cell (9 GB at 5000 x 6000) in ~150 blocks per solve.  Measured on
// B200: cudaMallocAsync from the default pool re-maps physical memory on most solves (setup 50 ms .. 4.7 s for
// the same grid), so the workspace is bump-allocated from a few large cudaMalloc chunks that stay cached per
// device between solves (one arena for blocks that live until the solve returns, one used as a stack for
// scoped temporaries); ssrs_release_workspace() returns them to the driver.
struct Arena {
    struct Chunk { char* base; size_t size, used; };
    std::vector<Chunk> chunks;
    int device = -1;
    static constexpr size_t ALIGN = 512, MIN_CHUNK = (size_t)256 << 20;
    // a mark encodes (chunk index, offset); allocation is strictly stack-like
    struct Mark { size_t chunk, used; };
    Mark mark() const { return chunks.empty() ? Mark{0, 0} : Mark{chunks.size() - 1, chunks.back().used}; }
    void* alloc(size_t bytes) {
        bytes = (bytes + ALIGN - 1) / ALIGN * ALIGN;
        if (bytes == 0) bytes = ALIGN;
        if (!chunks.empty() && chunks.back().size - chunks.back().used >= bytes) {
            void* p = chunks.back().base + chunks.back().used;
            chunks.back().used += bytes;
            return p;
        }
        size_t want = bytes > MIN_CHUNK ? bytes : MIN_CHUNK;
        if (!chunks.empty() && chunks.back().size > want) want = chunks.back().size;   // grow geometrically-ish
        void* base = nullptr;
        if (cudaMalloc(&base, want) != cudaSuccess) {
            cudaGetLastError();
            if (want == bytes || cudaMalloc(&base, bytes) != cudaSuccess) { cudaGetLastError(); return nullptr; }
            want = bytes;
        }
        chunks.push_back(Chunk{(char*)base, want, bytes});
        return base;
    }
    void release_to(Mark m) {
        if (chunks.empty()) return;
        for (size_t k = m.chunk + 1; k < chunks.size(); ++k) chunks[k].used = 0;
        if (m.chunk < chunks.size()) chunks[m.chunk].used = m.used;
    }
    size_t capacity() const { size_t t = 0; for (const Chunk& c : chunks) t += c.size; return t; }
    bool idle() const { for (const Chunk& c : chunks) if (c.used) return false; return true; }
    void free_all() { for (Chunk& c : chunks) cudaFree(c.base); chunks.clear(); }
    // one chunk of at least `bytes` before a solve starts (no-op when the cached capacity already suffices)
    void reserve(size_t bytes) {
        if (!idle()) return;
        if (chunks.size() == 1 && chunks[0].size >= bytes) return;
        const size_t have = capacity();
        if (chunks.size() > 1 || have < bytes) {
            free_all();
            void* base = nullptr;
            const size_t want = have > bytes ? have : bytes;
            if (cudaMalloc(&base, want) == cudaSuccess) chunks.push_back(Chunk{(char*)base, want, 0});
            else cudaGetLastError();        // fall back to growing on demand
        }
    }
};
Arena& arena(bool temp);      // per device and host thread (defined in potential.cu)
inline int dev_zero(void* p, size_t bytes, stream_t s) { return cudaMemsetAsync(p, 0, bytes, s) == cudaSuccess ? 0 : -1; }
inline int dev_fill_byte(void* p, int v, size_t bytes, stream_t s) { return cudaMemsetAsync(p, v, bytes, s) == cudaSuccess ? 0 : -1; }
inline int copy_d2d(void* d, const void* s, size_t b, stream_t st) { return cudaMemcpyAsync(d, s, b, cudaMemcpyDeviceToDevice, st) == cudaSuccess ? 0 : -1; }
inline int copy_h2d(void* d, const void* s, size_t b, stream_t st) { return cudaMemcpyAsync(d, s, b, cudaMemcpyHostToDevice, st) == cudaSuccess ? 0 : -1; }
inline int copy_d2h(void* d, const void* s, size_t b, stream_t st) { return cudaMemcpyAsync(d, s, b, cudaMemcpyDeviceToHost, st) == cudaSuccess ? 0 : -1; }
inline int sync(stream_t s) { return cudaStreamSynchronize(s) == cudaSuccess ? 0 : -1; }

int grid_cap();     // SM count x 8 (defined in potential.cu)

template <class F>
__global__ void __launch_bounds__(256) pfor_kernel(int64_t i0, int64_t i1, F f) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = i0 + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < i1; i += stride) f(i);
}
template <class F>
inline int pfor_range(int64_t i0, int64_t i1, stream_t s, F f) {
    const int64_t n = i1 - i0;
    if (n <= 0) return 0;
    int64_t blocks = (n + 255) / 256;
    if (blocks > grid_cap()) blocks = grid_cap();
    pfor_kernel<<<(int)blocks, 256, 0, s>>>(i0, i1, f);
    return cudaGetLastError() == cudaSuccess ? 0 : -1;
}

#ifndef SSRS_PFOR2D_MIN_BLOCKS
#define SSRS_PFOR2D_MIN_BLOCKS 8
#endif
// raster kernels: one thread per cell, 32 x 8 cells per CTA; f(row, col)
template <class F>
__global__ void __launch_bounds__(256, SSRS_PFOR2D_MIN_BLOCKS) pfor2d_kernel(int r0, int r1, int cols, F f) {
    const int c = blockIdx.x * 32 + threadIdx.x;
    const int r = r0 + blockIdx.y * 8 + threadIdx.y;
    if (r < r1 && c < cols) f(r, c);
}
template <class F>
inline int pfor2d_rows(int r0, int r1, int cols, stream_t s, F f) {
    if (r1 <= r0 || cols <= 0) return 0;
    dim3 grid((unsigned)((cols + 31) / 32), (unsigned)((r1 - r0 + 7) / 8));
    pfor2d_kernel<<<grid, dim3(32, 8), 0, s>>>(r0, r1, cols, f);
    return cudaGetLastError() == cudaSuccess ? 0 : -1;
}

// one warp per row (rows of the small coarse levels are long and few: a thread per row would walk its entries as a
// chain of dependent loads): partial(i, lane) is summed over the warp by a shuffle tree, lane 0 calls finish(i, total)
template <class P, class Fin>
__global__ void __launch_bounds__(256) pfor_warp_rows_kernel(int64_t i0, int64_t i1, P partial, Fin finish) {
    const int lane = threadIdx.x & 31;
    const int64_t warps = (int64_t)gridDim.x * (blockDim.x >> 5);
    for (int64_t i = i0 + (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); i < i1; i += warps) {
        double v = partial(i, lane);
        for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
        if (lane == 0) finish(i, v);
    }
}
template <class P, class Fin>
inline int pfor_warp_rows(int64_t i0, int64_t i1, stream_t s, P partial, Fin finish) {
    const int64_t n = i1 - i0;
    if (n <= 0) return 0;
    int64_t blocks = (n + 7) / 8;
    if (blocks > grid_cap()) blocks = grid_cap();
    pfor_warp_rows_kernel<<<(int)blocks, 256, 0, s>>>(i0, i1, partial, finish);
    return cudaGetLastError() == cudaSuccess ? 0 : -1;
}

constexpr int RED_BLOCKS = 1024;
double* reduce_scratch();       // device buffer of 2*RED_BLOCKS+2 doubles (defined in potential.cu)

__device__ __forceinline__ double block_sum(double v) {
    __shared__ double sh[8];
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
    __syncthreads();
    double t = (threadIdx.x < 8) ? sh[threadIdx.x] : 0.0;
    if (threadIdx.x < 32)
        for (int o = 4; o > 0; o >>= 1) t += __shfl_down_sync(0xffffffffu, t, o);
    return t;   // valid in thread 0
}
template <class F>
__global__ void __launch_bounds__(256) reduce2_kernel(int64_t i0, int64_t i1, double* partial, F f) {
    double s0 = 0.0, s1 = 0.0;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = i0 + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < i1; i += stride) {
        double a, b;
        f(i, a, b);
        s0 += a; s1 += b;
    }
    s0 = block_sum(s0);
    s1 = block_sum(s1);
    if (threadIdx.x == 0) { partial[blockIdx.x] = s0; partial[RED_BLOCKS + blockIdx.x] = s1; }
}
__global__ void __launch_bounds__(256) reduce_final_kernel(const double* partial, int nb, double* out) {
    double s0 = 0.0, s1 = 0.0;
    for (int i = threadIdx.x; i < nb; i += blockDim.x) { s0 += partial[i]; s1 += partial[RED_BLOCKS + i]; }
    s0 = block_sum(s0);
    s1 = block_sum(s1);
    if (threadIdx.x == 0) { out[0] = s0; out[1] = s1; }
}
// Deterministic two-stage reductions (fixed grid, fixed tree); the result is read back (stream sync).
template <class F>
inline int preduce_sum2_range(int64_t i0, int64_t i1, stream_t s, double* out0, double* out1, F f) {
    double* scratch = reduce_scratch();
    if (!scratch) return -1;
    int64_t blocks = (i1 - i0 + 255) / 256;
    if (blocks > RED_BLOCKS) blocks = RED_BLOCKS;
    if (blocks < 1) blocks = 1;
    reduce2_kernel<<<(int)blocks, 256, 0, s>>>(i0, i1, scratch, f);
    reduce_final_kernel<<<1, 256, 0, s>>>(scratch, (int)blocks, scratch + 2 * RED_BLOCKS);
    double h[2];
    if (cudaMemcpyAsync(h, scratch + 2 * RED_BLOCKS, sizeof(h), cudaMemcpyDeviceToHost, s) != cudaSuccess) return -1;
    if (cudaStreamSynchronize(s) != cudaSuccess) return -1;
    *out0 = h[0]; *out1 = h[1];
    return 0;
}
// the same over a raster: CTAs stride over 32 x 8 tiles in a fixed order
template <class F>
__global__ void __launch_bounds__(256, SSRS_PFOR2D_MIN_BLOCKS) reduce2d_kernel(int r0, int rows, int cols, int tiles_x, int64_t tiles, double* partial, F f) {
    double s0 = 0.0, s1 = 0.0;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    for (int64_t t = blockIdx.x; t < tiles; t += gridDim.x) {
        const int by = (int)(t / tiles_x), bx = (int)(t - (int64_t)by * tiles_x);
        const int r = r0 + by * 8 + ty, c = bx * 32 + tx;
        if (r < rows && c < cols) {
            double a, b;
            f(r, c, a, b);
            s0 += a; s1 += b;
        }
    }
    s0 = block_sum(s0);
    s1 = block_sum(s1);
    if (threadIdx.x == 0) { partial[blockIdx.x] = s0; partial[RED_BLOCKS + blockIdx.x] = s1; }
}
template <class F>
inline int preduce2d_sum2_rows(int r0, int r1, int cols, stream_t s, double* out0, double* out1, F f) {
    double* scratch = reduce_scratch();
    if (!scratch) return -1;
    const int tiles_x = (cols + 31) / 32;
    const int64_t tiles = (int64_t)tiles_x * ((r1 - r0 + 7) / 8);
    int64_t blocks = tiles < RED_BLOCKS ? tiles : RED_BLOCKS;
    if (blocks < 1) blocks = 1;
    reduce2d_kernel<<<(int)blocks, 256, 0, s>>>(r0, r1, cols, tiles_x, tiles, scratch, f);
    reduce_final_kernel<<<1, 256, 0, s>>>(scratch, (int)blocks, scratch + 2 * RED_BLOCKS);
    double h[2];
    if (cudaMemcpyAsync(h, scratch + 2 * RED_BLOCKS, sizeof(h), cudaMemcpyDeviceToHost, s) != cudaSuccess) return -1;
    if (cudaStreamSynchronize(s) != cudaSuccess) return -1;
    *out0 = h[0]; *out1 = h[1];
    return 0;
}
// In-place exclusive prefix sum of int64 (AMG setup: row pointers, aggregate numbering).  Three launches over tiles
// of 2048 items: per-tile sums (coalesced, strided items), ONE CTA that turns the tile sums into tile offsets and the
// grand total, then every tile scans itself (a thread owns 8 consecutive items) on top of its offset.  Integer sums:
// the result does not depend on the order.
constexpr int SCAN_THREADS = 256, SCAN_ITEMS = 8, SCAN_TILE = SCAN_THREADS * SCAN_ITEMS;

// exclusive scan of one value per thread over the CTA (blockDim.x a multiple of 32, <= 1024); *total on every thread
__device__ __forceinline__ int64_t block_exclusive_scan_i64(int64_t v, int64_t* total) {
    __shared__ int64_t warp_excl[32];
    __shared__ int64_t cta_total;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    int64_t incl = v;
    for (int o = 1; o < 32; o <<= 1) {
        const int64_t up = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += up;
    }
    __syncthreads();                                   // readers of a previous call are done with the shared words
    if (lane == 31) warp_excl[warp] = incl;            // this warp's sum
    __syncthreads();
    if (warp == 0) {
        const int64_t w = lane < nwarps ? warp_excl[lane] : 0;
        int64_t wi = w;
        for (int o = 1; o < 32; o <<= 1) {
            const int64_t up = __shfl_up_sync(0xffffffffu, wi, o);
            if (lane >= o) wi += up;
        }
        warp_excl[lane] = wi - w;                      // sum of the warps before this one
        if (lane == 31) cta_total = wi;
    }
    __syncthreads();
    *total = cta_total;
    return warp_excl[warp] + incl - v;
}
__global__ void __launch_bounds__(SCAN_THREADS) scan_tile_sums_kernel(const int64_t* __restrict__ data, int64_t n,
                                                                      int64_t* __restrict__ tile_sums) {
    const int64_t base = (int64_t)blockIdx.x * SCAN_TILE;
    int64_t v = 0;
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; ++k) {
        const int64_t i = base + (int64_t)k * SCAN_THREADS + threadIdx.x;
        if (i < n) v += data[i];
    }
    int64_t total;
    block_exclusive_scan_i64(v, &total);
    if (threadIdx.x == 0) tile_sums[blockIdx.x] = total;
}
// one CTA: tile_sums[0..ntiles) -> exclusive offsets in place, tile_sums[ntiles] = grand total
__global__ void __launch_bounds__(1024) scan_tile_offsets_kernel(int64_t* tile_sums, int64_t ntiles) {
    int64_t carry = 0;
    for (int64_t c0 = 0; c0 < ntiles; c0 += blockDim.x) {
        const int64_t i = c0 + threadIdx.x;
        const int64_t v = i < ntiles ? tile_sums[i] : 0;
        int64_t total;
        const int64_t excl = block_exclusive_scan_i64(v, &total);
        if (i < ntiles) tile_sums[i] = carry + excl;
        carry += total;
    }
    if (threadIdx.x == 0) tile_sums[ntiles] = carry;
}
__global__ void __launch_bounds__(SCAN_THREADS) scan_apply_kernel(int64_t* data, int64_t n, const int64_t* __restrict__ tile_offsets) {
    const int64_t first = (int64_t)blockIdx.x * SCAN_TILE + (int64_t)threadIdx.x * SCAN_ITEMS;
    int64_t item[SCAN_ITEMS];
    int64_t mine = 0;
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; ++k) {
        item[k] = first + k < n ? data[first + k] : 0;
        mine += item[k];
    }
    int64_t total;
    int64_t run = tile_offsets[blockIdx.x] + block_exclusive_scan_i64(mine, &total);
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; ++k) {
        if (first + k < n) data[first + k] = run;
        run += item[k];
    }
}
inline int exclusive_scan_i64(int64_t* data, int64_t n, int64_t* total, stream_t s) {
    if (n <= 0) { *total = 0; return 0; }
    const int64_t ntiles = (n + SCAN_TILE - 1) / SCAN_TILE;
    if (ntiles > 2147483647LL) return -1;
    Arena& ar = arena(true);
    const Arena::Mark mk = ar.mark();
    int64_t* tiles = (int64_t*)ar.alloc(sizeof(int64_t) * (size_t)(ntiles + 1));
    if (!tiles) return -1;
    scan_tile_sums_kernel<<<(unsigned)ntiles, SCAN_THREADS, 0, s>>>(data, n, tiles);
    scan_tile_offsets_kernel<<<1, 1024, 0, s>>>(tiles, ntiles);
    scan_apply_kernel<<<(unsigned)ntiles, SCAN_THREADS, 0, s>>>(data, n, tiles);
    cudaError_t e = cudaGetLastError();
    if (e == cudaSuccess) e = cudaMemcpyAsync(total, tiles + ntiles, 8, cudaMemcpyDeviceToHost, s);
    if (e == cudaSuccess) e = cudaStreamSynchronize(s);
    ar.release_to(mk);
    return e == cudaSuccess ? 0 : -1;
}
__host__ __device__ __forceinline__ int atomic_add_int(int* p, int v) {
#ifdef __CUDA_ARCH__
    return atomicAdd(p, v);
#else
    int o = *p; *p = o + v; return o;
#endif
}
__host__ __device__ __forceinline__ void atomic_min_i64(int64_t* p, int64_t v) {
#ifdef __CUDA_ARCH__
    atomicMin((long long*)p, (long long)v);
#else
    if (v < *p) *p = v;
#endif
}
__host__ __device__ __forceinline__ void atomic_max_i64(int64_t* p, int64_t v) {
#ifdef __CUDA_ARCH__
    atomicMax((long long*)p, (long long)v);
#else
    if (v > *p) *p = v;
#endif
}
#endif

// whole-range forms
template <class F> inline int pfor(int64_t n, stream_t s, F f) { return pfor_range(0, n, s, f); }
template <class F> inline int pfor2d(int rows, int cols, stream_t s, F f) { return pfor2d_rows(0, rows, cols, s, f); }
template <class F> inline int preduce_sum2(int64_t n, stream_t s, double* o0, double* o1, F f) { return preduce_sum2_range(0, n, s, o0, o1, f); }
template <class F> inline int preduce2d_sum2(int rows, int cols, stream_t s, double* o0, double* o1, F f) { return preduce2d_sum2_rows(0, rows, cols, s, o0, o1, f); }
template <class F>
struct OneOfTwoR {
    F f;
    SSRS_HD void operator()(int64_t i, double& a, double& b) const { a = f(i); b = 0.0; }
};
template <class F>
inline int preduce_sum_range(int64_t i0, int64_t i1, stream_t s, double* out, F f) {
    double dummy;
    return preduce_sum2_range(i0, i1, s, out, &dummy, OneOfTwoR<F>{f});
}
template <class F> inline int preduce_sum(int64_t n, stream_t s, double* out, F f) { return preduce_sum_range(0, n, s, out, f); }

}  // namespace par
}  // namespace ssrs
