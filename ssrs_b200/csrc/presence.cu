// Presence post-processing ("next" row f-1): disk-kernel smoothing of the visit counts.
//
// Replaces compute_smooth_presence_counts (ssrs/movmodel.py:422-439): convolve2d(counts, normalised disk,
// mode='same').  At 10 m resolution the default 1 km radius is a 201x201 disk, which the reference cannot
// evaluate (1.2e12 MACs per map); here each disk row is one difference of a per-row running sum, so a cell
// costs 2R+1 pairs of coalesced int64 loads instead of ~pi R^2 multiply-adds, and the integer sums are exact.
#include "common.cuh"

namespace ssrs {
namespace {

// prefix: int64 [rows][cols+1], prefix[r][c] = sum of counts[r][0..c-1]
__global__ void __launch_bounds__(256) smooth_disk_kernel(const long long* __restrict__ prefix, int rows, int cols,
                                                          int radius, double inv_area, float* __restrict__ out) {
    extern __shared__ int halfw[];                       // halfw[dy + radius]
    for (int i = threadIdx.x; i <= 2 * radius; i += blockDim.x) {
        const int dy = i - radius;
        const long long v = (long long)radius * radius - (long long)dy * dy;
        long long w = (long long)sqrt((double)v);
        while (w * w > v) --w;
        while ((w + 1) * (w + 1) <= v) ++w;
        halfw[i] = (int)w;
    }
    __syncthreads();
    const long long n = (long long)rows * cols;
    const long long stride = (long long)gridDim.x * blockDim.x;
    const int pitch = cols + 1;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const int r = (int)(i / cols), c = (int)(i - (long long)r * cols);
        long long acc = 0;
        const int y0 = max(r - radius, 0), y1 = min(r + radius, rows - 1);
        for (int y = y0; y <= y1; ++y) {
            const int w = halfw[y - r + radius];
            const int a = max(c - w, 0), b = min(c + w, cols - 1) + 1;
            const long long* p = prefix + (long long)y * pitch;
            acc += __ldg(p + b) - __ldg(p + a);
        }
        out[i] = (float)((double)acc * inv_area);
    }
}

// prefix[r][0] = 0, prefix[r][c + 1] = counts[r][0] + ... + counts[r][c]: one CTA per row, tiles of 256 columns scanned
// with warp shuffles, the tile total carried in a register (exact integer sums)
__global__ void __launch_bounds__(256) row_prefix_kernel(const unsigned* __restrict__ counts, int rows, int cols,
                                                         long long* __restrict__ prefix) {
    __shared__ long long warp_tot[8];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    for (int r = blockIdx.x; r < rows; r += gridDim.x) {
        const unsigned* src = counts + (long long)r * cols;
        long long* dst = prefix + (long long)r * (cols + 1);
        long long carry = 0;
        if (threadIdx.x == 0) dst[0] = 0;
        for (int c0 = 0; c0 < cols; c0 += 256) {
            const int c = c0 + threadIdx.x;
            long long v = c < cols ? (long long)src[c] : 0;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const long long u = __shfl_up_sync(0xffffffffu, v, o);
                if (lane >= o) v += u;
            }
            if (lane == 31) warp_tot[wid] = v;
            __syncthreads();
            long long before = 0, total = 0;
#pragma unroll
            for (int q = 0; q < 8; ++q) { const long long wt = warp_tot[q]; total += wt; if (q < wid) before += wt; }
            if (c < cols) dst[c + 1] = carry + before + v;
            carry += total;
            __syncthreads();
        }
    }
}

}  // namespace
}  // namespace ssrs

using namespace ssrs;

extern "C" int ssrs_row_prefix_sums(const uint32_t* counts, int rows, int cols, long long* row_prefix, void* stream) {
    SSRS_REQUIRE(counts && row_prefix, "ssrs_row_prefix_sums: NULL buffer");
    SSRS_REQUIRE(rows > 0 && cols > 0, "ssrs_row_prefix_sums: bad sizes");
    const int blocks = rows < sm_count() * 8 ? rows : sm_count() * 8;
    row_prefix_kernel<<<blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(counts, rows, cols, row_prefix);
    SSRS_CUDA_TRY(cudaGetLastError());
    return SSRS_OK;
}

extern "C" int ssrs_smooth_presence(const long long* row_prefix, int rows, int cols, int radius, float* out, void* stream) {
    SSRS_REQUIRE(row_prefix && out, "ssrs_smooth_presence: NULL buffer");
    SSRS_REQUIRE(rows > 0 && cols > 0 && radius >= 0, "ssrs_smooth_presence: bad sizes");
    SSRS_REQUIRE(radius <= 8192, "ssrs_smooth_presence: radius above 8192 cells");
    long long area = 0;                                      // number of cells with x^2 + y^2 <= r^2 (movmodel.py:432-435)
    for (int dy = -radius; dy <= radius; ++dy) {
        long long v = (long long)radius * radius - (long long)dy * dy, w = (long long)sqrt((double)v);
        while (w * w > v) --w;
        while ((w + 1) * (w + 1) <= v) ++w;
        area += 2 * w + 1;
    }
    long long blocks = cdiv((long long)rows * cols, 256);
    const long long cap = (long long)sm_count() * 8;
    if (blocks > cap) blocks = cap;
    smooth_disk_kernel<<<(int)blocks, 256, sizeof(int) * (2 * radius + 1), static_cast<cudaStream_t>(stream)>>>(
        row_prefix, rows, cols, radius, 1.0 / (double)area, out);
    SSRS_CUDA_TRY(cudaGetLastError());
    return SSRS_OK;
}
