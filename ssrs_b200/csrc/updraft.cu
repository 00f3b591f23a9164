// Stage 1 — fused slope / aspect / orographic updraft / threshold stencil (sm_100a).
//
// Replaces ssrs/layers.py:63-93 (compute_slope_degrees), :96-128 (compute_aspect_degrees),
// :11-22 (compute_orographic_updraft) and :171-185 (get_above_threshold_speed) with one pass over
// the DEM.  HBM-bound: 4 B read + 4 B per requested output plane per cell (20 B/cell with all four).
//
// Layout: a CTA owns a 32x128 output tile; the (32+2)x(128+8) halo tile of the DEM is staged in
// shared memory either by TMA (cp.async.bulk.tensor.2d, double-buffered behind mbarriers, persistent
// CTAs, out-of-bounds rows/cols zero-filled by the hardware) or, when the raster does not meet TMA's
// 16-byte pitch/alignment rules, by plain coalesced loads.  Each thread computes a 4-row x 4-col patch:
// one aligned float4 shared-memory read per row, halo columns exchanged with warp shuffles, three rows
// kept rolling in registers; outputs leave as float4 stores (512 B contiguous per warp and row).
//
// Numerics: Horn sums are formed differences-first, which is exact in fp32 for
// float32 DEMs, so `dz_dx == 0` fires on the same cells as the float64 reference; two adjacent cells share every
// packed FFMA2/FMUL2 (the kernel is issue-bound, not HBM-bound); the updraft itself is
// evaluated without inverse trig:  sin(atan h) = h / sqrt(1+h^2),
// cos(aspect - wdir) = -(gy cos(wdir) + gx' sin(wdir)) / hypot(gx', gy)   with gx' = (gx==0 ? 1e-10 : gx).
#include "common.cuh"

#include <cuda.h>
#include <math.h>

namespace ssrs {
namespace {

constexpr int TH = 32;            // tile rows
constexpr int TW = 128;           // tile cols
constexpr int HALO_L = 4;         // left halo kept 4 wide so each thread's float4 stays 16B-aligned
constexpr int SW = TW + 2 * HALO_L;   // 136 floats per staged row
constexpr int SH = TH + 2;            // 34 staged rows
constexpr int NTHREADS = 256;
constexpr int STAGES = 2;
constexpr uint32_t TILE_BYTES = SW * SH * sizeof(float);
// each TMA destination must be 128-byte aligned: pad the per-stage stride
constexpr int TILE_STRIDE = ((SW * SH + 31) / 32) * 32;

struct UpdraftParams {
    const float* dem;
    const float* wspeed;   // per-cell or null
    const float* wdirn;    // per-cell (degrees) or null
    float* slope;
    float* aspect;
    float* orograph;
    float* updraft;
    int rows, cols;
    int tiles_r, tiles_c;
    float inv8res;
    float uni_speed, uni_sin, uni_cos;   // uniform wind: speed, sin/cos of direction
    float thr, thr_inv, inv_em1;
    int vec_ok;            // cols % 4 == 0 and all pointers 16B aligned -> float4 global accesses
};

// raw MUFU approximations (<= 2 ulp, no denormal fix-up code around them)
__device__ __forceinline__ float rsqrt_approx(float x) { float y; asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float rcp_approx(float x) { float y; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }

__device__ __forceinline__ float threshold_fn(float w, float thr, float thr_inv, float inv_em1) {
    // layers.py:171-180: 0 if w <= 0.01 ; w if w > thr ; thr*(exp((w/thr)^5)-1)/(e-1) otherwise.
    // exp(x)-1 on (0,1] as x*P(x), degree-6 minimax fit, relative error 1.3e-7 (keeps the tiny values that
    // act as insulating films in the potential solve accurate, which exp(x)-1 in float32 would not).
    if (!(w > 0.01f)) return 0.0f;
    if (w > thr) return w;
    const float t = w * thr_inv;
    const float t2 = t * t;
    const float x = t2 * t2 * t;
    float p = 0.00031020541791804135f;
    p = fmaf(p, x, 0.0012223972007632256f);
    p = fmaf(p, x, 0.008451635017991066f);
    p = fmaf(p, x, 0.041624147444963455f);
    p = fmaf(p, x, 0.16667389869689941f);
    p = fmaf(p, x, 0.4999995529651642f);
    p = fmaf(p, x, 1.0f);
    return thr * inv_em1 * (p * x);
}

// ---- two cells at a time: packed float32 pairs (FFMA2 / FMUL2 / FADD2 on sm_100a; constants ride as broadcast immediates) ----
// The stencil kernels are issue-bound (ncu: 74 % issue-active at 64 % of HBM); two thirds of a cell's instructions are
// FMUL / FFMA / FADD, and a packed instruction does two cells' worth.  Each lane of a packed operation rounds exactly
// like the scalar one.
struct F2 { unsigned long long v; };
__device__ __forceinline__ F2 pk(float lo, float hi) { F2 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r.v) : "f"(lo), "f"(hi)); return r; }
__device__ __forceinline__ F2 pk1(float x) { return pk(x, x); }
__device__ __forceinline__ void upk(F2 a, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(a.v)); }
__device__ __forceinline__ F2 fma2(F2 a, F2 b, F2 c) { F2 r; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r.v) : "l"(a.v), "l"(b.v), "l"(c.v)); return r; }
__device__ __forceinline__ F2 mul2(F2 a, F2 b) { F2 r; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r.v) : "l"(a.v), "l"(b.v)); return r; }
__device__ __forceinline__ F2 add2(F2 a, F2 b) { F2 r; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r.v) : "l"(a.v), "l"(b.v)); return r; }

// atan(t) for t in [0, 1]: t * P(t^2), degree-6 minimax fit (max error 3.2e-7 rad = 1.8e-5 degrees,
// i.e. ~1e-6 of the slope range and ~5e-8 of the aspect range; the contract is 1e-5 of each field's maximum).
__device__ __forceinline__ F2 atan01_2(F2 t) {
    const F2 s = mul2(t, t);
    F2 p = pk1(0.006811788771301508f);
    p = fma2(p, s, pk1(-0.0336042158305645f));
    p = fma2(p, s, pk1(0.07962367683649063f));
    p = fma2(p, s, pk1(-0.132333442568779f));
    p = fma2(p, s, pk1(0.19807817041873932f));
    p = fma2(p, s, pk1(-0.3331736922264099f));
    p = fma2(p, s, pk1(0.9999961256980896f));
    return mul2(p, t);
}

// threshold_fn on both lanes (layers.py:171-180)
__device__ __forceinline__ void threshold2(float w0, float w1, float thr, float thr_inv, float thr_scale, float& u0, float& u1) {
    const F2 t = mul2(pk(w0, w1), pk1(thr_inv));
    const F2 t2 = mul2(t, t);
    const F2 x = mul2(mul2(t2, t2), t);
    F2 p = pk1(0.00031020541791804135f);
    p = fma2(p, x, pk1(0.0012223972007632256f));
    p = fma2(p, x, pk1(0.008451635017991066f));
    p = fma2(p, x, pk1(0.041624147444963455f));
    p = fma2(p, x, pk1(0.16667389869689941f));
    p = fma2(p, x, pk1(0.4999995529651642f));
    p = fma2(p, x, pk1(1.0f));
    float m0, m1;
    upk(mul2(mul2(p, x), pk1(thr_scale)), m0, m1);
    u0 = !(w0 > 0.01f) ? 0.0f : (w0 > thr ? w0 : m0);
    u1 = !(w1 > 0.01f) ? 0.0f : (w1 > thr ? w1 : m1);
}

// Two horizontally adjacent cells (a, b) from their Horn gradients gx = dz_dx (derivative along axis 0, rows) and
// gy = dz_dy (along axis 1, columns), layers.py:89-90.  Square roots and quotients go through MUFU.RSQ / MUFU.RCP
// (<= 2 ulp), inverse tangents through atan01_2.
template <bool WANT_ANGLES>
__device__ __forceinline__ void cell2(float gxa_, float gya_, float gxb_, float gyb_, F2 V, F2 sinw, F2 cosw,
                                      float& slope_a, float& slope_b, float& aspect_a, float& aspect_b, float& oro_a, float& oro_b) {
    const F2 gx = pk(gxa_, gxb_), gy = pk(gya_, gyb_);
    const F2 gy2 = mul2(gy, gy);
    const F2 h2 = fma2(gx, gx, gy2);
    float h2a, h2b;
    upk(h2, h2a, h2b);
    const F2 hraw = mul2(h2, pk(rsqrt_approx(h2a), rsqrt_approx(h2b)));
    float ha, hb;
    upk(hraw, ha, hb);
    ha = h2a > 1e-30f ? ha : 0.0f;
    hb = h2b > 1e-30f ? hb : 0.0f;
    const float gxaa = (gxa_ == 0.0f) ? 1e-10f : gxa_;                    // layers.py:124
    const float gxab = (gxb_ == 0.0f) ? 1e-10f : gxb_;
    const F2 gxa = pk(gxaa, gxab);
    float q2a, q2b;
    upk(fma2(gxa, gxa, gy2), q2a, q2b);
    const F2 rha = pk(rsqrt_approx(q2a), rsqrt_approx(q2b));
    // cos(aspect - wdir) = -(gy cos(wdir) + gx' sin(wdir)) / hypot(gx', gy)
    float cda, cdb;
    upk(mul2(fma2(gxa, sinw, mul2(gy, cosw)), rha), cda, cdb);
    float opa, opb;
    upk(add2(h2, pk1(1.0f)), opa, opb);
    const F2 sins = mul2(pk(ha, hb), pk(rsqrt_approx(opa), rsqrt_approx(opb)));      // sin(atan(h))
    float oa, ob;
    upk(mul2(mul2(V, sins), pk(fmaxf(0.0f, -cda), fmaxf(0.0f, -cdb))), oa, ob);
    oro_a = fmaxf(0.0f, oa);                                              // layers.py:19-22
    oro_b = fmaxf(0.0f, ob);
    if (WANT_ANGLES) {
        const bool steep_a = ha > 1.0f, steep_b = hb > 1.0f;
        float a0, a1;
        upk(atan01_2(pk(steep_a ? rcp_approx(ha) : ha, steep_b ? rcp_approx(hb) : hb)), a0, a1);
        a0 = steep_a ? 1.5707963267948966f - a0 : a0;
        a1 = steep_b ? 1.5707963267948966f - a1 : a1;
        upk(mul2(pk(a0, a1), pk1(57.29577951308232f)), slope_a, slope_b);
        const float aya = fabsf(gya_), axa = fabsf(gxaa), ayb = fabsf(gyb_), axb = fabsf(gxab);
        float p0, p1;
        upk(atan01_2(mul2(pk(fminf(aya, axa), fminf(ayb, axb)), pk(rcp_approx(fmaxf(aya, axa)), rcp_approx(fmaxf(ayb, axb))))), p0, p1);
        p0 = aya > axa ? 1.5707963267948966f - p0 : p0;
        p1 = ayb > axb ? 1.5707963267948966f - p1 : p1;
        const float ang0 = ((gya_ < 0.0f) != (gxaa < 0.0f)) ? -p0 : p0;   // atan(dz_dy / dz_dx)
        const float ang1 = ((gyb_ < 0.0f) != (gxab < 0.0f)) ? -p1 : p1;
        // aspect = 180 - ang * 57.29... + copysign(90, gx')              layers.py:125-127
        upk(fma2(pk(ang0, ang1), pk1(-57.29577951308232f), pk(180.0f + copysignf(90.0f, gxaa), 180.0f + copysignf(90.0f, gxab))),
            aspect_a, aspect_b);
    } else {
        slope_a = slope_b = aspect_a = aspect_b = 0.0f;
    }
}

struct Row6 { float l; float4 v; float r; };

__device__ __forceinline__ Row6 load_row(const float* srow, int lane) {
    Row6 o;
    o.v = *reinterpret_cast<const float4*>(srow + HALO_L + 4 * lane);
    float lft = __shfl_up_sync(0xffffffffu, o.v.w, 1);
    float rgt = __shfl_down_sync(0xffffffffu, o.v.x, 1);
    o.l = (lane == 0) ? srow[HALO_L - 1] : lft;
    o.r = (lane == 31) ? srow[HALO_L + TW] : rgt;
    return o;
}

template <bool VEC>
__device__ __forceinline__ void store4(float* plane, int64_t idx, int c0, int cols, float a, float b, float c, float d) {
    if (plane == nullptr) return;
    if (VEC) {
        if (c0 + 3 < cols) {
            __stcs(reinterpret_cast<float4*>(plane + idx), make_float4(a, b, c, d));
            return;
        }
    }
    if (c0 < cols) plane[idx] = a;
    if (c0 + 1 < cols) plane[idx + 1] = b;
    if (c0 + 2 < cols) plane[idx + 2] = c;
    if (c0 + 3 < cols) plane[idx + 3] = d;
}

// Computes one staged tile.  `tile` points at SH x SW floats; tile origin (r0, c0) in the raster.
template <bool VEC>
__device__ __forceinline__ void compute_tile(const UpdraftParams& p, const float* tile, int r0, int c0t) {
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const int c0 = c0t + 4 * lane;
    const int rbase = r0 + 4 * warp;
    if (rbase >= p.rows) return;                      // warp-uniform
    const float* s = tile + (4 * warp) * SW;          // staged row index = (row - r0) + 1
    const bool want_angles = (p.slope != nullptr) || (p.aspect != nullptr);
    Row6 below = load_row(s, lane);
    Row6 mid = load_row(s + SW, lane);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        Row6 above = load_row(s + (i + 2) * SW, lane);
        const int r = rbase + i;
        if (r < p.rows) {                             // warp-uniform
            float sl[4], as[4], oro[4], up[4];
            const float S[6] = {below.l, below.v.x, below.v.y, below.v.z, below.v.w, below.r};
            const float M[6] = {mid.l, mid.v.x, mid.v.y, mid.v.z, mid.v.w, mid.r};
            const float N[6] = {above.l, above.v.x, above.v.y, above.v.z, above.v.w, above.r};
            const int64_t idx = (int64_t)r * p.cols + c0;
            float V[4] = {p.uni_speed, p.uni_speed, p.uni_speed, p.uni_speed};
            float sn[4] = {p.uni_sin, p.uni_sin, p.uni_sin, p.uni_sin};
            float cs[4] = {p.uni_cos, p.uni_cos, p.uni_cos, p.uni_cos};
            if (p.wspeed != nullptr) {
                float wd[4];
                if (VEC && c0 + 3 < p.cols) {
                    float4 a = __ldcs(reinterpret_cast<const float4*>(p.wspeed + idx));
                    float4 b = __ldcs(reinterpret_cast<const float4*>(p.wdirn + idx));
                    V[0] = a.x; V[1] = a.y; V[2] = a.z; V[3] = a.w;
                    wd[0] = b.x; wd[1] = b.y; wd[2] = b.z; wd[3] = b.w;
                } else {
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        bool ok = c0 + j < p.cols;
                        V[j] = ok ? p.wspeed[idx + j] : 0.0f;
                        wd[j] = ok ? p.wdirn[idx + j] : 0.0f;
                    }
                }
#pragma unroll
                for (int j = 0; j < 4; ++j) sincospif(wd[j] * (1.0f / 180.0f), &sn[j], &cs[j]);
            }
            const bool edge_row = (r == 0) || (r == p.rows - 1);
            // Horn gradients of the four cells, differences first (exact for float32 DEMs): dz_dx along rows, dz_dy
            // along columns (layers.py:80-90)
            float gx[4], gy[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                gx[j] = ((N[j + 2] - S[j + 2]) + 2.0f * (N[j + 1] - S[j + 1]) + (N[j] - S[j])) * p.inv8res;
                gy[j] = ((S[j + 2] - S[j]) + 2.0f * (M[j + 2] - M[j]) + (N[j + 2] - N[j])) * p.inv8res;
            }
            const float thr_scale = p.thr * p.inv_em1;
#pragma unroll
            for (int j = 0; j < 4; j += 2) {
                if (want_angles)
                    cell2<true>(gx[j], gy[j], gx[j + 1], gy[j + 1], pk(V[j], V[j + 1]), pk(sn[j], sn[j + 1]), pk(cs[j], cs[j + 1]),
                                sl[j], sl[j + 1], as[j], as[j + 1], oro[j], oro[j + 1]);
                else
                    cell2<false>(gx[j], gy[j], gx[j + 1], gy[j + 1], pk(V[j], V[j + 1]), pk(sn[j], sn[j + 1]), pk(cs[j], cs[j + 1]),
                                 sl[j], sl[j + 1], as[j], as[j + 1], oro[j], oro[j + 1]);
#pragma unroll
                for (int q = j; q < j + 2; ++q) {
                    const int c = c0 + q;
                    if (edge_row || c == 0 || c >= p.cols - 1) { sl[q] = 0.0f; as[q] = 0.0f; oro[q] = 0.0f; }
                }
                threshold2(oro[j], oro[j + 1], p.thr, p.thr_inv, thr_scale, up[j], up[j + 1]);
            }
            store4<VEC>(p.slope, idx, c0, p.cols, sl[0], sl[1], sl[2], sl[3]);
            store4<VEC>(p.aspect, idx, c0, p.cols, as[0], as[1], as[2], as[3]);
            store4<VEC>(p.orograph, idx, c0, p.cols, oro[0], oro[1], oro[2], oro[3]);
            store4<VEC>(p.updraft, idx, c0, p.cols, up[0], up[1], up[2], up[3]);
        }
        below = mid;
        mid = above;
    }
}

// ---- plain staging: one tile per CTA --------------------------------------------------------
template <bool VEC>
__global__ void __launch_bounds__(NTHREADS) updraft_plain_kernel(const UpdraftParams p) {
    __shared__ __align__(16) float tile[SH * SW];
    const int tr = blockIdx.x / p.tiles_c;
    const int tc = blockIdx.x - tr * p.tiles_c;
    const int r0 = tr * TH, c0 = tc * TW;
    for (int i = threadIdx.x; i < SH * SW; i += NTHREADS) {
        int sr = i / SW, sc = i - sr * SW;
        int r = r0 - 1 + sr, c = c0 - HALO_L + sc;
        float v = 0.0f;
        if (r >= 0 && r < p.rows && c >= 0 && c < p.cols) v = __ldg(p.dem + (int64_t)r * p.cols + c);
        tile[i] = v;
    }
    __syncthreads();
    compute_tile<VEC>(p, tile, r0, c0);
}

// ---- TMA staging: persistent CTAs, two-stage mbarrier pipeline ---------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* ptr) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(ptr));
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred P1;\n\t"
        "WAIT_LOOP:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t"
        "@P1 bra DONE;\n\t"
        "bra WAIT_LOOP;\n\t"
        "DONE:\n\t}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, uint64_t* bar, int x, int y) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(x), "r"(y)
        : "memory");
}

template <bool VEC>
__global__ void __launch_bounds__(NTHREADS, 4) updraft_tma_kernel(const UpdraftParams p,
                                                               const __grid_constant__ CUtensorMap dem_map) {
    __shared__ __align__(128) float tiles[STAGES][TILE_STRIDE];
    __shared__ __align__(8) uint64_t full[STAGES];
    const int ntiles = p.tiles_r * p.tiles_c;
    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; ++s) mbar_init(&full[s], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    auto issue = [&](int t, int stage) {
        const int tr = t / p.tiles_c, tc = t - tr * p.tiles_c;
        mbar_expect_tx(&full[stage], TILE_BYTES);
        tma_load_2d(tiles[stage], &dem_map, &full[stage], tc * TW - HALO_L, tr * TH - 1);
    };
    int t = blockIdx.x;
    if (threadIdx.x == 0 && t < ntiles) issue(t, 0);
    uint32_t phase[STAGES] = {0, 0};
    int stage = 0;
    for (; t < ntiles; t += gridDim.x) {
        const int tnext = t + gridDim.x;
        if (threadIdx.x == 0 && tnext < ntiles) issue(tnext, stage ^ 1);   // buffer freed by the sync below
        mbar_wait(&full[stage], phase[stage]);
        phase[stage] ^= 1;
        const int tr = t / p.tiles_c, tc = t - tr * p.tiles_c;
        compute_tile<VEC>(p, tiles[stage], tr * TH, tc * TW);
        __syncthreads();     // everyone done reading tiles[stage] before it is refilled next-next round
        stage ^= 1;
    }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode_fn() {
    static EncodeTiledFn fn = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void* sym = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(sym);
    }
    return fn;
}

__global__ void threshold_kernel(const float* __restrict__ in, float* __restrict__ out, int64_t n, float thr,
                                 float thr_inv, float inv_em1) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (; i < n; i += stride) out[i] = threshold_fn(in[i], thr, thr_inv, inv_em1);
}

// compute_orographic_updraft (ssrs/layers.py:11-22) on already computed slope / aspect rasters (degrees):
// max(V sin(slope) max(cos(aspect - dirn), 0), min_val), per-cell or uniform wind
__global__ void orographic_kernel(const float* __restrict__ slope, const float* __restrict__ aspect,
                                  const float* __restrict__ wspeed, const float* __restrict__ wdirn, float ws, float wd,
                                  float min_val, float* __restrict__ out, int64_t n) {
    const float d2r = 0.017453292519943295f;
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (; i < n; i += stride) {
        const float v = wspeed ? wspeed[i] : ws, dir = wdirn ? wdirn[i] : wd;
        const float c = fmaxf(cosf((aspect[i] - dir) * d2r), 0.0f);
        out[i] = fmaxf(v * (sinf(slope[i] * d2r) * c), min_val);
    }
}

inline bool aligned16(const void* p) { return p == nullptr || (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

}  // namespace
}  // namespace ssrs

using namespace ssrs;

// SSRS_STENCIL_PATH=plain|tma forces a staging path (tests); default: TMA when eligible.
extern "C" int ssrs_updraft(const float* dem, int rows, int cols, float resolution, const float* wspeed,
                            const float* wdirn, float uniform_wspeed, float uniform_wdirn_deg, float threshold,
                            float* slope_deg, float* aspect_deg, float* orograph, float* updraft, void* stream) {
    SSRS_REQUIRE(dem != nullptr, "ssrs_updraft: dem is NULL");
    SSRS_REQUIRE(rows >= 3 && cols >= 3, "ssrs_updraft: grid %dx%d is smaller than the 3x3 stencil", rows, cols);
    SSRS_REQUIRE(resolution > 0.0f, "ssrs_updraft: resolution must be positive");
    SSRS_REQUIRE((wspeed == nullptr) == (wdirn == nullptr), "ssrs_updraft: wspeed and wdirn must both be given or both NULL");
    SSRS_REQUIRE(threshold > 0.0f, "ssrs_updraft: threshold must be positive");
    UpdraftParams p;
    p.dem = dem; p.wspeed = wspeed; p.wdirn = wdirn;
    p.slope = slope_deg; p.aspect = aspect_deg; p.orograph = orograph; p.updraft = updraft;
    p.rows = rows; p.cols = cols;
    p.tiles_r = (int)cdiv(rows, TH); p.tiles_c = (int)cdiv(cols, TW);
    p.inv8res = 1.0f / (8.0f * resolution);
    const double wd = (double)uniform_wdirn_deg * M_PI / 180.0;
    p.uni_speed = uniform_wspeed; p.uni_sin = (float)sin(wd); p.uni_cos = (float)cos(wd);
    p.thr = threshold; p.thr_inv = 1.0f / threshold; p.inv_em1 = (float)(1.0 / (exp(1.0) - 1.0));
    p.vec_ok = (cols % 4 == 0) && aligned16(dem) && aligned16(wspeed) && aligned16(wdirn) && aligned16(slope_deg) &&
               aligned16(aspect_deg) && aligned16(orograph) && aligned16(updraft);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int ntiles = p.tiles_r * p.tiles_c;

    const char* force = getenv("SSRS_STENCIL_PATH");
    bool want_tma = p.vec_ok && ntiles > 1;
    if (force && force[0] == 'p') want_tma = false;
    if (force && force[0] == 't') {
        SSRS_REQUIRE(p.vec_ok, "ssrs_updraft: SSRS_STENCIL_PATH=tma needs cols %% 4 == 0 and 16-byte aligned rasters");
        want_tma = true;
    }
    if (want_tma) {
        EncodeTiledFn enc = get_encode_fn();
        if (enc == nullptr) {
            set_error("ssrs_updraft: cuTensorMapEncodeTiled not available from the driver");
            return SSRS_ERR_CUDA;
        }
        CUtensorMap map;
        cuuint64_t gdim[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
        cuuint64_t gstride[1] = {(cuuint64_t)cols * sizeof(float)};
        cuuint32_t box[2] = {(cuuint32_t)SW, (cuuint32_t)SH};
        cuuint32_t estr[2] = {1, 1};
        CUresult r = enc(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(dem), gdim, gstride, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) {
            set_error("ssrs_updraft: cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
            return SSRS_ERR_CUDA;
        }
        // persistent grid: a multiple of the SM count (4 CTAs/SM: 37 KB smem, 256 threads, <= 64 registers each)
        int grid = sm_count() * 4;
        if (grid > ntiles) grid = ntiles;
        updraft_tma_kernel<true><<<grid, NTHREADS, 0, st>>>(p, map);
    } else if (p.vec_ok) {
        updraft_plain_kernel<true><<<ntiles, NTHREADS, 0, st>>>(p);
    } else {
        updraft_plain_kernel<false><<<ntiles, NTHREADS, 0, st>>>(p);
    }
    SSRS_CUDA_TRY(cudaGetLastError());
    return SSRS_OK;
}

extern "C" int ssrs_threshold(const float* in, float* out, int64_t n, float threshold, void* stream) {
    SSRS_REQUIRE(in != nullptr && out != nullptr, "ssrs_threshold: NULL raster");
    SSRS_REQUIRE(n >= 0 && threshold > 0.0f, "ssrs_threshold: bad size or threshold");
    if (n == 0) return SSRS_OK;
    int64_t blocks = cdiv(n, 256);
    int64_t cap = (int64_t)sm_count() * 16;
    if (blocks > cap) blocks = cap;
    threshold_kernel<<<(int)blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(
        in, out, n, threshold, 1.0f / threshold, (float)(1.0 / (exp(1.0) - 1.0)));
    SSRS_CUDA_TRY(cudaGetLastError());
    return SSRS_OK;
}

extern "C" int ssrs_orographic_updraft(const float* slope_deg, const float* aspect_deg, const float* wspeed, const float* wdirn,
                                       float uniform_wspeed, float uniform_wdirn_deg, float min_updraft, float* out, int64_t n,
                                       void* stream) {
    SSRS_REQUIRE(slope_deg && aspect_deg && out, "ssrs_orographic_updraft: NULL raster");
    SSRS_REQUIRE(n >= 0, "ssrs_orographic_updraft: negative size");
    if (n == 0) return SSRS_OK;
    int64_t blocks = cdiv(n, 256);
    const int64_t cap = (int64_t)sm_count() * 16;
    if (blocks > cap) blocks = cap;
    orographic_kernel<<<(int)blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(slope_deg, aspect_deg, wspeed, wdirn, uniform_wspeed,
                                                                                 uniform_wdirn_deg, min_updraft, out, n);
    SSRS_CUDA_TRY(cudaGetLastError());
    return SSRS_OK;
}
