// Patch I/O kernels either side of the sampler: tile extraction with per-patch normalisation, and the
// overlap-blend stitch.  Both are HBM-bound gather kernels (no tensor-core work): coalesced along image rows.
//
// Reference semantics
//   tile extract: Patch.py:80-84 (window order, supplied by the caller as origins), :201-203 (slicing), :41-49 +
//                 :192 (validity = every input channel finite [and an optional caller mask]), :51-62 + :228-229
//                 (masked z-score of HH, HV), :231-232 (incidence / 90, elevation / 1000), :236-239 (invalid -> 0,
//                 non-finite -> 0).
//   stitch:       not in the reference; definition in DESIGN.md: uniform weights,
//                 per-pixel gather in ascending patch index, fp32 sum then one fp32 division.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

#include "ptx_sm100.cuh"

namespace s1s2 {

constexpr int kExtractThreads = 1024;
constexpr int kStitchMaxC = 8;

// All four kernels are templated on G, the number of consecutive pixels of an image row one thread handles per step:
// G = 4 moves float4 / uchar4 (used when the row pitch, the window size, every window origin and the base pointers
// are multiples of 4 elements -- Patch.py's 256 / stride 32 geometry), G = 1 is the general scalar path.  Same
// arithmetic either way; only the association of the fp64 partial sums differs.
template <int G>
__device__ __forceinline__ void ld_f(const float* p, float (&o)[G]) {
    if constexpr (G == 4) {
        const float4 t = __ldg(reinterpret_cast<const float4*>(p));
        o[0] = t.x; o[1] = t.y; o[2] = t.z; o[3] = t.w;
    } else {
        o[0] = __ldg(p);
    }
}
template <int G>
__device__ __forceinline__ void st_f(float* p, const float (&v)[G]) {
    if constexpr (G == 4) *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
    else p[0] = v[0];
}
template <int G>
__device__ __forceinline__ void ld_b(const uint8_t* p, uint8_t (&o)[G]) {
    if constexpr (G == 4) {
        const uchar4 t = __ldg(reinterpret_cast<const uchar4*>(p));
        o[0] = t.x; o[1] = t.y; o[2] = t.z; o[3] = t.w;
    } else {
        o[0] = __ldg(p);
    }
}
template <int G>
__device__ __forceinline__ void st_b(uint8_t* p, const uint8_t (&v)[G]) {
    if constexpr (G == 4) *reinterpret_cast<uchar4*>(p) = make_uchar4(v[0], v[1], v[2], v[3]);
    else p[0] = v[0];
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// Sums NV doubles over the block (up to 32 warps); every thread receives the totals.  red: [32][NV].
template <int NV>
__device__ __forceinline__ void block_sum(double (&v)[NV], double (*red)[NV]) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    __syncthreads();                 // red may still be read from a previous call
#pragma unroll
    for (int i = 0; i < NV; ++i) {
        const double s = warp_sum(v[i]);
        if (lane == 0) red[warp][i] = s;
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < NV; ++i) v[i] = warp_sum(lane < nw ? red[lane][i] : 0.0);
}

// One CTA per window.  scene f32[4,SH,SW]; vmask u8[SH,SW] or nullptr; origins i32[N,2] (row, col);
// cond f32[N,4,ps,ps]; mask u8[N,ps,ps]; valid_ratio f32[N] or nullptr.
// Two coalesced passes over the window (the second one is served by L2): masked sum / sum of squares of HH and HV in
// fp64 (numpy uses fp32 pairwise sums; the two agree to ~1 ulp of the mean), then normalise + write.
template <int G>
__device__ __forceinline__ void tile_extract_body(const float* __restrict__ scene, const uint8_t* __restrict__ vmask, int SH,
                                                  int SW, int r0, int c0, int ps, float* __restrict__ cond,
                                                  uint8_t* __restrict__ mask, float* __restrict__ valid_ratio,
                                                  double (*red)[5]) {
    const int p = blockIdx.x;
    const size_t plane = static_cast<size_t>(SH) * SW;
    const int npix = ps * ps;
    const int gpr = ps / G, ngroups = ps * gpr;          // groups per window row, per window

    double a[5] = {0.0, 0.0, 0.0, 0.0, 0.0};             // sum HH, sum HH^2, sum HV, sum HV^2, count
    for (int i = threadIdx.x; i < ngroups; i += blockDim.x) {
        const int y = i / gpr, x = (i - y * gpr) * G;
        const size_t g = static_cast<size_t>(r0 + y) * SW + (c0 + x);
        float v0[G], v1[G], v2[G], v3[G];
        uint8_t vm[G];
        ld_f<G>(scene + g, v0);
        ld_f<G>(scene + plane + g, v1);
        ld_f<G>(scene + 2 * plane + g, v2);
        ld_f<G>(scene + 3 * plane + g, v3);
        if (vmask != nullptr) ld_b<G>(vmask + g, vm);
#pragma unroll
        for (int k = 0; k < G; ++k) {
            bool ok = isfinite(v0[k]) && isfinite(v1[k]) && isfinite(v2[k]) && isfinite(v3[k]);
            if (vmask != nullptr) ok = ok && vm[k] != 0;
            if (ok) {
                const double d0 = v0[k], d1 = v1[k];
                a[0] += d0;
                a[1] = fma(d0, d0, a[1]);
                a[2] += d1;
                a[3] = fma(d1, d1, a[3]);
                a[4] += 1.0;
            }
        }
    }
    block_sum<5>(a, red);
    const double cnt = a[4];
    const double m0 = cnt > 0.0 ? a[0] / cnt : 0.0, m1 = cnt > 0.0 ? a[2] / cnt : 0.0;
    float mu0 = static_cast<float>(m0), mu1 = static_cast<float>(m1);
    float sd0 = cnt > 0.0 ? static_cast<float>(sqrt(fmax(a[1] / cnt - m0 * m0, 0.0))) : 1.f;
    float sd1 = cnt > 0.0 ? static_cast<float>(sqrt(fmax(a[3] / cnt - m1 * m1, 0.0))) : 1.f;
    if (!isfinite(mu0)) mu0 = 0.f;
    if (!isfinite(mu1)) mu1 = 0.f;
    if (!isfinite(sd0) || static_cast<double>(sd0) < 1e-6) sd0 = 1.f;
    if (!isfinite(sd1) || static_cast<double>(sd1) < 1e-6) sd1 = 1.f;

    float* cp = cond + static_cast<size_t>(p) * 4 * npix;
    uint8_t* mp = mask + static_cast<size_t>(p) * npix;
    for (int i = threadIdx.x; i < ngroups; i += blockDim.x) {
        const int y = i / gpr, x = (i - y * gpr) * G;
        const size_t g = static_cast<size_t>(r0 + y) * SW + (c0 + x);
        float v0[G], v1[G], v2[G], v3[G];
        uint8_t vm[G], mo[G];
        ld_f<G>(scene + g, v0);
        ld_f<G>(scene + plane + g, v1);
        ld_f<G>(scene + 2 * plane + g, v2);
        ld_f<G>(scene + 3 * plane + g, v3);
        if (vmask != nullptr) ld_b<G>(vmask + g, vm);
        float o0[G], o1[G], o2[G], o3[G];
#pragma unroll
        for (int k = 0; k < G; ++k) {
            bool ok = isfinite(v0[k]) && isfinite(v1[k]) && isfinite(v2[k]) && isfinite(v3[k]);
            if (vmask != nullptr) ok = ok && vm[k] != 0;
            o0[k] = o1[k] = o2[k] = o3[k] = 0.f;
            if (ok) {
                o0[k] = __fdiv_rn(__fsub_rn(v0[k], mu0), sd0);
                o1[k] = __fdiv_rn(__fsub_rn(v1[k], mu1), sd1);
                o2[k] = __fdiv_rn(v2[k], 90.f);
                o3[k] = __fdiv_rn(v3[k], 1000.f);
                if (!isfinite(o0[k])) o0[k] = 0.f;
                if (!isfinite(o1[k])) o1[k] = 0.f;
            }
            mo[k] = ok ? 1 : 0;
        }
        const int o = y * ps + x;
        st_f<G>(cp + o, o0);
        st_f<G>(cp + npix + o, o1);
        st_f<G>(cp + 2 * npix + o, o2);
        st_f<G>(cp + 3 * npix + o, o3);
        st_b<G>(mp + o, mo);
    }
    if (valid_ratio != nullptr && threadIdx.x == 0) valid_ratio[p] = static_cast<float>(cnt / static_cast<double>(npix));
}
// allow_vec: the host found SW, ps and the base pointers 4-element aligned; the window's column origin decides per CTA.
__global__ void __launch_bounds__(kExtractThreads) tile_extract_kernel(const float* __restrict__ scene,
                                                                       const uint8_t* __restrict__ vmask, int SH, int SW,
                                                                       const int32_t* __restrict__ origins, int ps,
                                                                       float* __restrict__ cond, uint8_t* __restrict__ mask,
                                                                       float* __restrict__ valid_ratio, int allow_vec) {
    __shared__ double red[32][5];
    const int r0 = origins[2 * blockIdx.x], c0 = origins[2 * blockIdx.x + 1];
    if (r0 < 0 || c0 < 0 || r0 > SH - ps || c0 > SW - ps) {
        // a window that does not lie inside the scene (Patch.py's patch_iter never yields one): all-invalid patch
        const size_t npix = static_cast<size_t>(ps) * ps;
        float* cp = cond + static_cast<size_t>(blockIdx.x) * 4 * npix;
        uint8_t* mp = mask + static_cast<size_t>(blockIdx.x) * npix;
        for (size_t i = threadIdx.x; i < 4 * npix; i += blockDim.x) cp[i] = 0.f;
        for (size_t i = threadIdx.x; i < npix; i += blockDim.x) mp[i] = 0;
        if (valid_ratio != nullptr && threadIdx.x == 0) valid_ratio[blockIdx.x] = 0.f;
        return;
    }
    if (allow_vec && (c0 & 3) == 0) tile_extract_body<4>(scene, vmask, SH, SW, r0, c0, ps, cond, mask, valid_ratio, red);
    else tile_extract_body<1>(scene, vmask, SH, SW, r0, c0, ps, cond, mask, valid_ratio, red);
}

// ---------------------------------------------------------------------------------------------- quality filters
// Patch.py's four window tests on the target (Patch.py:205-224): valid ratio, all-band variance, dark fraction
// (:88-98) and Laplacian variance of band 3 with a symmetric window boundary (:100-114; scipy's convolve2d also
// multiplies the 3x3 corners by zero, so a non-finite corner voids the sample).  One CTA per window, ONE pass: every
// statistic is a ratio of fp64 sums (count, sum y, sum y^2 per band, dark count, Laplacian moments).
// stats[p][0..7] = valid_ratio, var[0..3], dark_fraction, laplacian_var, decision code (0 keep, 1 valid ratio, 2 flat,
// 3 dark, 4 no texture).  Validity = build_mask (:41-49): every input and target band finite and colloc > 0.
struct FilterThresholds {
    float valid_ratio, variance, dark_thr, dark_max_ratio, texture;
};
constexpr int kFilterThreads = 256;
constexpr int kFilterMaxCi = 8;

constexpr int kFilterVals = 13;   // count, sum y[4], sum y^2[4], dark, sum L, sum L^2, count L
template <int G>
__device__ __forceinline__ void tile_filter_body(const float* __restrict__ scene, int Ci, const float* __restrict__ target,
                                                 const uint8_t* __restrict__ colloc, int SH, int SW, int r0, int c0, int ps,
                                                 const FilterThresholds& th, float* __restrict__ stats,
                                                 double (*red)[kFilterVals]) {
    constexpr int NV = kFilterVals;
    const int p = blockIdx.x;
    const size_t plane = static_cast<size_t>(SH) * SW;
    const int npix = ps * ps;
    const int gpr = ps / G, ngroups = ps * gpr;
    const float* b8 = target + 3 * plane;
    double a[NV];
#pragma unroll
    for (int i = 0; i < NV; ++i) a[i] = 0.0;
    for (int i = threadIdx.x; i < ngroups; i += blockDim.x) {
        const int y = i / gpr, x = (i - y * gpr) * G;
        const size_t g = static_cast<size_t>(r0 + y) * SW + (c0 + x);
        bool ok[G];
#pragma unroll
        for (int k = 0; k < G; ++k) ok[k] = true;
        if (colloc != nullptr) {
            uint8_t cl[G];
            ld_b<G>(colloc + g, cl);
#pragma unroll
            for (int k = 0; k < G; ++k) ok[k] = cl[k] != 0;
        }
        for (int c = 0; c < Ci; ++c) {
            float v[G];
            ld_f<G>(scene + c * plane + g, v);
#pragma unroll
            for (int k = 0; k < G; ++k) ok[k] = ok[k] && isfinite(v[k]);
        }
        float yv[4][G];
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            ld_f<G>(target + c * plane + g, yv[c]);
#pragma unroll
            for (int k = 0; k < G; ++k) ok[k] = ok[k] && isfinite(yv[c][k]);
        }
        bool any = false;
#pragma unroll
        for (int k = 0; k < G; ++k) any = any || ok[k];
        if (!any) continue;
        // band-3 rows above / below and the columns left / right of the group, symmetric boundary (-1 -> 0, ps -> ps-1)
        const int ym = y > 0 ? y - 1 : 0, yp = y < ps - 1 ? y + 1 : ps - 1;
        const int xm = x > 0 ? x - 1 : 0, xp = x + G < ps ? x + G : ps - 1;
        float rm[G + 2], rc[G + 2], rp[G + 2];
        {
            const float* rowm = b8 + static_cast<size_t>(r0 + ym) * SW + c0;
            const float* rowc = b8 + static_cast<size_t>(r0 + y) * SW + c0;
            const float* rowp = b8 + static_cast<size_t>(r0 + yp) * SW + c0;
            float t[G];
            ld_f<G>(rowm + x, t);
#pragma unroll
            for (int k = 0; k < G; ++k) rm[k + 1] = t[k];
            ld_f<G>(rowp + x, t);
#pragma unroll
            for (int k = 0; k < G; ++k) rp[k + 1] = t[k];
#pragma unroll
            for (int k = 0; k < G; ++k) rc[k + 1] = yv[3][k];
            rm[0] = __ldg(rowm + xm); rm[G + 1] = __ldg(rowm + xp);
            rc[0] = __ldg(rowc + xm); rc[G + 1] = __ldg(rowc + xp);
            rp[0] = __ldg(rowp + xm); rp[G + 1] = __ldg(rowp + xp);
        }
#pragma unroll
        for (int k = 0; k < G; ++k) {
            if (!ok[k]) continue;
            const float y0 = yv[0][k], y1 = yv[1][k], y2 = yv[2][k], y3 = yv[3][k];
            a[0] += 1.0;
            const double d0 = y0, d1 = y1, d2 = y2, d3 = y3;
            a[1] += d0; a[2] += d1; a[3] += d2; a[4] += d3;
            a[5] = fma(d0, d0, a[5]); a[6] = fma(d1, d1, a[6]); a[7] = fma(d2, d2, a[7]); a[8] = fma(d3, d3, a[8]);
            const float vis = __fdiv_rn(__fadd_rn(__fadd_rn(y0, y1), y2), 3.0f);
            if (vis < th.dark_thr && y3 < th.dark_thr) a[9] += 1.0;
            const float c = rc[k + 1], n = rm[k + 1], s = rp[k + 1], w = rc[k], e = rc[k + 2];
            if (isfinite(c) && isfinite(n) && isfinite(s) && isfinite(w) && isfinite(e) && isfinite(rm[k]) && isfinite(rm[k + 2]) &&
                isfinite(rp[k]) && isfinite(rp[k + 2])) {
                // float32 like scipy's convolve2d on the reference's float32 arrays (Patch.py:104,112)
                const double L = static_cast<double>(__fsub_rn(__fadd_rn(__fadd_rn(n, s), __fadd_rn(w, e)), __fmul_rn(4.f, c)));
                a[10] += L; a[11] = fma(L, L, a[11]); a[12] += 1.0;
            }
        }
    }
    block_sum<NV>(a, red);
    if (threadIdx.x == 0) {
        const double cnt = a[0];
        const double dark = cnt > 0.0 ? a[9] / cnt : 1.0;
        const float nanv = __int_as_float(0x7fc00000);
        float v[4];
        for (int c = 0; c < 4; ++c) {
            const double m = a[1 + c] / cnt;
            v[c] = cnt > 0.0 ? static_cast<float>(fmax(a[5 + c] / cnt - m * m, 0.0)) : nanv;
        }
        float lv = 0.f;
        if (cnt > 0.0) lv = a[12] > 0.0 ? static_cast<float>(fmax(a[11] / a[12] - (a[10] / a[12]) * (a[10] / a[12]), 0.0)) : nanv;
        const float vr = static_cast<float>(cnt / npix);
        int code = 0;
        if (vr < th.valid_ratio) code = 1;
        else if (v[0] < th.variance && v[1] < th.variance && v[2] < th.variance && v[3] < th.variance) code = 2;
        else if (static_cast<float>(dark) > th.dark_max_ratio) code = 3;
        else if (lv < th.texture) code = 4;
        float* o = stats + static_cast<size_t>(p) * 8;
        o[0] = vr; o[1] = v[0]; o[2] = v[1]; o[3] = v[2]; o[4] = v[3];
        o[5] = static_cast<float>(dark); o[6] = lv; o[7] = static_cast<float>(code);
    }
}
__global__ void __launch_bounds__(kFilterThreads, 4) tile_filter_kernel(const float* __restrict__ scene, int Ci,
                                                                     const float* __restrict__ target,
                                                                     const uint8_t* __restrict__ colloc, int SH, int SW,
                                                                     const int32_t* __restrict__ origins, int ps,
                                                                     FilterThresholds th, float* __restrict__ stats, int allow_vec) {
    __shared__ double red[32][kFilterVals];
    const int r0 = origins[2 * blockIdx.x], c0 = origins[2 * blockIdx.x + 1];
    if (r0 < 0 || c0 < 0 || r0 > SH - ps || c0 > SW - ps) {           // window outside the scene: fails the valid-ratio test
        if (threadIdx.x < 8) stats[static_cast<size_t>(blockIdx.x) * 8 + threadIdx.x] = threadIdx.x == 7 ? 1.f : 0.f;
        return;
    }
    if (allow_vec && (c0 & 3) == 0) tile_filter_body<4>(scene, Ci, target, colloc, SH, SW, r0, c0, ps, th, stats, red);
    else tile_filter_body<1>(scene, Ci, target, colloc, SH, SW, r0, c0, ps, th, stats, red);
}

// ---------------------------------------------------------------------------------------------- evaluation metrics
// One CTA per patch, one pass over (pred, gt, mask): the reductions behind masked MAE / MSE / PSNR
// (Evaluation/DDIM_Multi-step.py:72-95), the global (non-windowed) ssim_simple (:97-101), SAM and ERGAS
// (Evaluation_Updated/Evaluation_Pure_Generation.py:229-254), finalised by thread 0 in double precision.
// A thread sums its G pixels x C channels of a step in fp32 and adds the step's partial sums to fp64 accumulators
// (the fp32 -> fp64 conversion rate, not HBM, bounded the all-fp64 version).
// out[p][0..7] = mae, mse, psnr, ssim_simple, sam, ergas, valid pixel count, 0;
// out[p][8 + c] = sum |pred - gt| and out[p][16 + c] = sum (pred - gt)^2 of channel c over the patch's valid pixels: the
// per-batch sums of the dataset-level, pixel-weighted aggregation (Evaluation/Limitation_Test.py:118-133).
constexpr int kMetricsMaxC = 8;
constexpr int kMetricsThreads = 256;
constexpr int kMetricsOut = 8 + 2 * kMetricsMaxC;

// Thread 0's finalisation of the block-reduced sums r[0 .. 3*CMAX+7) into one output row (shared by both kernels).
template <int CMAX>
__device__ __forceinline__ void metrics_finalize(const double* r, int C, int HW, double* o) {
    constexpr int kG = 3 * CMAX;
    const double W = r[kG + 5];
    double sabs = 0.0, ssq = 0.0, eg = 0.0;
    for (int c = 0; c < C; ++c) {
        sabs += r[3 * c];
        ssq += r[3 * c + 1];
        const double rmse = sqrt(fmax(r[3 * c + 1] / (W + 1e-8), 0.0));
        const double q = rmse / (r[3 * c + 2] / HW + 1e-8);
        eg += q * q;
    }
    const double mae = sabs / (W * C + 1e-8), mse = ssq / (W * C + 1e-8);
    const double n = static_cast<double>(C) * HW;
    double sum_b = 0.0;                              // global sum of gt = sum of the per-channel sums
    for (int c = 0; c < C; ++c) sum_b += r[3 * c + 2];
    const double mx = r[kG] / n, my = sum_b / n;
    const double vx = (r[kG + 2] - n * mx * mx) / (n - 1.0), vy = (r[kG + 3] - n * my * my) / (n - 1.0);
    const double cxy = r[kG + 4] / n - mx * my;
    const double C1 = 0.01 * 0.01, C2 = 0.03 * 0.03;
    o[0] = mae;
    o[1] = mse;
    o[2] = mse <= 1e-12 ? 99.0 : 10.0 * log10(1.0 / mse);
    o[3] = ((2 * mx * my + C1) * (2 * cxy + C2)) / ((mx * mx + my * my + C1) * (vx + vy + C2) + 1e-8);
    o[4] = r[kG + 6] / W;                            // NaN for an empty mask, like torch's mean of nothing
    o[5] = 100.0 * sqrt(eg / C) * 4.0;
    o[6] = W;
    o[7] = 0.0;
    for (int c = 0; c < kMetricsMaxC; ++c) {
        o[8 + c] = c < C ? r[3 * c] : 0.0;
        o[8 + kMetricsMaxC + c] = c < C ? r[3 * c + 1] : 0.0;
    }
}

template <int G, int CMAX>
__global__ void __launch_bounds__(kMetricsThreads, CMAX <= 4 ? 3 : 1) patch_metrics_kernel(const float* __restrict__ pred,
                                                                        const float* __restrict__ gt,
                                                                        const uint8_t* __restrict__ mask, int C, int HW,
                                                                        double* __restrict__ out) {
    constexpr int kG = 3 * CMAX;                         // per channel: sum|d|, sum d^2, sum gt ; then 5 global + count + sam
    constexpr int kVals = kG + 7;
    __shared__ double red[32][kVals];
    const int p = blockIdx.x;
    const float* pp = pred + static_cast<size_t>(p) * C * HW;
    const float* gp = gt + static_cast<size_t>(p) * C * HW;
    const uint8_t* mp = mask != nullptr ? mask + static_cast<size_t>(p) * HW : nullptr;
    double acc[kVals];
#pragma unroll
    for (int i = 0; i < kVals; ++i) acc[i] = 0.0;
    const int ngroups = HW / G;
    constexpr int kFlush = G == 4 ? 4 : 1;               // steps (of G pixels) summed in fp32 before the fp64 add
    for (int ib = threadIdx.x; ib < ngroups; ib += kFlush * blockDim.x) {
        float part[kVals];
#pragma unroll
        for (int j = 0; j < kVals; ++j) part[j] = 0.f;
#pragma unroll
      for (int f = 0; f < kFlush; ++f) {
        const int i = ib + f * blockDim.x;
        if (i >= ngroups) break;
        const int i0 = i * G;
        bool w[G];
#pragma unroll
        for (int k = 0; k < G; ++k) w[k] = true;
        if (mp != nullptr) {
            uint8_t m[G];
            ld_b<G>(mp + i0, m);
#pragma unroll
            for (int k = 0; k < G; ++k) w[k] = m[k] != 0;
        }
        // per pixel: sum_c a, a.b, |a|^2, |b|^2 over the channels (they feed both the global ssim_simple sums and SAM)
        float sa[G], dot[G], np2[G], ng2[G];
#pragma unroll
        for (int k = 0; k < G; ++k) sa[k] = dot[k] = np2[k] = ng2[k] = 0.f;
#pragma unroll
        for (int c = 0; c < CMAX; ++c) {
            if (c < C) {
                float av[G], bv[G];
                ld_f<G>(pp + static_cast<size_t>(c) * HW + i0, av);
                ld_f<G>(gp + static_cast<size_t>(c) * HW + i0, bv);
#pragma unroll
                for (int k = 0; k < G; ++k) {
                    const float a = av[k], b = bv[k];
                    const float d = a - b;
                    if (w[k]) {
                        part[3 * c] += fabsf(d);
                        part[3 * c + 1] = fmaf(d, d, part[3 * c + 1]);
                    }
                    part[3 * c + 2] += b;
                    sa[k] += a;
                    dot[k] = fmaf(a, b, dot[k]);
                    np2[k] = fmaf(a, a, np2[k]);
                    ng2[k] = fmaf(b, b, ng2[k]);
                }
            }
        }
#pragma unroll
        for (int k = 0; k < G; ++k) {
            part[kG + 0] += sa[k];
            part[kG + 2] += np2[k];
            part[kG + 3] += ng2[k];
            part[kG + 4] += dot[k];
            if (w[k]) {
                part[kG + 5] += 1.f;
                const float cosv = dot[k] / (fmaxf(sqrtf(np2[k]), 1e-8f) * fmaxf(sqrtf(ng2[k]), 1e-8f));
                part[kG + 6] += acosf(fminf(fmaxf(cosv, -1.f), 1.f));
            }
        }
      }
#pragma unroll
        for (int j = 0; j < kVals; ++j) acc[j] += static_cast<double>(part[j]);
    }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int i = 0; i < kVals; ++i) {
        const double v = warp_sum(acc[i]);
        if (lane == 0) red[warp][i] = v;
    }
    __syncthreads();
    if (warp == 0) {
#pragma unroll
        for (int i = 0; i < kVals; ++i) {
            const double v = warp_sum(lane < (blockDim.x >> 5) ? red[lane][i] : 0.0);
            if (lane == 0) red[0][i] = v;
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) metrics_finalize<CMAX>(red[0], C, HW, out + static_cast<size_t>(p) * kMetricsOut);
}

// Streaming version for the usual geometry (C <= 4, HW a multiple of kMsChunk, 16-byte aligned planes): the planes of a
// patch are contiguous, so one elected thread streams them through a ring of shared-memory stages with 1-D bulk copies
// (cp.async.bulk + mbarrier transaction counts: no tensor map, loads fully asynchronous and as deep as the ring), and 16
// consumer warps reduce from shared memory.  The register-path kernel above is bound by load latency at 24 warps per SM
// (ncu: issue 42 %, DRAM 41 %); here memory and math overlap by construction.  Same arithmetic, same output rows.
constexpr int kMsChunk = 2048;                      // pixels per stage
constexpr int kMsStages = 3;                        // 3 x 66 KB of the 227 KB
constexpr int kMsConsumers = 512;                   // 4 pixels (one float4 per plane) per consumer thread per stage
constexpr int kMsThreads = kMsConsumers + 32;       // + one producer warp
constexpr int kMsC = 4;
constexpr int kMsStageBytes = 2 * kMsC * kMsChunk * 4 + kMsChunk;      // pred + gt planes, mask bytes
constexpr int kMsSmemBytes = kMsStages * kMsStageBytes + 64;

__global__ void __launch_bounds__(kMsThreads, 1) patch_metrics_stream_kernel(const float* __restrict__ pred,
                                                                             const float* __restrict__ gt,
                                                                             const uint8_t* __restrict__ mask, int C, int HW,
                                                                             double* __restrict__ out) {
    constexpr int CMAX = kMsC;
    constexpr int kG = 3 * CMAX;
    constexpr int kVals = kG + 7;
    extern __shared__ __align__(128) uint8_t ms_smem[];
    __shared__ double red[32][kVals];
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(ms_smem + kMsStages * kMsStageBytes);
    uint64_t* empty_bar = full_bar + kMsStages;
    const int p = blockIdx.x;
    const float* pp = pred + static_cast<size_t>(p) * C * HW;
    const float* gp = gt + static_cast<size_t>(p) * C * HW;
    const uint8_t* mp = mask != nullptr ? mask + static_cast<size_t>(p) * HW : nullptr;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int nchunks = HW / kMsChunk;
    if (threadIdx.x == 0) {
        for (int s = 0; s < kMsStages; ++s) {
            mbar_init(&full_bar[s], 1);
            mbar_init(&empty_bar[s], kMsConsumers / 32);       // one arrive per consumer warp
        }
        mbar_fence_init();
    }
    __syncthreads();
    double acc[kVals];
#pragma unroll
    for (int i = 0; i < kVals; ++i) acc[i] = 0.0;
    if (warp == kMsConsumers / 32) {
        // ---------------------------------------------------------------- producer warp
        int s = 0;
        uint32_t ph = 0;
        const uint32_t bytes = static_cast<uint32_t>(2 * C * kMsChunk * 4 + (mp != nullptr ? kMsChunk : 0));
        for (int ch = 0; ch < nchunks; ++ch) {
            mbar_wait(&empty_bar[s], ph ^ 1);
            if (lane == 0) {
                uint8_t* st = ms_smem + s * kMsStageBytes;
                mbar_expect_tx(&full_bar[s], bytes);
                for (int c = 0; c < C; ++c) {
                    bulk_load_1d(st + c * kMsChunk * 4, pp + static_cast<size_t>(c) * HW + ch * kMsChunk, kMsChunk * 4, &full_bar[s]);
                    bulk_load_1d(st + (CMAX + c) * kMsChunk * 4, gp + static_cast<size_t>(c) * HW + ch * kMsChunk, kMsChunk * 4,
                                 &full_bar[s]);
                }
                if (mp != nullptr) bulk_load_1d(st + 2 * CMAX * kMsChunk * 4, mp + ch * kMsChunk, kMsChunk, &full_bar[s]);
            }
            __syncwarp();
            if (++s == kMsStages) { s = 0; ph ^= 1; }
        }
    } else {
        // ---------------------------------------------------------------- consumers: 4 pixels per thread per stage
        int s = 0;
        uint32_t ph = 0;
        float part[kVals];
#pragma unroll
        for (int j = 0; j < kVals; ++j) part[j] = 0.f;
        for (int ch = 0; ch < nchunks; ++ch) {
            mbar_wait(&full_bar[s], ph);
            const uint8_t* st = ms_smem + s * kMsStageBytes;
            bool w[4] = {true, true, true, true};
            if (mp != nullptr) {
                const uchar4 m = *reinterpret_cast<const uchar4*>(st + 2 * CMAX * kMsChunk * 4 + threadIdx.x * 4);
                w[0] = m.x != 0; w[1] = m.y != 0; w[2] = m.z != 0; w[3] = m.w != 0;
            }
            float sa[4] = {0.f, 0.f, 0.f, 0.f}, dot[4] = {0.f, 0.f, 0.f, 0.f}, np2[4] = {0.f, 0.f, 0.f, 0.f}, ng2[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
            for (int c = 0; c < CMAX; ++c) {
                if (c < C) {
                    const float4 a4 = *reinterpret_cast<const float4*>(st + c * kMsChunk * 4 + threadIdx.x * 16);
                    const float4 b4 = *reinterpret_cast<const float4*>(st + (CMAX + c) * kMsChunk * 4 + threadIdx.x * 16);
                    const float av[4] = {a4.x, a4.y, a4.z, a4.w}, bv[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        const float a = av[k], b = bv[k];
                        const float d = a - b;
                        if (w[k]) {
                            part[3 * c] += fabsf(d);
                            part[3 * c + 1] = fmaf(d, d, part[3 * c + 1]);
                        }
                        part[3 * c + 2] += b;
                        sa[k] += a;
                        dot[k] = fmaf(a, b, dot[k]);
                        np2[k] = fmaf(a, a, np2[k]);
                        ng2[k] = fmaf(b, b, ng2[k]);
                    }
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&empty_bar[s]);           // this warp's reads of the stage are in registers
            if (++s == kMsStages) { s = 0; ph ^= 1; }
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                part[kG + 0] += sa[k];
                part[kG + 2] += np2[k];
                part[kG + 3] += ng2[k];
                part[kG + 4] += dot[k];
                if (w[k]) {
                    part[kG + 5] += 1.f;
                    const float cosv = dot[k] / (fmaxf(sqrtf(np2[k]), 1e-8f) * fmaxf(sqrtf(ng2[k]), 1e-8f));
                    part[kG + 6] += acosf(fminf(fmaxf(cosv, -1.f), 1.f));
                }
            }
            if ((ch & 3) == 3 || ch == nchunks - 1) {            // fold the fp32 partial sums (<= 16 pixels) into fp64
#pragma unroll
                for (int j = 0; j < kVals; ++j) { acc[j] += static_cast<double>(part[j]); part[j] = 0.f; }
            }
        }
    }
    // block reduction (the producer warp contributes zeros) and finalisation: identical to patch_metrics_kernel
#pragma unroll
    for (int i = 0; i < kVals; ++i) {
        const double v = warp_sum(acc[i]);
        if (lane == 0) red[warp][i] = v;
    }
    __syncthreads();
    if (warp == 0) {
#pragma unroll
        for (int i = 0; i < kVals; ++i) {
            const double v = warp_sum(lane < (kMsThreads >> 5) ? red[lane][i] : 0.0);
            if (lane == 0) red[0][i] = v;
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) metrics_finalize<CMAX>(red[0], C, HW, out + static_cast<size_t>(p) * kMetricsOut);
}

// grid_map[(row/stride) * ncols + col/stride] = patch index (entries stay -1 where no patch was kept).
__global__ void stitch_map_kernel(const int32_t* __restrict__ origins, int N, int stride, int nrows, int ncols,
                                  int32_t* __restrict__ grid_map) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= N) return;
    const int r = origins[2 * p], c = origins[2 * p + 1];
    if (r < 0 || c < 0 || r % stride != 0 || c % stride != 0) return;
    const int i = r / stride, j = c / stride;
    if (i < nrows && j < ncols) grid_map[i * ncols + j] = p;
}

// One thread per G consecutive canvas pixels of a row (they share their covering patches when stride, ps and x are
// multiples of G); covering patches visited in ascending (row, col) = ascending patch index, fp32 adds in that order, one
// fp32 division: no atomics, bit-reproducible, identical for G = 1 and G = 4.
// win == nullptr: uniform weights (sum / count).  win = f32[ps]: separable window, weight of a patch pixel =
// win[ly] * win[lx] (e.g. Hann): sum of weight * pred in the same order / sum of weights.
// blockIdx.y strides over the canvas rows (gridDim.y is capped at 65535).
template <int G, int CMAX>
__global__ void __launch_bounds__(128, 8) stitch_gather_kernel(const float* __restrict__ preds,
                                                            const int32_t* __restrict__ grid_map, int C, int ps, int stride,
                                                            int nrows, int ncols, int SH, int SW, float* __restrict__ canvas,
                                                            uint8_t* __restrict__ cover, const float* __restrict__ win) {
    const int x = (blockIdx.x * blockDim.x + threadIdx.x) * G;
    if (x >= SW) return;
    for (int y = blockIdx.y; y < SH; y += gridDim.y) {
        const int i_lo = y >= ps ? (y - ps) / stride + 1 : 0;
        const int i_hi = min(nrows - 1, y / stride);
        const int j_lo = x >= ps ? (x - ps) / stride + 1 : 0;
        const int j_hi = min(ncols - 1, x / stride);
        float acc[CMAX][G];
#pragma unroll
        for (int c = 0; c < CMAX; ++c)
#pragma unroll
            for (int k = 0; k < G; ++k) acc[c][k] = 0.f;
        float cnt[G];
#pragma unroll
        for (int k = 0; k < G; ++k) cnt[k] = 0.f;
        const size_t pp = static_cast<size_t>(ps) * ps;
        for (int i = i_lo; i <= i_hi; ++i) {
            const int ly = y - i * stride;
#pragma unroll 4
            for (int j = j_lo; j <= j_hi; ++j) {
                const int p = __ldg(grid_map + i * ncols + j);
                if (p < 0) continue;
                const int lx = x - j * stride;
                const float* src = preds + static_cast<size_t>(p) * C * pp + static_cast<size_t>(ly) * ps + lx;
                if (win == nullptr) {
#pragma unroll
                    for (int c = 0; c < CMAX; ++c) {
                        if (c < C) {
                            float v[G];
                            ld_f<G>(src + c * pp, v);
#pragma unroll
                            for (int k = 0; k < G; ++k) acc[c][k] = __fadd_rn(acc[c][k], v[k]);
                        }
                    }
#pragma unroll
                    for (int k = 0; k < G; ++k) cnt[k] += 1.f;
                } else {
                    const float wy = __ldg(win + ly);
                    float wgt[G];
#pragma unroll
                    for (int k = 0; k < G; ++k) wgt[k] = __fmul_rn(wy, __ldg(win + lx + k));
#pragma unroll
                    for (int c = 0; c < CMAX; ++c) {
                        if (c < C) {
                            float v[G];
                            ld_f<G>(src + c * pp, v);
#pragma unroll
                            for (int k = 0; k < G; ++k) acc[c][k] = __fadd_rn(acc[c][k], __fmul_rn(wgt[k], v[k]));
                        }
                    }
#pragma unroll
                    for (int k = 0; k < G; ++k) cnt[k] = __fadd_rn(cnt[k], wgt[k]);
                }
            }
        }
        const size_t plane = static_cast<size_t>(SH) * SW;
        const size_t o = static_cast<size_t>(y) * SW + x;
#pragma unroll
        for (int c = 0; c < CMAX; ++c) {
            if (c < C) {
                float v[G];
#pragma unroll
                for (int k = 0; k < G; ++k) v[k] = cnt[k] > 0.f ? __fdiv_rn(acc[c][k], cnt[k]) : 0.f;
                st_f<G>(canvas + c * plane + o, v);
            }
        }
        uint8_t cv[G];
#pragma unroll
        for (int k = 0; k < G; ++k) cv[k] = cnt[k] > 0.f ? 1 : 0;
        st_b<G>(cover + o, cv);
    }
}

// ---------------------------------------------------------------------------------------------- debug / noise helpers
// Number of fp16 elements of an NHWC activation view that sit at the saturation value of the epilogues' cvt.rn.satfinite
// (|v| = 65504) or are not finite.  Measurement aid: a clamp in a conv epilogue leaves exactly this footprint.
__global__ void saturation_count_kernel(const __half* __restrict__ src, int cpitch, int C, size_t npix,
                                        unsigned long long* __restrict__ out) {
    unsigned long long n = 0;
    const size_t total = npix * C;
    for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < total;
         i += static_cast<size_t>(gridDim.x) * blockDim.x) {
        const size_t px = i / C;
        const int c = static_cast<int>(i - px * C);
        const uint16_t bits = __half_as_ushort(src[px * cpitch + c]) & 0x7FFFu;
        n += bits >= 0x7BFFu ? 1u : 0u;
    }
    n = __reduce_add_sync(0xffffffffu, static_cast<unsigned>(n));
    if ((threadIdx.x & 31) == 0 && n != 0) atomicAdd(out, n);
}

}  // namespace s1s2
