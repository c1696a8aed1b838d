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

namespace s1s2 {

constexpr int kExtractThreads = 1024;
constexpr int kStitchMaxC = 8;

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// Sums three doubles over the block; every thread receives the totals.
__device__ __forceinline__ void block_sum3(double& a, double& b, double& c, double* scratch /* [3*32] */) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    a = warp_sum(a);
    b = warp_sum(b);
    c = warp_sum(c);
    __syncthreads();                 // scratch may still be read from a previous call
    if (lane == 0) {
        scratch[warp] = a;
        scratch[32 + warp] = b;
        scratch[64 + warp] = c;
    }
    __syncthreads();
    a = lane < nw ? scratch[lane] : 0.0;
    b = lane < nw ? scratch[32 + lane] : 0.0;
    c = lane < nw ? scratch[64 + lane] : 0.0;
    a = warp_sum(a);
    b = warp_sum(b);
    c = warp_sum(c);
}

// One CTA per window.  scene f32[4,SH,SW]; vmask u8[SH,SW] or nullptr; origins i32[N,2] (row, col);
// cond f32[N,4,ps,ps]; mask u8[N,ps,ps]; valid_ratio f32[N] or nullptr.
// Statistics are accumulated in fp64 (numpy uses fp32 pairwise sums; the two agree to ~1 ulp of the mean).
__global__ void __launch_bounds__(kExtractThreads) tile_extract_kernel(const float* __restrict__ scene,
                                                                       const uint8_t* __restrict__ vmask, int SH, int SW,
                                                                       const int32_t* __restrict__ origins, int ps,
                                                                       float* __restrict__ cond, uint8_t* __restrict__ mask,
                                                                       float* __restrict__ valid_ratio) {
    __shared__ double scratch[96];
    const int p = blockIdx.x;
    const int r0 = origins[2 * p], c0 = origins[2 * p + 1];
    const size_t plane = static_cast<size_t>(SH) * SW;
    const int npix = ps * ps;

    auto valid_at = [&](size_t g, float v0, float v1, float v2, float v3) {
        bool ok = isfinite(v0) && isfinite(v1) && isfinite(v2) && isfinite(v3);
        if (vmask != nullptr) ok = ok && vmask[g] != 0;
        return ok;
    };

    double s0 = 0.0, s1 = 0.0, cnt = 0.0;
    for (int i = threadIdx.x; i < npix; i += blockDim.x) {
        const int y = i / ps, x = i - y * ps;
        const size_t g = static_cast<size_t>(r0 + y) * SW + (c0 + x);
        const float v0 = scene[g], v1 = scene[plane + g], v2 = scene[2 * plane + g], v3 = scene[3 * plane + g];
        if (valid_at(g, v0, v1, v2, v3)) {
            s0 += v0;
            s1 += v1;
            cnt += 1.0;
        }
    }
    block_sum3(s0, s1, cnt, scratch);
    const double m0 = cnt > 0.0 ? s0 / cnt : 0.0, m1 = cnt > 0.0 ? s1 / cnt : 0.0;

    double q0 = 0.0, q1 = 0.0, dummy = 0.0;
    for (int i = threadIdx.x; i < npix; i += blockDim.x) {
        const int y = i / ps, x = i - y * ps;
        const size_t g = static_cast<size_t>(r0 + y) * SW + (c0 + x);
        const float v0 = scene[g], v1 = scene[plane + g], v2 = scene[2 * plane + g], v3 = scene[3 * plane + g];
        if (valid_at(g, v0, v1, v2, v3)) {
            const double d0 = v0 - m0, d1 = v1 - m1;
            q0 += d0 * d0;
            q1 += d1 * d1;
        }
    }
    block_sum3(q0, q1, dummy, scratch);
    float mu0 = static_cast<float>(m0), mu1 = static_cast<float>(m1);
    float sd0 = cnt > 0.0 ? static_cast<float>(sqrt(q0 / cnt)) : 1.f;
    float sd1 = cnt > 0.0 ? static_cast<float>(sqrt(q1 / cnt)) : 1.f;
    if (!isfinite(mu0)) mu0 = 0.f;
    if (!isfinite(mu1)) mu1 = 0.f;
    if (!isfinite(sd0) || static_cast<double>(sd0) < 1e-6) sd0 = 1.f;
    if (!isfinite(sd1) || static_cast<double>(sd1) < 1e-6) sd1 = 1.f;

    float* cp = cond + static_cast<size_t>(p) * 4 * npix;
    uint8_t* mp = mask + static_cast<size_t>(p) * npix;
    for (int i = threadIdx.x; i < npix; i += blockDim.x) {
        const int y = i / ps, x = i - y * ps;
        const size_t g = static_cast<size_t>(r0 + y) * SW + (c0 + x);
        const float v0 = scene[g], v1 = scene[plane + g], v2 = scene[2 * plane + g], v3 = scene[3 * plane + g];
        const bool ok = valid_at(g, v0, v1, v2, v3);
        float o0 = 0.f, o1 = 0.f, o2 = 0.f, o3 = 0.f;
        if (ok) {
            o0 = __fdiv_rn(__fsub_rn(v0, mu0), sd0);
            o1 = __fdiv_rn(__fsub_rn(v1, mu1), sd1);
            o2 = __fdiv_rn(v2, 90.f);
            o3 = __fdiv_rn(v3, 1000.f);
            if (!isfinite(o0)) o0 = 0.f;
            if (!isfinite(o1)) o1 = 0.f;
        }
        cp[i] = o0;
        cp[npix + i] = o1;
        cp[2 * npix + i] = o2;
        cp[3 * npix + i] = o3;
        mp[i] = ok ? 1 : 0;
    }
    if (valid_ratio != nullptr && threadIdx.x == 0) valid_ratio[p] = static_cast<float>(cnt / static_cast<double>(npix));
}

// ---------------------------------------------------------------------------------------------- quality filters
// Patch.py's four window tests on the target (Patch.py:205-224): valid ratio, all-band variance, dark fraction
// (:88-98) and Laplacian variance of band 3 with a symmetric window boundary (:100-114; scipy's convolve2d also
// multiplies the 3x3 corners by zero, so a non-finite corner voids the sample).  One CTA per window, two passes.
// stats[p][0..7] = valid_ratio, var[0..3], dark_fraction, laplacian_var, decision code (0 keep, 1 valid ratio, 2 flat,
// 3 dark, 4 no texture).  Validity = build_mask (:41-49): every input and target band finite and colloc > 0.
struct FilterThresholds {
    float valid_ratio, variance, dark_thr, dark_max_ratio, texture;
};
constexpr int kFilterThreads = 256;

__device__ __forceinline__ void block_sum_n(double* v, int n, double (*red)[12]) {   // n <= 12; result in v on all threads
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    __syncthreads();
    for (int i = 0; i < n; ++i) {
        const double s = warp_sum(v[i]);
        if (lane == 0) red[warp][i] = s;
    }
    __syncthreads();
    for (int i = 0; i < n; ++i) v[i] = warp_sum(lane < nw ? red[lane][i] : 0.0);
}

__global__ void __launch_bounds__(kFilterThreads) tile_filter_kernel(const float* __restrict__ scene, int Ci,
                                                                     const float* __restrict__ target,
                                                                     const uint8_t* __restrict__ colloc, int SH, int SW,
                                                                     const int32_t* __restrict__ origins, int ps,
                                                                     FilterThresholds th, float* __restrict__ stats) {
    __shared__ double red[kFilterThreads / 32][12];
    const int p = blockIdx.x;
    const int r0 = origins[2 * p], c0 = origins[2 * p + 1];
    const size_t plane = static_cast<size_t>(SH) * SW;
    const int npix = ps * ps;
    auto valid_at = [&](size_t g) {
        bool ok = colloc == nullptr || colloc[g] != 0;
        for (int c = 0; c < Ci && ok; ++c) ok = isfinite(scene[c * plane + g]);
        for (int c = 0; c < 4 && ok; ++c) ok = isfinite(target[c * plane + g]);
        return ok;
    };
    // pass A: count, band sums, dark pixels
    double a[12];
    for (int i = 0; i < 12; ++i) a[i] = 0.0;
    for (int i = threadIdx.x; i < npix; i += blockDim.x) {
        const int y = i / ps, x = i - y * ps;
        const size_t g = static_cast<size_t>(r0 + y) * SW + (c0 + x);
        if (!valid_at(g)) continue;
        const float y0 = target[g], y1 = target[plane + g], y2 = target[2 * plane + g], y3 = target[3 * plane + g];
        a[0] += 1.0;
        a[1] += y0; a[2] += y1; a[3] += y2; a[4] += y3;
        const float vis = __fdiv_rn(__fadd_rn(__fadd_rn(y0, y1), y2), 3.0f);
        if (vis < th.dark_thr && y3 < th.dark_thr) a[5] += 1.0;
    }
    block_sum_n(a, 6, red);
    const double cnt = a[0];
    const double m0 = a[1] / cnt, m1 = a[2] / cnt, m2 = a[3] / cnt, m3 = a[4] / cnt;
    const double dark = cnt > 0.0 ? a[5] / cnt : 1.0;
    // pass B: squared deviations, Laplacian moments
    double b[12];
    for (int i = 0; i < 12; ++i) b[i] = 0.0;
    const float* b8 = target + 3 * plane;
    for (int i = threadIdx.x; i < npix; i += blockDim.x) {
        const int y = i / ps, x = i - y * ps;
        const size_t g = static_cast<size_t>(r0 + y) * SW + (c0 + x);
        if (!valid_at(g)) continue;
        const double d0 = target[g] - m0, d1 = target[plane + g] - m1, d2 = target[2 * plane + g] - m2, d3 = target[3 * plane + g] - m3;
        b[0] += d0 * d0; b[1] += d1 * d1; b[2] += d2 * d2; b[3] += d3 * d3;
        // symmetric boundary: index -1 -> 0, ps -> ps-1
        const int ym = y > 0 ? y - 1 : 0, yp = y < ps - 1 ? y + 1 : ps - 1;
        const int xm = x > 0 ? x - 1 : 0, xp = x < ps - 1 ? x + 1 : ps - 1;
        auto at = [&](int yy, int xx) { return b8[static_cast<size_t>(r0 + yy) * SW + (c0 + xx)]; };
        const float c = at(y, x), n = at(ym, x), s = at(yp, x), w = at(y, xm), e = at(y, xp);
        const float k0 = at(ym, xm), k1 = at(ym, xp), k2 = at(yp, xm), k3 = at(yp, xp);
        if (isfinite(c) && isfinite(n) && isfinite(s) && isfinite(w) && isfinite(e) && isfinite(k0) && isfinite(k1) &&
            isfinite(k2) && isfinite(k3)) {
            const double L = static_cast<double>(n) + s + w + e - 4.0 * c;
            b[4] += L; b[5] += L * L; b[6] += 1.0;
        }
    }
    block_sum_n(b, 7, red);
    if (threadIdx.x == 0) {
        const float nanv = __int_as_float(0x7fc00000);
        float v[4];
        for (int c = 0; c < 4; ++c) v[c] = cnt > 0.0 ? static_cast<float>(b[c] / cnt) : nanv;
        float lv = 0.f;
        if (cnt > 0.0) lv = b[6] > 0.0 ? static_cast<float>(fmax(b[5] / b[6] - (b[4] / b[6]) * (b[4] / b[6]), 0.0)) : nanv;
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

// ---------------------------------------------------------------------------------------------- evaluation metrics
// One CTA per patch, one pass over (pred, gt, mask): the reductions behind masked MAE / MSE / PSNR
// (Evaluation/DDIM_Multi-step.py:72-95), the global (non-windowed) ssim_simple (:97-101), SAM and ERGAS
// (Evaluation_Updated/Evaluation_Pure_Generation.py:229-254), finalised by thread 0 in double precision.
// out[p][0..7] = mae, mse, psnr, ssim_simple, sam, ergas, valid pixel count, 0.
constexpr int kMetricsMaxC = 8;
constexpr int kMetricsThreads = 256;      // 31 double accumulators per thread: keep the register budget wide
constexpr int kMetricsOut = 8;

__global__ void __launch_bounds__(kMetricsThreads) patch_metrics_kernel(const float* __restrict__ pred,
                                                                        const float* __restrict__ gt,
                                                                        const uint8_t* __restrict__ mask, int C, int HW,
                                                                        double* __restrict__ out) {
    constexpr int kVals = 3 * kMetricsMaxC + 7;          // per channel: sum|d|, sum d^2, sum gt ; then 5 global + count + sam
    __shared__ double red[32][kVals];
    const int p = blockIdx.x;
    const float* pp = pred + static_cast<size_t>(p) * C * HW;
    const float* gp = gt + static_cast<size_t>(p) * C * HW;
    const uint8_t* mp = mask != nullptr ? mask + static_cast<size_t>(p) * HW : nullptr;
    double acc[kVals];
#pragma unroll
    for (int i = 0; i < kVals; ++i) acc[i] = 0.0;
    for (int i = threadIdx.x; i < HW; i += blockDim.x) {
        const bool w = mp == nullptr || mp[i] != 0;
        float dot = 0.f, np2 = 0.f, ng2 = 0.f;
#pragma unroll
        for (int c = 0; c < kMetricsMaxC; ++c) {
            if (c < C) {
                const float a = pp[static_cast<size_t>(c) * HW + i], b = gp[static_cast<size_t>(c) * HW + i];
                const float d = a - b;
                if (w) {
                    acc[3 * c] += fabsf(d);
                    acc[3 * c + 1] += static_cast<double>(d) * d;
                }
                acc[3 * c + 2] += b;
                acc[3 * kMetricsMaxC + 0] += a;
                acc[3 * kMetricsMaxC + 1] += b;
                acc[3 * kMetricsMaxC + 2] += static_cast<double>(a) * a;
                acc[3 * kMetricsMaxC + 3] += static_cast<double>(b) * b;
                acc[3 * kMetricsMaxC + 4] += static_cast<double>(a) * b;
                dot = fmaf(a, b, dot);
                np2 = fmaf(a, a, np2);
                ng2 = fmaf(b, b, ng2);
            }
        }
        if (w) {
            acc[3 * kMetricsMaxC + 5] += 1.0;
            const float cosv = dot / (fmaxf(sqrtf(np2), 1e-8f) * fmaxf(sqrtf(ng2), 1e-8f));
            acc[3 * kMetricsMaxC + 6] += acosf(fminf(fmaxf(cosv, -1.f), 1.f));
        }
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
    if (threadIdx.x == 0) {
        const double* r = red[0];
        const double W = r[3 * kMetricsMaxC + 5];
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
        const double mx = r[3 * kMetricsMaxC] / n, my = r[3 * kMetricsMaxC + 1] / n;
        const double vx = (r[3 * kMetricsMaxC + 2] - n * mx * mx) / (n - 1.0), vy = (r[3 * kMetricsMaxC + 3] - n * my * my) / (n - 1.0);
        const double cxy = r[3 * kMetricsMaxC + 4] / n - mx * my;
        const double C1 = 0.01 * 0.01, C2 = 0.03 * 0.03;
        double* o = out + static_cast<size_t>(p) * kMetricsOut;
        o[0] = mae;
        o[1] = mse;
        o[2] = mse <= 1e-12 ? 99.0 : 10.0 * log10(1.0 / mse);
        o[3] = ((2 * mx * my + C1) * (2 * cxy + C2)) / ((mx * mx + my * my + C1) * (vx + vy + C2) + 1e-8);
        o[4] = r[3 * kMetricsMaxC + 6] / W;              // NaN for an empty mask, like torch's mean of nothing
        o[5] = 100.0 * sqrt(eg / C) * 4.0;
        o[6] = W;
        o[7] = 0.0;
    }
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

// One thread per canvas pixel; covering patches visited in ascending (row, col) = ascending patch index.
__global__ void __launch_bounds__(128) stitch_gather_kernel(const float* __restrict__ preds,
                                                            const int32_t* __restrict__ grid_map, int C, int ps, int stride,
                                                            int nrows, int ncols, int SH, int SW, float* __restrict__ canvas,
                                                            uint8_t* __restrict__ cover) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int y = blockIdx.y;
    if (x >= SW) return;
    const int i_lo = y >= ps ? (y - ps) / stride + 1 : 0;
    const int i_hi = min(nrows - 1, y / stride);
    const int j_lo = x >= ps ? (x - ps) / stride + 1 : 0;
    const int j_hi = min(ncols - 1, x / stride);
    float acc[kStitchMaxC];
#pragma unroll
    for (int c = 0; c < kStitchMaxC; ++c) acc[c] = 0.f;
    float cnt = 0.f;
    const size_t pp = static_cast<size_t>(ps) * ps;
    for (int i = i_lo; i <= i_hi; ++i) {
        const int ly = y - i * stride;
        for (int j = j_lo; j <= j_hi; ++j) {
            const int p = grid_map[i * ncols + j];
            if (p < 0) continue;
            const int lx = x - j * stride;
            const float* src = preds + static_cast<size_t>(p) * C * pp + static_cast<size_t>(ly) * ps + lx;
#pragma unroll
            for (int c = 0; c < kStitchMaxC; ++c)
                if (c < C) acc[c] = __fadd_rn(acc[c], __ldg(src + c * pp));
            cnt += 1.f;
        }
    }
    const size_t plane = static_cast<size_t>(SH) * SW;
    const size_t o = static_cast<size_t>(y) * SW + x;
#pragma unroll
    for (int c = 0; c < kStitchMaxC; ++c)
        if (c < C) canvas[c * plane + o] = cnt > 0.f ? __fdiv_rn(acc[c], cnt) : 0.f;
    cover[o] = cnt > 0.f ? 1 : 0;
}

}  // namespace s1s2
