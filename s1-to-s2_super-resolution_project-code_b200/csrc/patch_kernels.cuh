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
