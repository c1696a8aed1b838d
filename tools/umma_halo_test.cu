// Micro-test: can a K-major 128B-swizzled UMMA operand be read from INSIDE a larger TMA-written halo tile, i.e. with a
// descriptor start address offset by whole 128-byte rows and an 8-row-group stride (SBO) that is not a multiple of
// 1024 bytes?  (If yes, one (th+2) x (tw+2) activation tile serves all nine taps of a 3x3 convolution.)
//   build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -std=c++17 -I../s1-to-s2_super-resolution_project-code_b200/csrc -o umma_halo_test umma_halo_test.cu
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "ptx_sm100.cuh"
using namespace s1s2;

constexpr int TW = 8, TH = 16, HW_ = TW + 2, HH_ = TH + 2;     // output tile 8 x 16, halo tile 10 x 18
#ifndef CH
#define CH 64
#endif
constexpr int IMG_W = 16, IMG_H = 24, C = CH, NOUT = 32;
constexpr int ROWB = C * 2;                                   // bytes per K-major row: 128 (SW128) or 64 (SW64)

struct P { CUtensorMap ta, tb; float* out; int variant; };

__global__ void __launch_bounds__(128, 1) k(const __grid_constant__ P p) {
    extern __shared__ uint8_t raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
    uint8_t* sa = smem;                       // 180 rows x 128 B
    uint8_t* sb = smem + 24576;               // 9 taps x 32 rows x 128 B
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 24576 + 9 * 4096);
    uint64_t* bar2 = bar + 1;
    uint32_t* slot = reinterpret_cast<uint32_t*>(bar + 2);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) { mbar_init(bar, 1); mbar_init(bar2, 1); mbar_fence_init(); }
    if (warp == 0) tmem_alloc<512>(slot);
    tc_fence_before(); __syncthreads(); tc_fence_after();
    const uint32_t tmem = *slot;
    if (threadIdx.x == 0) {
        mbar_expect_tx(bar, HW_ * HH_ * ROWB + 9 * NOUT * ROWB);
        tma_load_4d(sa, &p.ta, bar, 0, 4 - 1, 3 - 1, 0);          // tile origin (x=4, y=3): halo starts at (3, 2)
        for (int t = 0; t < 9; ++t) tma_load_2d(sb + t * 4096, &p.tb, bar, t * C, 0);   // 4096-byte slots either way
        mbar_wait(bar, 0);
        tc_fence_after();
        const uint32_t idesc = umma_idesc_f16(128, NOUT);
        for (int t = 0; t < 9; ++t) {
            const int ky = t / 3, kx = t % 3;
            const uint32_t start = smem_u32(sa) + (ky * HW_ + kx) * ROWB;
            uint64_t ad = 0;
            ad |= static_cast<uint64_t>((start & 0x3FFFFu) >> 4);
            ad |= static_cast<uint64_t>(1) << 16;
            ad |= static_cast<uint64_t>((HW_ * ROWB) >> 4) << 32;                    // SBO = 10 rows
            ad |= static_cast<uint64_t>(1) << 46;
            if (p.variant == 1) ad |= static_cast<uint64_t>((start >> 7) & 7) << 49;  // matrix base offset
            ad |= (ROWB == 128 ? 2ull : 4ull) << 61;
            const uint64_t bd = umma_smem_desc<ROWB>(smem_u32(sb + t * 4096));
            for (int kk = 0; kk < C / 16; ++kk) umma_f16(tmem + t * NOUT, ad + 2 * kk, bd + 2 * kk, idesc, kk != 0);
        }
        umma_commit(bar2);
    }
    mbar_wait(bar2, 0);
    tc_fence_after();
    for (int t = 0; t < 9; ++t) {
        uint32_t r[32];
        tmem_ld32(tmem + (static_cast<uint32_t>(warp * 32) << 16) + t * NOUT, r);
        tmem_ld_wait();
        const int m = warp * 32 + lane;
        for (int j = 0; j < 32; ++j) p.out[(t * 128 + m) * NOUT + j] = __uint_as_float(r[j]);
    }
    tc_fence_before(); __syncthreads();
    if (warp == 0) tmem_dealloc<512>(tmem);
}

typedef CUresult (*Enc)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                        const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                        CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
int main() {
    void* fp = nullptr; cudaDriverEntryPointQueryResult q;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &q);
    Enc enc = reinterpret_cast<Enc>(fp);
    std::vector<__half> ha(IMG_H * IMG_W * C), hb(NOUT * 9 * C);
    std::vector<float> fa(ha.size()), fb(hb.size());
    srand(1);
    for (size_t i = 0; i < ha.size(); ++i) { fa[i] = (rand() % 17 - 8) / 8.f; ha[i] = __float2half(fa[i]); }
    for (size_t i = 0; i < hb.size(); ++i) { fb[i] = (rand() % 9 - 4) / 4.f; hb[i] = __float2half(fb[i]); }
    __half *da, *db; float* dout;
    cudaMalloc(&da, ha.size() * 2); cudaMalloc(&db, hb.size() * 2); cudaMalloc(&dout, 9 * 128 * NOUT * 4);
    cudaMemcpy(da, ha.data(), ha.size() * 2, cudaMemcpyHostToDevice);
    cudaMemcpy(db, hb.data(), hb.size() * 2, cudaMemcpyHostToDevice);
    P p;
    { cuuint64_t d[4] = {C, IMG_W, IMG_H, 1}; cuuint64_t s[3] = {C * 2, IMG_W * C * 2, (cuuint64_t)IMG_H * IMG_W * C * 2};
      cuuint32_t b[4] = {C, HW_, HH_, 1}; cuuint32_t e[4] = {1, 1, 1, 1};
      CUresult r = enc(&p.ta, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, da, d, s, b, e, CU_TENSOR_MAP_INTERLEAVE_NONE,
                       (C == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B), CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      if (r) { printf("tmap a %d\n", (int)r); return 1; } }
    { cuuint64_t d[2] = {9 * C, NOUT}; cuuint64_t s[1] = {9 * C * 2}; cuuint32_t b[2] = {C, NOUT}; cuuint32_t e[2] = {1, 1};
      CUresult r = enc(&p.tb, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, db, d, s, b, e, CU_TENSOR_MAP_INTERLEAVE_NONE,
                       (C == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B), CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      if (r) { printf("tmap b %d\n", (int)r); return 1; } }
    p.out = dout;
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 70000);
    for (int variant = 0; variant < 2; ++variant) {
        p.variant = variant;
        cudaMemset(dout, 0, 9 * 128 * NOUT * 4);
        k<<<1, 128, 70000>>>(p);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("variant %d: CUDA error %s\n", variant, cudaGetErrorString(e)); return 2; }
        std::vector<float> out(9 * 128 * NOUT);
        cudaMemcpy(out.data(), dout, out.size() * 4, cudaMemcpyDeviceToHost);
        double worst = 0; int bad = 0;
        for (int t = 0; t < 9; ++t) for (int m = 0; m < 128; ++m) for (int n = 0; n < NOUT; ++n) {
            const int ly = m / TW, lx = m % TW, y = 3 + ly + t / 3 - 1, x = 4 + lx + t % 3 - 1;
            double ref = 0;
            if (y >= 0 && y < IMG_H && x >= 0 && x < IMG_W)
                for (int c = 0; c < C; ++c) ref += fa[(y * IMG_W + x) * C + c] * fb[(n * 9 + t) * C + c];
            const double err = fabs(ref - out[(t * 128 + m) * NOUT + n]);
            if (err > 1e-3) ++bad;
            if (err > worst) worst = err;
        }
        printf("variant %d (base_offset %s): max |err| %.4g, mismatches %d / %d\n", variant, variant ? "set" : "0", worst, bad, 9 * 128 * NOUT);
    }
    return 0;
}
