// libs1s2_b200.so -- host side of the C ABI declared in include/s1s2_b200.h.
//
// Owns: the fp16 NHWC activation arena, the repacked fp16 weights, one TMA tensor map pair per convolution and
// the launch sequence of the denoiser (16 launches per model call: 13 3x3 convs, 3 transposed convs; the 1x1
// head and the scheduler update ride in the last conv's epilogue).  sm_100a only, no CPU fallback: s1s2_create
// fails on anything that is not a compute-capability-10.x device.
//
// Reference semantics: UNetSmall.forward (Evaluation/DDIM_Multi-step.py:42-53) and the sampler loops listed in
// the header.
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>

#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/s1s2_b200.h"
#include "conv_umma.cuh"
#include "conv_px.cuh"
#include "patch_kernels.cuh"

using namespace s1s2;

namespace {

thread_local std::string g_error;

void set_err(std::string* dst, const char* fmt, ...) {
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    *dst = buf;
}

#define CK(call)                                                                                   \
    do {                                                                                           \
        cudaError_t e_ = (call);                                                                   \
        if (e_ != cudaSuccess) {                                                                   \
            set_err(err, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
            return S1S2_ERR_CUDA;                                                                  \
        }                                                                                          \
    } while (0)

inline bool aligned_to(const void* p, size_t a) { return reinterpret_cast<uintptr_t>(p) % a == 0; }

// ---------------------------------------------------------------------------------------------- small kernels
// Input assembly (replaces t_idx.view.float.repeat + 2x torch.cat, DDIM_Multi-step.py:44-45,131):
// NCHW f32 -> NHWC16 fp16 pixel record [xlo0..3 | t t 0 0 | c0 c1 c2 c3 | xhi0..3].  x_t rides as an fp16 hi/lo pair
// (x = 4096*hi + lo: ~22 significant bits, range 2.7e8) and so does the WEIGHT of the time plane (exact t up to 2048,
// weight error 2^-22 instead of 2^-11).  Also records max|x_t| per patch for the range scale of this call.
__global__ void __launch_bounds__(256) pack_input_kernel(const float* __restrict__ x, size_t x_bstride,
                                                         const float* __restrict__ cond, size_t c_bstride,
                                                         const int64_t* __restrict__ t_idx, float t_const, float scale,
                                                         float* __restrict__ state, __half* __restrict__ xin16,
                                                         uint32_t* __restrict__ amax, int HW, size_t total) {
    const size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x;   // HW % 256 == 0: one patch per block
    const int b = static_cast<int>((blockIdx.x * static_cast<size_t>(blockDim.x)) / HW);
    float mx = 0.f;
    if (i < total) {
        const int pix = static_cast<int>(i - static_cast<size_t>(b) * HW);
        float t = t_const;
        if (t_idx != nullptr) {         // outside the fp16-exact integer range the time planes would be silently rounded:
            const int64_t ti = t_idx[b];                    // poison the patch instead -- a NaN amax makes the head write NaN
            if (ti >= 0 && ti <= 2048) t = static_cast<float>(ti);
            else { t = 0.f; mx = __uint_as_float(kAmaxPoison); }
        }
        float xv[4], cv[4], hi[4], lo[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            xv[k] = __fmul_rn(x[b * x_bstride + static_cast<size_t>(k) * HW + pix], scale);
            cv[k] = cond[b * c_bstride + static_cast<size_t>(k) * HW + pix];
            if (state != nullptr) state[(static_cast<size_t>(b) * 4 + k) * HW + pix] = xv[k];
            split_x(xv[k], hi[k], lo[k]);
            if (__float_as_uint(mx) != kAmaxPoison) mx = fmaxf(mx, fabsf(xv[k]));
        }
        uint4 lo4, hi4;
        lo4.x = pack_half2_sat(lo[0], lo[1]);
        lo4.y = pack_half2_sat(lo[2], lo[3]);
        lo4.z = pack_half2_sat(t, t);
        lo4.w = 0u;
        hi4.x = pack_half2_sat(cv[0], cv[1]);
        hi4.y = pack_half2_sat(cv[2], cv[3]);
        hi4.z = pack_half2_sat(hi[0], hi[1]);
        hi4.w = pack_half2_sat(hi[2], hi[3]);
        uint4* dst = reinterpret_cast<uint4*>(xin16 + i * 16);
        dst[0] = lo4;
        dst[1] = hi4;
    }
    const uint32_t wmax = __reduce_max_sync(0xffffffffu, __float_as_uint(mx));
    if ((threadIdx.x & 31) == 0 && wmax != 0u) atomicMax(amax + b, wmax);
}

// Unit-normal initial noise keyed by GLOBAL patch id (scene sharding: the draw of a patch does not depend on the rank or
// batch slot it lands in).  Philox4x32-10, key = seed, counter = (element / 4, id lo, id hi, tag); two Box-Muller pairs.
__global__ void __launch_bounds__(256) patch_noise_kernel(const int64_t* __restrict__ ids, uint32_t seed_lo, uint32_t seed_hi,
                                                          size_t elems4, float4* __restrict__ out) {
    const size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x;
    if (i >= elems4) return;
    const uint64_t id = static_cast<uint64_t>(ids[blockIdx.y]);
    uint32_t u[4];
    philox4x32_10(static_cast<uint32_t>(i), static_cast<uint32_t>(id), static_cast<uint32_t>(id >> 32), 0x4E4F4953u, seed_lo,
                  seed_hi, u);
    float z[4];
#pragma unroll
    for (int k = 0; k < 2; ++k) {
        const float a = (static_cast<float>(u[2 * k]) + 0.5f) * 2.3283064365386963e-10f;        // (0, 1)
        const float b = (static_cast<float>(u[2 * k + 1]) + 0.5f) * 2.3283064365386963e-10f;
        const float r = sqrtf(-2.f * logf(a));
        float sn, cs;
        sincospif(2.f * b, &sn, &cs);
        z[2 * k] = r * cs;
        z[2 * k + 1] = r * sn;
    }
    out[blockIdx.y * elems4 + i] = make_float4(z[0], z[1], z[2], z[3]);
}

// Conv2d weight OIHW f32 -> [cout][tap][cin] fp16 (K-major rows for the UMMA B operand).
// `cin_pad` >= cin: input channels past cin get zero weights (down1.0.0 reads 96 real + 32 always-zero channels so that
// it runs as two 64-channel chunks).
__global__ void repack_conv3_kernel(const float* __restrict__ w, __half* __restrict__ out, int cout, int cin, int cin_pad) {
    const size_t n = static_cast<size_t>(cout) * 9 * cin_pad;
    for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < n;
         i += static_cast<size_t>(gridDim.x) * blockDim.x) {
        const int ci = static_cast<int>(i % cin_pad);
        const int tap = static_cast<int>((i / cin_pad) % 9);
        const int co = static_cast<int>(i / (static_cast<size_t>(cin_pad) * 9));
        out[i] = ci < cin ? __float2half_rn(w[(static_cast<size_t>(co) * cin + ci) * 9 + tap]) : __float2half_rn(0.f);
    }
}
// inc.0.weight [cout][9][3][3] -> [cout][tap][16] in the pixel-record order of pack_input_kernel.
__global__ void repack_inc_kernel(const float* __restrict__ w, __half* __restrict__ out, int cout) {
    const int n = cout * 9 * 16;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const int slot = i % 16;
        const int tap = (i / 16) % 9;
        const int co = i / (16 * 9);
        auto W = [&](int ci) { return w[(static_cast<size_t>(co) * 9 + ci) * 9 + tap]; };
        float v = 0.f;
        if (slot < 4) v = W(slot);                       // x_t, lo part
        else if (slot >= 12) v = __half2float(__float2half_rn(W(slot - 12))) * kXSplit;            // x_t, hi part
        else if (slot == 4) v = __half2float(__float2half_rn(W(8)));                               // t, hi part
        else if (slot == 5) v = W(8) - __half2float(__float2half_rn(W(8)));                        // t, lo part
        else if (slot >= 8 && slot < 12) v = W(4 + slot - 8);                                      // cond
        out[i] = __float2half_rn(v);
    }
}
// ConvTranspose2d weight IOHW f32 [cin][cout][2][2] -> [(ky*2+kx)*cout + co][cin] fp16.
__global__ void repack_convt_kernel(const float* __restrict__ w, __half* __restrict__ out, int cin, int cout) {
    const size_t n = static_cast<size_t>(4) * cout * cin;
    for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < n;
         i += static_cast<size_t>(gridDim.x) * blockDim.x) {
        const int ci = static_cast<int>(i % cin);
        const int co = static_cast<int>((i / cin) % cout);
        const int tap = static_cast<int>(i / (static_cast<size_t>(cin) * cout));
        out[i] = __float2half_rn(w[(static_cast<size_t>(ci) * cout + co) * 4 + tap]);
    }
}
__global__ void tile_bias_kernel(const float* __restrict__ b, float* __restrict__ out, int cout, int reps) {
    const int n = cout * reps;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) out[i] = b[i % cout];
}
// Debug / per-layer parity tap: NHWC fp16 view -> NCHW f32.
__global__ void unpack_activation_kernel(const __half* __restrict__ src, int cpitch, int C, int HW, float* __restrict__ out,
                                         size_t total) {
    const size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x;   // over B*C*HW, NCHW order
    if (i >= total) return;
    const int pix = static_cast<int>(i % HW);
    const int c = static_cast<int>((i / HW) % C);
    const size_t b = i / (static_cast<size_t>(HW) * C);
    out[i] = __half2float(src[(b * HW + pix) * cpitch + c]);
}

// ---------------------------------------------------------------------------------------------- tensor maps
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode_fn() {
    static EncodeTiledFn fn = nullptr;
    if (fn == nullptr) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
    }
    return fn;
}

CUtensorMapSwizzle swizzle_for(int kbox) {
    return kbox == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : (kbox == 32 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B);
}

// ---------------------------------------------------------------------------------------------- layers
enum KernelId { K_INC = 0, K_C96IN, K_STORE, K_POOL, K_CONVT, K_STORE256, K_POOL256, K_CONVT256, K_N96, K_HEAD, K_PX_STORE, K_PX_HEAD32, K_HSTORE, K_HPOOL, K_HSTORE256, K_HPOOL256, K_HINC, K_HC96IN, K_HSTORE96, K_HPOOL96, K_HHEAD96, K_HINC64, K_HSTORE128, K_HPOOL128, K_HSTORE64, K_PX_HEAD64, K_COUNT };

struct KernelInfo {
    void (*fn)(const ConvParams);
    int block_n, kbox, boxes, smem, mode, ctas, threads;
    int subc;                // channels per output staging sub-tile = inner box of the store tensor map (32 or 64)
    bool px;                 // conv_px_kernel (pixels on N) instead of conv_umma_kernel
    bool halo;               // conv_umma_kernel halo mode (8 x 16 tile, one activation halo tile per chunk)
};

template <int BN, int KB, int BX, int ST, int MODE, int CTAS = 2, bool HALO = false, bool WRES = false, int SBUF = 1,
          int TPS = 1, int EPIWG = 1, int HSLOTS = kHaloSlotsDefault, int SUBC = 32>
KernelInfo make_kernel() {
    KernelInfo k;
    k.fn = conv_umma_kernel<BN, KB, BX, ST, MODE, CTAS, HALO, WRES, SBUF, TPS, EPIWG, HSLOTS, SUBC>;
    k.subc = SUBC;
    k.threads = 128 + 128 * EPIWG;
    k.ctas = CTAS;
    k.px = false;
    k.halo = HALO;
    k.block_n = BN;
    k.kbox = KB;
    k.boxes = BX;
    k.smem = ConvSmem<BN, KB, BX, ST, MODE, CTAS, HALO, SBUF, TPS, HSLOTS, SUBC>::kBytes;
    k.mode = MODE;
    return k;
}

template <int KB, int ST, int MODE, int TPS = 1, int HSLOTS = 2, int NQ = 3>
KernelInfo make_px_kernel() {
    KernelInfo k;
    k.fn = conv_px_kernel<KB, ST, MODE, TPS, HSLOTS, NQ>;
    k.subc = 32;
    k.threads = kPxThreads;
    k.block_n = 32 * NQ;
    k.kbox = KB;
    k.boxes = 1;
    k.smem = PxSmem<KB, ST, TPS, HSLOTS, NQ>::kBytes;
    k.mode = MODE;
    k.ctas = 1;
    k.px = true;
    k.halo = false;
    return k;
}

struct KernelTable {
    KernelInfo t[K_COUNT];
    KernelTable() {
        // One tcgen05.mma (128 rows per CTA, K = 16) occupies the tensor pipe for N/2 cycles; the widest tile that divides
        // the GEMM N (Cout, or 4*Cout for the transposed convs) costs the least weight traffic and issue work per flop:
        // N = 256 where N % 256 == 0, else 192 / 128 / 96 / 64.  The first block are the round-1 one-box-per-tap kernels,
        // kept as A/B baselines (S1S2_NO_HALO) and for the transposed convs.
        t[K_INC] = make_kernel<96, 16, 3, 4, MODE_STORE>();     // Cin = 16-channel pixel record
        t[K_C96IN] = make_kernel<192, 32, 3, 4, MODE_STORE>();  // Cin = 96 (64-byte swizzle rows)
        t[K_STORE] = make_kernel<192, 64, 1, 6, MODE_STORE>();
        t[K_POOL] = make_kernel<192, 64, 1, 7, MODE_POOL>();
        // transposed convs: short K (3 / 6 / 12 stages per tile) under an epilogue of 6 / 8 column chunks -- two epilogue
        // warpgroups, each draining half of the chunks (ncu source view: one group needed longer than the tile's MMAs)
        t[K_CONVT] = make_kernel<192, 64, 1, 6, MODE_CONVT, 2, false, false, 1, 1, 2>();
        t[K_STORE256] = make_kernel<256, 64, 1, 4, MODE_STORE>();
        t[K_POOL256] = make_kernel<256, 64, 1, 6, MODE_POOL>();
        t[K_CONVT256] = make_kernel<256, 64, 1, 4, MODE_CONVT, 2, false, false, 1, 1, 2, kHaloSlotsDefault, 64>();   // 128-byte store rows
        t[K_N96] = make_kernel<96, 64, 1, 8, MODE_STORE>();     // Cout = 96
        // halo mode (default for 3x3, Cin % 64 == 0); 64-channel (128-byte) rows in the output staging: the TMA unit writes a
        // tile row by row, and with 64-byte rows the stores of a short-K tile (down1.0.0, the transposed convs) took as long
        // as its MMAs
        t[K_HSTORE] = make_kernel<192, 64, 1, 8, MODE_STORE, 2, true, false, 1, 1, 1, kHaloSlotsDefault, 64>();
        t[K_HPOOL] = make_kernel<192, 64, 1, 10, MODE_POOL, 2, true, false, 1, 1, 1, kHaloSlotsDefault, 64>();
        t[K_HSTORE256] = make_kernel<256, 64, 1, 5, MODE_STORE, 2, true, false, 1, 1, 1, kHaloSlotsDefault, 64>();
        t[K_HPOOL256] = make_kernel<256, 64, 1, 8, MODE_POOL, 2, true, false, 1, 1, 1, kHaloSlotsDefault, 64>();
        // Short-K layers: several taps per ring stage (the MMA issue loop costs ~300 cycles per stage: ncu source view)
        t[K_HINC] = make_kernel<96, 16, 1, 1, MODE_STORE, 2, true, true, 4, 9, 2, 8>();   // inc.0: 32-byte rows, nine resident weight
                                                                  // tiles, 4 staging buffers, two epilogue warpgroups, 8 halo slots
        t[K_HC96IN] = make_kernel<192, 32, 1, 6, MODE_STORE, 2, true, false, 1, 3, 1, kHaloSlotsDefault, 64>();  // down1.0.0: exact K = 96 per tap as three
                                                                                       // 32-channel chunks, one kernel row per stage
        t[K_PX_STORE] = make_px_kernel<64, 5, MODE_STORE>();            // conv1.0: pixels on N (see conv_px.cuh)
        t[K_PX_HEAD32] = make_px_kernel<32, 4, MODE_HEAD, 3, 3>();        // same with exact 32-channel chunks, one kernel row per
                                                                          // stage, three halo slots (default)
        t[K_HEAD] = make_kernel<96, 32, 3, 6, MODE_HEAD>();     // conv1.2 + outc + scheduler
        // Cout-on-N tiles of 96 columns, three taps per stage (one tcgen05.mma of M = 256, N = 96, K = 16 is 48 cycles, so a
        // one-tap stage of four would sit under the ~300-cycle issue loop).  Used (a) when a layer has too few N = 192 / 256
        // tiles to fill the 74 CTA pairs (small batches: see pick_variant) and (b) as the A/B partner of the pixels-on-N
        // kernels for the Cout = 96 layers (S1S2_C1_UMMA=1).
        t[K_HSTORE96] = make_kernel<96, 64, 1, 6, MODE_STORE, 2, true, false, 1, 3>();
        t[K_HPOOL96] = make_kernel<96, 64, 1, 6, MODE_POOL, 2, true, false, 1, 3>();
        t[K_HHEAD96] = make_kernel<96, 32, 1, 3, MODE_HEAD, 2, true, false, 1, 9>();   // conv1.2: exact 32-channel chunks, nine taps per stage
        // base_ch = 64 (the class default of Train_Orignal.py:99): channel counts 64 / 128 / 256 / 512.  The 256-column
        // kernels above serve Cout = 256 / 512 and the three transposed convs; these cover Cout = 128 and 64 (three taps
        // per stage for the same reason as the 96-column tiles), the first layer and the head.
        t[K_HINC64] = make_kernel<64, 16, 1, 1, MODE_STORE, 2, true, true, 4, 9, 2, 8>();
        t[K_HSTORE128] = make_kernel<128, 64, 1, 4, MODE_STORE, 2, true, false, 1, 3, 1, kHaloSlotsDefault, 64>();
        t[K_HPOOL128] = make_kernel<128, 64, 1, 5, MODE_POOL, 2, true, false, 1, 3, 1, kHaloSlotsDefault, 64>();
        t[K_HSTORE64] = make_kernel<64, 64, 1, 8, MODE_STORE, 2, true, false, 1, 3, 1, kHaloSlotsDefault, 64>();
        t[K_PX_HEAD64] = make_px_kernel<32, 4, MODE_HEAD, 3, 3, 2>();
    }
};
const KernelInfo* kernel_table() {
    static const KernelTable table;        // function-local static: initialised once, thread-safe (C++11)
    return table.t;
}

struct Layer {
    const char* name;        // state_dict prefix
    KernelId kid;
    int level;               // input resolution = (H >> level, W >> level)
    int cin;                 // K per tap (padded for inc and down1.0.0)
    int cin_real = 0;        // input channels of the state_dict tensor when cin is padded (0: same as cin)
    bool first = false;      // inc.0: reads the 16-slot pixel record, applies the range scale
    int ntot;                // GEMM N (CONVT: 4 * cout)
    int cout;                // real output channels per pixel
    int taps_w;
    const __half* src;       // input view
    int src_pitch;
    __half* dst;             // output view (channel offset applied)
    int dst_pitch;
    __half* w = nullptr;
    size_t w_bytes = 0;
    float* bias = nullptr;
    ConvParams p;
    // Alternative tilings of the same layer (same weights, same arithmetic per output element): narrower N tiles that
    // fill the 74 CTA pairs when the batch is small.  pick_variant chooses per launch.
    struct Alt { KernelId kid; ConvParams p; };
    std::vector<Alt> alts;
};

struct View {
    const char* name;
    const __half* ptr;
    int pitch, C, level;
};

}  // namespace

struct s1s2_handle {
    int device = 0;
    int base_ch = 96;
    int H = 0, W = 0, max_batch = 0, nalloc = 0;
    int num_sms = 0;
    bool weights_loaded = false;
    std::string err;
    int64_t launches = 0;
    std::vector<void*> allocs;
    std::vector<Layer> layers;
    std::vector<View> views;
    __half* xin16 = nullptr;
    uint64_t noise_seed = 0x5EED5EEDull;   // S1S2_STEP_PHILOX key and patch id of batch slot 0
    uint32_t patch_base = 0;
    uint32_t* amax = nullptr;     // [2][max_batch] float bits of max|x_t| per patch, ping-pong across model calls
    float head_w[kHeadOut * kHeadIn];
    float head_b[kHeadOut];
    // staging for s1s2_sample_host / s1s2_sample_host_stream (two sets: batch i+1 uploads while batch i computes)
    float *st_cond[2] = {nullptr, nullptr}, *st_x[2] = {nullptr, nullptr}, *st_out[2] = {nullptr, nullptr};
    cudaStream_t s_in = nullptr, s_out = nullptr;                  // copy streams of the pipelined host entry
    cudaEvent_t ev_in[2] = {nullptr, nullptr}, ev_packed[2] = {nullptr, nullptr}, ev_done[2] = {nullptr, nullptr},
                ev_out[2] = {nullptr, nullptr}, ev_entry = nullptr;
    unsigned long long* sat_counts = nullptr;                      // [views] scratch of s1s2_debug_saturation_count
#ifdef S1S2_TIMELINE
    unsigned long long* tl_dbg = nullptr;                          // [layers][16] stamps of the chain being traced, or nullptr
#endif
};

namespace {

struct TileGeom {
    int tw_log2, th_log2, tn, tiles_x, tiles_y;
};

int ilog2(int v) {
    int l = 0;
    while ((1 << (l + 1)) <= v) ++l;
    return l;
}

TileGeom tile_geom(int Hl, int Wl) {
    TileGeom g;
    int tw = Wl & -Wl;
    if (tw > 16) tw = 16;
    int th = 1 << ilog2(Hl);
    if (th > 128 / tw) th = 128 / tw;
    g.tw_log2 = ilog2(tw);
    g.th_log2 = ilog2(th);
    g.tn = 128 / (tw * th);
    g.tiles_x = Wl / tw;
    g.tiles_y = (Hl + th - 1) / th;
    return g;
}

int dmalloc(s1s2_handle* h, void** p, size_t bytes, std::string* err) {
    CK(cudaMalloc(p, bytes));
    h->allocs.push_back(*p);
    return S1S2_OK;
}

// Timing events owned for the duration of one API call.
struct EventSet {
    std::vector<cudaEvent_t> ev;
    ~EventSet() { for (cudaEvent_t e : ev) cudaEventDestroy(e); }
    int create(int n, std::string* err) {
        ev.reserve(n);
        for (int i = 0; i < n; ++i) {
            cudaEvent_t e;
            CK(cudaEventCreate(&e));
            ev.push_back(e);
        }
        return S1S2_OK;
    }
    cudaEvent_t operator[](int i) const { return ev[i]; }
};

// One-time per-device settings of the handle-less entry points (the attributes live in the device's context).
struct DeviceOnce {
    std::mutex mu;
    bool metrics_attr[64] = {};
    bool pool_keep[64] = {};
};
DeviceOnce& device_once() {
    static DeviceOnce d;
    return d;
}

// conv_px_kernel: pixels (16 x 16 tile of one image) on N, weight rows on M.
int build_px_params(s1s2_handle* h, Layer& L, const KernelInfo& k, ConvParams& p, EncodeTiledFn enc, std::string* err) {
    const int Hl = h->H >> L.level, Wl = h->W >> L.level;
    if (Hl % 16 != 0 || Wl % 8 != 0 || L.cout > 128 || L.taps_w != 3) {
        set_err(err, "layer %s: geometry does not fit the pixels-on-N kernel", L.name);
        return S1S2_ERR_INVALID;
    }
    memset(&p, 0, sizeof(p));
    const CUtensorMapSwizzle sw = swizzle_for(k.kbox);
    {   // activations: (C, W, H, N), box (kbox, 16, 16, 1); channels past Cin are zero-filled
        cuuint64_t dims[4] = {static_cast<cuuint64_t>(L.cin), static_cast<cuuint64_t>(Wl), static_cast<cuuint64_t>(Hl),
                              static_cast<cuuint64_t>(h->nalloc)};
        cuuint64_t strides[3] = {static_cast<cuuint64_t>(L.src_pitch) * 2, static_cast<cuuint64_t>(Wl) * L.src_pitch * 2,
                                 static_cast<cuuint64_t>(Hl) * Wl * L.src_pitch * 2};
        cuuint32_t box[4] = {static_cast<cuuint32_t>(k.kbox), kPxHaloW, kPxHaloH, 1};      // 8 x 32 tile + 1-pixel ring
        cuuint32_t estr[4] = {1, 1, 1, 1};
        CUresult r = enc(&p.tmap_a, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, const_cast<__half*>(L.src), dims, strides, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) { set_err(err, "layer %s: activation tensor map rejected (CUresult %d)", L.name, static_cast<int>(r)); return S1S2_ERR_CUDA; }
    }
    {   // weights: (K, Cout), box (kbox, 128): rows past Cout are zero-filled
        const int ktot = 9 * L.cin;
        cuuint64_t dims[2] = {static_cast<cuuint64_t>(ktot), static_cast<cuuint64_t>(L.cout)};
        cuuint64_t strides[1] = {static_cast<cuuint64_t>(ktot) * 2};
        cuuint32_t box[2] = {static_cast<cuuint32_t>(k.kbox), 128};
        cuuint32_t estr[2] = {1, 1};
        CUresult r = enc(&p.tmap_b, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, L.w, dims, strides, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) { set_err(err, "layer %s: weight tensor map rejected (CUresult %d)", L.name, static_cast<int>(r)); return S1S2_ERR_CUDA; }
    }
    if (k.mode == MODE_STORE) {
        cuuint64_t dims[4] = {static_cast<cuuint64_t>(L.cout), static_cast<cuuint64_t>(Wl), static_cast<cuuint64_t>(Hl),
                              static_cast<cuuint64_t>(h->nalloc)};
        cuuint64_t strides[3] = {static_cast<cuuint64_t>(L.dst_pitch) * 2, static_cast<cuuint64_t>(Wl) * L.dst_pitch * 2,
                                 static_cast<cuuint64_t>(Hl) * Wl * L.dst_pitch * 2};
        cuuint32_t box[4] = {32, 8, 16, 1};     // half a tile (8 wide x 16 tall) per store
        cuuint32_t estr[4] = {1, 1, 1, 1};
        CUresult r = enc(&p.tmap_out, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, L.dst, dims, strides, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) { set_err(err, "layer %s: output tensor map rejected (CUresult %d)", L.name, static_cast<int>(r)); return S1S2_ERR_CUDA; }
    }
    p.bias = L.bias;
    p.out = L.dst;
    p.out_cpitch = L.dst_pitch;
    p.H = Hl;
    p.W = Wl;
    p.taps_w = 3;
    p.chunks = (L.cin + k.kbox - 1) / k.kbox;
    p.tap_kstride = L.cin;
    p.cout = L.cout;
    p.num_n_tiles = 1;
    return S1S2_OK;
}

int build_layer_params(s1s2_handle* h, Layer& L, KernelId kid, ConvParams& p, std::string* err) {
    const KernelInfo& k = kernel_table()[kid];
    EncodeTiledFn enc = get_encode_fn();
    if (enc == nullptr) {
        set_err(err, "cuTensorMapEncodeTiled is not available from this driver");
        return S1S2_ERR_CUDA;
    }
    const int Hl = h->H >> L.level, Wl = h->W >> L.level;
    if (k.px) return build_px_params(h, L, k, p, enc, err);
    TileGeom g = tile_geom(Hl, Wl);
    if (k.halo) {            // 8 wide x 16 tall tile of one image; partial tiles are zero-filled on load, clipped on store
        g.tw_log2 = 3;
        g.th_log2 = 4;
        g.tn = 1;
        g.tiles_x = (Wl + 7) / 8;
        g.tiles_y = (Hl + 15) / 16;
    }
    const int chunks_ = L.cin / k.kbox;
    const bool boxes_ok = chunks_ == 1 ? (L.taps_w * L.taps_w) % k.boxes == 0 : chunks_ % k.boxes == 0;
    if (L.cin % k.kbox != 0 || !boxes_ok || L.ntot % k.block_n != 0) {
        set_err(err, "layer %s: geometry does not fit kernel (cin %d, N %d)", L.name, L.cin, L.ntot);
        return S1S2_ERR_INVALID;
    }
    memset(&p, 0, sizeof(p));
    {   // activations: (C, W, H, N), box (kbox, tw, th, tn)
        cuuint64_t dims[4] = {static_cast<cuuint64_t>(L.cin), static_cast<cuuint64_t>(Wl), static_cast<cuuint64_t>(Hl),
                              static_cast<cuuint64_t>(h->nalloc)};
        cuuint64_t strides[3] = {static_cast<cuuint64_t>(L.src_pitch) * 2, static_cast<cuuint64_t>(Wl) * L.src_pitch * 2,
                                 static_cast<cuuint64_t>(Hl) * Wl * L.src_pitch * 2};
        cuuint32_t box[4] = {static_cast<cuuint32_t>(k.kbox), 1u << g.tw_log2, 1u << g.th_log2,
                             static_cast<cuuint32_t>(g.tn)};
        if (k.halo) { box[1] = kHaloW; box[2] = kHaloH; }        // tile + 1-pixel ring, fetched once per channel chunk
        cuuint32_t estr[4] = {1, 1, 1, 1};
        CUresult r = enc(&p.tmap_a, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, const_cast<__half*>(L.src), dims, strides, box,
                         estr, CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle_for(k.kbox), CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) {
            set_err(err, "layer %s: activation tensor map rejected (CUresult %d)", L.name, static_cast<int>(r));
            return S1S2_ERR_CUDA;
        }
    }
    {   // weights: (K, N), box (kbox, block_n)
        const int ktot = L.taps_w * L.taps_w * L.cin;
        cuuint64_t dims[2] = {static_cast<cuuint64_t>(ktot), static_cast<cuuint64_t>(L.ntot)};
        cuuint64_t strides[1] = {static_cast<cuuint64_t>(ktot) * 2};
        cuuint32_t box[2] = {static_cast<cuuint32_t>(k.kbox), static_cast<cuuint32_t>(k.block_n / k.ctas)};   // per CTA of the group
        cuuint32_t estr[2] = {1, 1};
        CUresult r = enc(&p.tmap_b, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, L.w, dims, strides, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle_for(k.kbox), CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) {
            set_err(err, "layer %s: weight tensor map rejected (CUresult %d)", L.name, static_cast<int>(r));
            return S1S2_ERR_CUDA;
        }
    }
    if (k.mode == MODE_CONVT) {
        // destination rows of parity ky as a 5-D view (c, kx, x, y, n) of the double-resolution output: the pixel shuffle
        // of one tap is then a plain tiled store, box = 32 channels x 1 x the input tile
        const int Wo = 2 * Wl, Ho = 2 * Hl;
        const cuuint64_t pb = static_cast<cuuint64_t>(L.dst_pitch) * 2;            // bytes per output pixel
        cuuint64_t dims[5] = {static_cast<cuuint64_t>(L.cout), 2, static_cast<cuuint64_t>(Wl), static_cast<cuuint64_t>(Hl),
                              static_cast<cuuint64_t>(h->nalloc)};
        cuuint64_t strides[4] = {pb, 2 * pb, 2 * Wo * pb, static_cast<cuuint64_t>(Ho) * Wo * pb};
        if (L.cout % k.subc != 0) {
            set_err(err, "layer %s: %d output channels per tap do not fit %d-channel store sub-tiles", L.name, L.cout, k.subc);
            return S1S2_ERR_INVALID;
        }
        cuuint32_t box[5] = {static_cast<cuuint32_t>(k.subc), 1u, 1u << g.tw_log2, 1u << g.th_log2, static_cast<cuuint32_t>(g.tn)};
        cuuint32_t estr[5] = {1, 1, 1, 1, 1};
        for (int ky = 0; ky < 2; ++ky) {
            CUresult r = enc(ky ? &p.tmap_out2 : &p.tmap_out, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 5, L.dst + static_cast<size_t>(ky) * Wo * L.dst_pitch,
                             dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle_for(k.subc),
                             CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
            if (r != CUDA_SUCCESS) {
                set_err(err, "layer %s: output tensor map (ky %d) rejected (CUresult %d)", L.name, ky, static_cast<int>(r));
                return S1S2_ERR_CUDA;
            }
        }
    } else if (k.mode != MODE_HEAD) {
        // destination: (C, W, H, N) of the output, box = 32 channels x the output pixels of one M tile, 64B swizzle.
        //   STORE: same resolution.  POOL: half resolution, box = the pooled tile.
        const int sh = k.mode == MODE_POOL ? 1 : 0;
        if (sh && (g.tw_log2 < 1 || g.th_log2 < 1)) {
            set_err(err, "layer %s: pooled tile needs an M tile at least 2 x 2 pixels", L.name);
            return S1S2_ERR_INVALID;
        }
        const int Ho = Hl >> sh, Wo = Wl >> sh;
        cuuint64_t dims[4] = {static_cast<cuuint64_t>(L.cout), static_cast<cuuint64_t>(Wo), static_cast<cuuint64_t>(Ho),
                              static_cast<cuuint64_t>(h->nalloc)};
        cuuint64_t strides[3] = {static_cast<cuuint64_t>(L.dst_pitch) * 2, static_cast<cuuint64_t>(Wo) * L.dst_pitch * 2,
                                 static_cast<cuuint64_t>(Ho) * Wo * L.dst_pitch * 2};
        cuuint32_t box[4] = {static_cast<cuuint32_t>(k.subc), (1u << g.tw_log2) >> sh, (1u << g.th_log2) >> sh, static_cast<cuuint32_t>(g.tn)};
        cuuint32_t estr[4] = {1, 1, 1, 1};
        CUresult r = enc(&p.tmap_out, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, L.dst, dims, strides, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle_for(k.subc), CU_TENSOR_MAP_L2_PROMOTION_NONE,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) {
            set_err(err, "layer %s: output tensor map rejected (CUresult %d)", L.name, static_cast<int>(r));
            return S1S2_ERR_CUDA;
        }
    }
    p.bias = L.bias;
    p.out = L.dst;
    p.out_cpitch = L.dst_pitch;
    p.H = Hl;
    p.W = Wl;
    p.B = 0;
    p.tw_log2 = g.tw_log2;
    p.th_log2 = g.th_log2;
    p.tiles_x = g.tiles_x;
    p.tiles_y = g.tiles_y;
    p.num_m_tiles = 0;
    p.num_n_tiles = L.ntot / k.block_n;
    p.fd_n = FastDiv::make(static_cast<uint32_t>(p.num_n_tiles));
    p.fd_tx = FastDiv::make(static_cast<uint32_t>(p.tiles_x));
    p.fd_ty = FastDiv::make(static_cast<uint32_t>(p.tiles_y));
    p.taps_w = L.taps_w;
    p.chunks = L.cin / k.kbox;
    p.tap_kstride = L.cin;
    p.cout = L.cout;
    p.flags = L.first ? LAYER_FLAG_FIRST : 0;
    return S1S2_OK;
}

// Launch with programmatic stream serialization: the kernel may be scheduled while its predecessor in the stream drains;
// it orders itself behind the predecessor's memory operations with griddepcontrol.wait (S1S2_NO_PDL=1: plain launches).
int launch_conv(const KernelInfo& k, int grid, const ConvParams& p, cudaStream_t st, std::string* err) {
    static const bool pdl = getenv("S1S2_NO_PDL") == nullptr;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(k.threads);
    cfg.dynamicSmemBytes = k.smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl ? 1 : 0;
    CK(cudaLaunchKernelEx(&cfg, k.fn, p));
    return S1S2_OK;
}

// CTA-pair kernels: group tiles (2 M tiles x 1 N tile) of a launch at batch B.
int group_tiles_of(const KernelInfo& k, const ConvParams& p, int B) {
    const int tn = 128 >> (p.tw_log2 + p.th_log2);
    const int m_tiles = p.tiles_x * p.tiles_y * ((B + tn - 1) / tn);
    return ((m_tiles + k.ctas - 1) / k.ctas) * p.num_n_tiles;
}

// Which tiling of the layer runs at batch B.  A launch takes waves x (time of one group tile); the K loop of a tile is
// N/2 cycles per tcgen05.mma (B300_MICROARCH.md: max(M,128) * N / (256 * cta_group)), so the cost of a variant is
// waves * (BLOCK_N + c), c = the per-tile fixed part (accumulator hand-over, epilogue tail) in the same unit.  At batch 64
// every layer has hundreds of tiles and the widest tile wins (least weight traffic per flop); at batch 1 the 64 x 64
// layers have 32 M tiles and N = 96 / 192 tiles are what fills the 74 CTA pairs.  Every variant accumulates each output
// element over (chunk, tap, k) in the same order, so the choice never changes a bit of the result.
const Layer::Alt* pick_variant(const s1s2_handle* h, const Layer& L, int B) {
    if (L.alts.empty()) return nullptr;
    const KernelInfo* kt = kernel_table();
    auto cost = [&](const KernelInfo& k, const ConvParams& p) {
        const int slots = h->num_sms / k.ctas;
        const int waves = (group_tiles_of(k, p, B) + slots - 1) / slots;
        return static_cast<long>(waves) * (k.block_n + 24);
    };
    long best = cost(kt[L.kid], L.p);
    const Layer::Alt* pick = nullptr;
    for (const Layer::Alt& a : L.alts) {          // alternates are listed widest first: ties keep the wider tile
        const long c = cost(kt[a.kid], a.p);
        if (c < best) { best = c; pick = &a; }
    }
    return pick;
}

int launch_layer(s1s2_handle* h, Layer& L, int B, const uint32_t* amax_in, uint32_t* amax_zero, cudaStream_t st,
                 std::string* err) {
    const Layer::Alt* alt = (L.p.perf_mode == 0) ? pick_variant(h, L, B) : nullptr;
    const KernelInfo& k = kernel_table()[alt != nullptr ? alt->kid : L.kid];
    ConvParams p = alt != nullptr ? alt->p : L.p;          // per-launch copy: the kernel takes it by value anyway
    p.B = B;
    p.amax_in = amax_in;
    p.amax_zero = amax_zero;
    {   // the next launch of the chain (the first layer again after the last one): its weights are prefetched into L2
        static const bool no_prefetch = getenv("S1S2_NO_WPREFETCH") != nullptr;
        const size_t li = static_cast<size_t>(&L - h->layers.data());
        const Layer& nxt = h->layers[(li + 1) % h->layers.size()];
        p.next_w = no_prefetch ? nullptr : reinterpret_cast<const uint8_t*>(nxt.w);
        p.next_w_bytes = static_cast<uint32_t>(nxt.w_bytes);
    }
#ifdef S1S2_TIMELINE
    p.dbg = h->tl_dbg != nullptr ? h->tl_dbg + 16 * (&L - h->layers.data()) : nullptr;
#endif
    int grid;
    if (k.px) {
        const int tiles = (p.W >> 3) * ((p.H + 31) >> 5) * B;
        grid = tiles < h->num_sms ? tiles : h->num_sms;
    } else {
        const int tn = 128 >> (p.tw_log2 + p.th_log2);
        p.num_m_tiles = p.tiles_x * p.tiles_y * ((B + tn - 1) / tn);
        const int group_tiles = group_tiles_of(k, p, B);                    // one CTA group per (ctas M tiles, 1 N tile)
        const int clusters = group_tiles < h->num_sms / k.ctas ? group_tiles : h->num_sms / k.ctas;
        grid = k.ctas * clusters;                                           // __cluster_dims__(ctas, 1, 1)
    }
    int rc = launch_conv(k, grid, p, st, err);
    if (rc != S1S2_OK) return rc;
    ++h->launches;
    return S1S2_OK;
}

// `amax_zero` ([B] or nullptr): cleared by the first layer's kernel for the head of THIS call to reduce max|x_next| into --
// its last readers were the kernels of the previous call, complete by the time any kernel of this call passes
// griddepcontrol.wait.  (A cudaMemsetAsync between calls would put a memset node between conv1.2 and the next inc.0 and
// break the programmatic-dependent-launch chain once per model call.)
int run_network(s1s2_handle* h, int B, const HeadParams& head_io, const uint32_t* amax_in, uint32_t* amax_zero,
                cudaStream_t st, std::string* err) {
    for (size_t i = 0; i < h->layers.size(); ++i) {
        Layer& L = h->layers[i];
        if (kernel_table()[L.kid].mode == MODE_HEAD) {
            HeadParams& hp = L.p.head;
            memcpy(hp.w, h->head_w, sizeof(hp.w));
            memcpy(hp.b, h->head_b, sizeof(hp.b));
            hp.x_t = head_io.x_t;
            hp.pred_out = head_io.pred_out;
            hp.noise = head_io.noise;
            hp.seed_lo = head_io.seed_lo; hp.seed_hi = head_io.seed_hi;
            hp.noise_stream = head_io.noise_stream; hp.patch_base = head_io.patch_base;
            hp.xin16 = head_io.xin16;
            hp.amax_out = head_io.amax_out;
            hp.step = head_io.step;
        }
        int rc = launch_layer(h, L, B, amax_in, L.first ? amax_zero : nullptr, st, err);
        if (rc != S1S2_OK) return rc;
    }
    return S1S2_OK;
}

int check_batch(s1s2_handle* h, int B) {
    if (h == nullptr) return S1S2_ERR_INVALID;
    if (!h->weights_loaded) {
        h->err = "weights not loaded: call s1s2_load_weights first";
        return S1S2_ERR_STATE;
    }
    if (B < 1 || B > h->max_batch) {
        set_err(&h->err, "batch %d outside [1, max_batch=%d]", B, h->max_batch);
        return S1S2_ERR_INVALID;
    }
    return S1S2_OK;
}

}  // namespace

// ================================================================================================ C ABI
extern "C" {

int s1s2_abi_version(void) { return 3; }   // 3: + sample_host_stream, patch_noise, stitch_weighted, debug_saturation_count

const char* s1s2_global_error(void) { return g_error.c_str(); }
const char* s1s2_last_error(const s1s2_handle* h) { return h != nullptr ? h->err.c_str() : g_error.c_str(); }
int64_t s1s2_launch_count(const s1s2_handle* h) { return h != nullptr ? h->launches : 0; }

int s1s2_create(s1s2_handle** out, int device, int in_ch, int out_ch, int base_ch, int H, int W, int max_batch) {
    std::string* err = &g_error;
    if (out == nullptr) return S1S2_ERR_INVALID;
    *out = nullptr;
    if (in_ch != 8 || out_ch != 4 || (base_ch != 96 && base_ch != 64)) {
        set_err(err, "unsupported architecture (in_ch %d, out_ch %d, base_ch %d): this library implements "
                     "UNetSmall(8, 4, base_ch) for base_ch = 96 (the scripts' default) and 64 (the class default)", in_ch, out_ch,
                base_ch);
        return S1S2_ERR_INVALID;
    }
    if (H < 16 || W < 16 || H % 16 != 0 || W % 16 != 0 || max_batch < 1) {
        set_err(err, "H and W must be multiples of 16 (got %d x %d), max_batch >= 1 (got %d)", H, W, max_batch);
        return S1S2_ERR_INVALID;
    }
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || device < 0 || device >= ndev) {
        set_err(err, "no CUDA device %d (found %d); this library has no CPU fallback", device, ndev);
        return S1S2_ERR_CUDA;
    }
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10) {
        set_err(err, "device %d is sm_%d%d; libs1s2_b200 is built for sm_100a (B200) only", device, prop.major, prop.minor);
        return S1S2_ERR_CUDA;
    }
    CK(cudaSetDevice(device));
    s1s2_handle* h = new s1s2_handle();
    h->device = device;
    h->base_ch = base_ch;
    h->H = H;
    h->W = W;
    h->max_batch = max_batch;
    h->num_sms = prop.multiProcessorCount;
    const int tn_max = tile_geom(H >> 3, W >> 3).tn;   // coarsest level packs the most images into one M tile
    h->nalloc = ((max_batch + tn_max - 1) / tn_max) * tn_max;

    const KernelInfo* kt = kernel_table();
    for (int i = 0; i < K_COUNT; ++i) {
        cudaError_t e = cudaFuncSetAttribute(reinterpret_cast<const void*>(kt[i].fn),
                                             cudaFuncAttributeMaxDynamicSharedMemorySize, kt[i].smem);
        if (e != cudaSuccess) {
            set_err(err, "cudaFuncSetAttribute(kernel %d, %d B smem): %s", i, kt[i].smem, cudaGetErrorString(e));
            delete h;
            return S1S2_ERR_CUDA;
        }
    }

    const size_t N = static_cast<size_t>(h->nalloc);
    const size_t P0 = static_cast<size_t>(H) * W, P1 = P0 / 4, P2 = P0 / 16, P3 = P0 / 64;
    // channel counts of the four resolutions; cat1 = [up1 (c1) | inc (c1) | pad]: for base_ch = 96 the pad is 32 channels
    // that stay zero (the K-padded A/B variant of down1.0.0 reads [inc | zeros] as 128 channels)
    const int c1 = base_ch, c2 = 2 * base_ch, c4 = 4 * base_ch, c8 = 8 * base_ch;
    const bool b96 = base_ch == 96;
    const int cat1p = b96 ? 224 : 2 * c1;
    __half *cat1, *d1a, *cat2, *d2a, *cat3, *d3a, *e4, *c3a, *c3b, *c2a, *c2b, *c1a;
    struct { __half** p; size_t elems; } bufs[] = {
        {&h->xin16, N * P0 * 16}, {&cat1, N * P0 * cat1p}, {&d1a, N * P0 * c2}, {&cat2, N * P1 * c4},
        {&d2a, N * P1 * c4},      {&cat3, N * P2 * c8},    {&d3a, N * P2 * c8}, {&e4, N * P3 * c8},
        {&c3a, N * P2 * c4},      {&c3b, N * P2 * c4},     {&c2a, N * P1 * c2}, {&c2b, N * P1 * c2},
        {&c1a, N * P0 * c1}};
    for (auto& b : bufs) {
        void* p = nullptr;
        if (dmalloc(h, &p, b.elems * sizeof(__half), err) != S1S2_OK) {
            s1s2_destroy(h);
            return S1S2_ERR_CUDA;
        }
        cudaMemset(p, 0, b.elems * sizeof(__half));
        *b.p = static_cast<__half*>(p);
    }
    {
        void* p = nullptr;
        if (dmalloc(h, &p, sizeof(uint32_t) * 2 * max_batch, err) != S1S2_OK) { s1s2_destroy(h); return S1S2_ERR_CUDA; }
        cudaMemset(p, 0, sizeof(uint32_t) * 2 * max_batch);
        h->amax = static_cast<uint32_t*>(p);
    }
    {
        const size_t img = static_cast<size_t>(max_batch) * 4 * P0 * sizeof(float);
        for (int k = 0; k < 2; ++k) {
            float** slots[3] = {&h->st_cond[k], &h->st_x[k], &h->st_out[k]};
            for (float** slot : slots) {
                void* p;
                if (dmalloc(h, &p, img, err) != S1S2_OK) { s1s2_destroy(h); return S1S2_ERR_CUDA; }
                *slot = static_cast<float*>(p);
            }
        }
        void* p;
        if (dmalloc(h, &p, sizeof(unsigned long long) * 32, err) != S1S2_OK) { s1s2_destroy(h); return S1S2_ERR_CUDA; }
        h->sat_counts = static_cast<unsigned long long*>(p);
        bool ok = cudaStreamCreateWithFlags(&h->s_in, cudaStreamNonBlocking) == cudaSuccess &&
                  cudaStreamCreateWithFlags(&h->s_out, cudaStreamNonBlocking) == cudaSuccess;
        ok = ok && cudaEventCreateWithFlags(&h->ev_entry, cudaEventDisableTiming) == cudaSuccess;
        for (int k = 0; k < 2 && ok; ++k)
            for (cudaEvent_t* e : {&h->ev_in[k], &h->ev_packed[k], &h->ev_done[k], &h->ev_out[k]})
                ok = ok && cudaEventCreateWithFlags(e, cudaEventDisableTiming) == cudaSuccess;
        if (!ok) {
            set_err(err, "stream / event creation failed: %s", cudaGetErrorString(cudaGetLastError()));
            s1s2_destroy(h);
            return S1S2_ERR_CUDA;
        }
    }

    auto add = [&](const char* name, KernelId kid, int level, int cin, int ntot, int cout, int taps, const __half* src,
                   int sp, __half* dst, int dp) {
        Layer L;
        L.name = name; L.kid = kid; L.level = level; L.cin = cin; L.ntot = ntot; L.cout = cout; L.taps_w = taps;
        L.src = src; L.src_pitch = sp; L.dst = dst; L.dst_pitch = dp;
        h->layers.push_back(L);
    };
    // Kernel of a 3x3 layer by its output width (non-halo ids: the halo / pixels-on-N substitutions follow below).
    // base_ch = 96: N = 192 tiles for Cout = 192 / 384, N = 256 for 768.  base_ch = 64: N = 128 for Cout = 128, N = 256 for 256 / 512.
    const KernelId kS2 = b96 ? K_STORE : K_HSTORE128, kP2 = b96 ? K_POOL : K_HPOOL128;          // Cout = 2 * base_ch
    const KernelId kS4 = b96 ? K_STORE : K_HSTORE256, kP4 = b96 ? K_POOL : K_HPOOL256;          // Cout = 4 * base_ch
    const KernelId kS8 = b96 ? K_STORE256 : K_HSTORE256, kP8 = b96 ? K_POOL256 : K_HPOOL256;    // Cout = 8 * base_ch
    //   name          kernel    lvl cin  N       cout taps src          pitch  dst           pitch
    add("inc.0",       b96 ? K_INC : K_HINC64, 0, 16, c1, c1, 3, h->xin16, 16,  cat1 + c1,    cat1p);
    h->layers.back().first = true;
    if (!b96) {
        add("down1.0.0", kS2,      0,  c1,  c2,     c2,  3,   cat1 + c1,   cat1p, d1a,          c2);
    } else if (getenv("S1S2_PAD96") != nullptr) {    // A/B: K padded to 128 per tap (two 64-channel chunks, a quarter of the MMAs wasted)
        add("down1.0.0", K_STORE,  0,  128, c2,     c2,  3,   cat1 + c1,   cat1p, d1a,          c2);
        h->layers.back().cin_real = 96;
    } else {                                   // exact K = 96 per tap as three 32-channel chunks
        add("down1.0.0", K_C96IN,  0,  c1,  c2,     c2,  3,   cat1 + c1,   cat1p, d1a,          c2);
    }
    add("down1.0.2",   kP2,      0,  c2,  c2,     c2,  3,   d1a,         c2,    cat2 + c2,    c4);
    add("down2.0.0",   kS4,      1,  c2,  c4,     c4,  3,   cat2 + c2,   c4,    d2a,          c4);
    add("down2.0.2",   kP4,      1,  c4,  c4,     c4,  3,   d2a,         c4,    cat3 + c4,    c8);
    add("down3.0.0",   kS8,      2,  c4,  c8,     c8,  3,   cat3 + c4,   c8,    d3a,          c8);
    add("down3.0.2",   kP8,      2,  c8,  c8,     c8,  3,   d3a,         c8,    e4,           c8);
    add("up3",         K_CONVT256, 3, c8, 4 * c4, c4,  1,   e4,          c8,    cat3,         c8);
    add("conv3.0",     kS4,      2,  c8,  c4,     c4,  3,   cat3,        c8,    c3a,          c4);
    add("conv3.2",     kS4,      2,  c4,  c4,     c4,  3,   c3a,         c4,    c3b,          c4);
    add("up2",         K_CONVT256, 2, c4, 4 * c2, c2,  1,   c3b,         c4,    cat2,         c4);
    add("conv2.0",     kS2,      1,  c4,  c2,     c2,  3,   cat2,        c4,    c2a,          c2);
    add("conv2.2",     kS2,      1,  c2,  c2,     c2,  3,   c2a,         c2,    c2b,          c2);
    add("up1",         b96 ? K_CONVT : K_CONVT256, 1, c2, 4 * c1, c1, 1, c2b, c2,    cat1,         cat1p);
    add("conv1.0",     b96 ? K_N96 : K_HSTORE64, 0, c2, c1, c1,  3,   cat1,        cat1p, c1a,          c1);
    add("conv1.2",     b96 ? K_HEAD : K_PX_HEAD64, 0, c1, c1, c1, 3,  c1a,         c1,    nullptr,      0);

    if (b96 && getenv("S1S2_NO_HALO") == nullptr) {
        for (Layer& L : h->layers) {
            if (L.taps_w != 3) continue;
            if (L.kid == K_INC) L.kid = K_HINC;
            else if (L.kid == K_C96IN) L.kid = K_HC96IN;
            if (L.cin % 64 != 0) continue;
            if (L.kid == K_STORE) L.kid = K_HSTORE;
            else if (L.kid == K_POOL) L.kid = K_HPOOL;
            else if (L.kid == K_STORE256) L.kid = K_HSTORE256;
            else if (L.kid == K_POOL256) L.kid = K_HPOOL256;
        }
    }
    if (b96 && getenv("S1S2_NO_PX") == nullptr) {       // default: pixels-on-N kernel for the head layer
        for (Layer& L : h->layers) {
            // conv1.0: Cout on N with 96-column tiles, three taps per stage (measured 8 % faster than the pixels-on-N kernel,
            // whose M = 128 MMAs carry 96 real rows; S1S2_C10_PX=1 restores that one for A/B).  conv1.2 + head: pixels on N
            // (the Cout-on-N head, S1S2_C12_UMMA=1, is 20 % slower: one epilogue warpgroup does the 96 x 4 head FMAs per pixel).
            if (L.kid == K_N96) L.kid = getenv("S1S2_C10_PX") != nullptr ? K_PX_STORE : K_HSTORE96;
            if (L.kid == K_HEAD) L.kid = getenv("S1S2_C12_UMMA") != nullptr ? K_HHEAD96 : K_PX_HEAD32;
        }
    }
    if (const char* force = getenv("S1S2_FORCE")) {        // measurement aid: "layer=KERNEL[,layer=KERNEL...]", e.g. conv2.2=HC96IN
        static const struct { const char* name; KernelId kid; } names[] = {
            {"HSTORE", K_HSTORE}, {"HPOOL", K_HPOOL}, {"HSTORE256", K_HSTORE256}, {"HPOOL256", K_HPOOL256}, {"HC96IN", K_HC96IN},
            {"HSTORE96", K_HSTORE96}, {"HPOOL96", K_HPOOL96}, {"PX_STORE", K_PX_STORE}, {"STORE", K_STORE}, {"N96", K_N96},
            {"HSTORE128", K_HSTORE128}, {"HPOOL128", K_HPOOL128}, {"HSTORE64", K_HSTORE64}};
        std::string spec = force;
        for (Layer& L : h->layers)
            for (const auto& nm : names) {
                const std::string key = std::string(L.name) + "=" + nm.name;
                const size_t at = spec.find(key);
                if (at != std::string::npos && (at + key.size() == spec.size() || spec[at + key.size()] == ',')) L.kid = nm.kid;
            }
    }
    if (getenv("S1S2_NO_ALTS") == nullptr) {     // narrower tilings for small batches (pick_variant), widest first
        for (Layer& L : h->layers) {
            auto alt = [&](KernelId kid) { if (L.ntot % kernel_table()[kid].block_n == 0) L.alts.push_back({kid, ConvParams()}); };
            if (b96) {
                if (L.kid == K_HSTORE256) { alt(K_HSTORE); alt(K_HSTORE96); }
                else if (L.kid == K_HPOOL256) { alt(K_HPOOL); alt(K_HPOOL96); }
                else if (L.kid == K_HSTORE) alt(K_HSTORE96);
                else if (L.kid == K_HPOOL) alt(K_HPOOL96);
            } else {
                if (L.kid == K_HSTORE256) { alt(K_HSTORE128); alt(K_HSTORE64); }
                else if (L.kid == K_HPOOL256) alt(K_HPOOL128);
                else if (L.kid == K_HSTORE128) alt(K_HSTORE64);
            }
        }
    }
    for (Layer& L : h->layers) {
        const size_t welems = static_cast<size_t>(L.ntot) * L.taps_w * L.taps_w * L.cin;
        void* p;
        if (dmalloc(h, &p, welems * sizeof(__half), err) != S1S2_OK) { s1s2_destroy(h); return S1S2_ERR_CUDA; }
        L.w = static_cast<__half*>(p);
        L.w_bytes = welems * sizeof(__half);
        if (dmalloc(h, &p, static_cast<size_t>(L.ntot) * sizeof(float), err) != S1S2_OK) { s1s2_destroy(h); return S1S2_ERR_CUDA; }
        L.bias = static_cast<float*>(p);
        int rc = build_layer_params(h, L, L.kid, L.p, err);
        for (Layer::Alt& a : L.alts)
            if (rc == S1S2_OK) rc = build_layer_params(h, L, a.kid, a.p, err);
        if (rc != S1S2_OK) { s1s2_destroy(h); return rc; }
    }
    // views for s1s2_debug_activation: name = the oracle's tap key
    h->views = {{"inc", cat1 + c1, cat1p, c1, 0},    {"down1.0", d1a, c2, c2, 0},   {"down1", cat2 + c2, c4, c2, 1},
                {"down2.0", d2a, c4, c4, 1},         {"down2", cat3 + c4, c8, c4, 2}, {"down3.0", d3a, c8, c8, 2},
                {"down3", e4, c8, c8, 3},            {"up3", cat3, c8, c4, 2},      {"conv3.0", c3a, c4, c4, 2},
                {"conv3", c3b, c4, c4, 2},           {"up2", cat2, c4, c2, 1},      {"conv2.0", c2a, c2, c2, 1},
                {"conv2", c2b, c2, c2, 1},           {"up1", cat1, cat1p, c1, 0},   {"conv1.0", c1a, c1, c1, 0},
                {"xin16", h->xin16, 16, 16, 0}};
    {   // first use of the device's stream-ordered memory pool costs ~0.1 s (s1s2_stitch allocates its scratch from it):
        // pay it here, at model creation, not inside the first stitched scene
        void* scratch = nullptr;
        if (cudaMallocAsync(&scratch, 256, nullptr) == cudaSuccess) cudaFreeAsync(scratch, nullptr);
    }
    if (cudaDeviceSynchronize() != cudaSuccess) {
        set_err(err, "arena initialisation failed: %s", cudaGetErrorString(cudaGetLastError()));
        s1s2_destroy(h);
        return S1S2_ERR_CUDA;
    }
    *out = h;
    return S1S2_OK;
}

void s1s2_destroy(s1s2_handle* h) {
    if (h == nullptr) return;
    cudaSetDevice(h->device);
    cudaDeviceSynchronize();
    for (void* p : h->allocs) cudaFree(p);
    for (int k = 0; k < 2; ++k)
        for (cudaEvent_t e : {h->ev_in[k], h->ev_packed[k], h->ev_done[k], h->ev_out[k]})
            if (e != nullptr) cudaEventDestroy(e);
    if (h->ev_entry != nullptr) cudaEventDestroy(h->ev_entry);
    if (h->s_in != nullptr) cudaStreamDestroy(h->s_in);
    if (h->s_out != nullptr) cudaStreamDestroy(h->s_out);
    delete h;
}

int s1s2_load_weights(s1s2_handle* h, int n, const char* const* names, const float* const* ptrs, const int64_t* numel,
                      void* stream) {
    if (h == nullptr) return S1S2_ERR_INVALID;
    std::string* err = &h->err;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    CK(cudaSetDevice(h->device));
    auto find = [&](const std::string& key, int64_t want, const float** out) -> int {
        for (int i = 0; i < n; ++i) {
            if (key == names[i]) {
                if (numel[i] != want) {
                    set_err(err, "size mismatch for %s: got %lld elements, expected %lld", key.c_str(),
                            static_cast<long long>(numel[i]), static_cast<long long>(want));
                    return S1S2_ERR_INVALID;
                }
                *out = ptrs[i];
                return S1S2_OK;
            }
        }
        set_err(err, "missing key in state_dict: %s", key.c_str());
        return S1S2_ERR_INVALID;
    };
    if (n != 34) {
        set_err(err, "strict load: expected the 34 tensors of UNetSmall(8,4,%d), got %d", h->base_ch, n);
        return S1S2_ERR_INVALID;
    }
    for (Layer& L : h->layers) {
        const float *w = nullptr, *b = nullptr;
        const std::string nm = L.name;
        int rc;
        if (L.first) {
            if ((rc = find(nm + ".weight", static_cast<int64_t>(L.cout) * 9 * 9, &w)) || (rc = find(nm + ".bias", L.cout, &b))) return rc;
            repack_inc_kernel<<<64, 256, 0, st>>>(w, L.w, L.cout);
            tile_bias_kernel<<<4, 256, 0, st>>>(b, L.bias, L.cout, 1);
        } else if (kernel_table()[L.kid].mode == MODE_CONVT) {
            if ((rc = find(nm + ".weight", static_cast<int64_t>(L.cin) * L.cout * 4, &w)) ||
                (rc = find(nm + ".bias", L.cout, &b)))
                return rc;
            repack_convt_kernel<<<512, 256, 0, st>>>(w, L.w, L.cin, L.cout);
            tile_bias_kernel<<<8, 256, 0, st>>>(b, L.bias, L.cout, 4);
        } else {
            const int cin_real = L.cin_real > 0 ? L.cin_real : L.cin;
            if ((rc = find(nm + ".weight", static_cast<int64_t>(L.cout) * cin_real * 9, &w)) ||
                (rc = find(nm + ".bias", L.cout, &b)))
                return rc;
            repack_conv3_kernel<<<1024, 256, 0, st>>>(w, L.w, L.cout, cin_real, L.cin);
            tile_bias_kernel<<<8, 256, 0, st>>>(b, L.bias, L.cout, 1);
        }
        h->launches += 2;
        CK(cudaGetLastError());
    }
    {
        const float *w = nullptr, *b = nullptr;
        int rc;
        if ((rc = find("outc.weight", kHeadOut * h->base_ch, &w)) || (rc = find("outc.bias", kHeadOut, &b))) return rc;
        memset(h->head_w, 0, sizeof(h->head_w));            // [4][base_ch] contiguous, the layout the head kernels index
        CK(cudaMemcpyAsync(h->head_w, w, sizeof(float) * kHeadOut * h->base_ch, cudaMemcpyDeviceToHost, st));
        CK(cudaMemcpyAsync(h->head_b, b, sizeof(h->head_b), cudaMemcpyDeviceToHost, st));
    }
    CK(cudaStreamSynchronize(st));
    h->weights_loaded = true;
    return S1S2_OK;
}

int s1s2_forward(s1s2_handle* h, const float* xt_and_cond, const int64_t* t_idx, float* out, int B, void* stream) {
    int rc = check_batch(h, B);
    if (rc != S1S2_OK) return rc;
    std::string* err = &h->err;
    if (xt_and_cond == nullptr || t_idx == nullptr || out == nullptr) {
        h->err = "s1s2_forward: null pointer argument";
        return S1S2_ERR_INVALID;
    }
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    CK(cudaSetDevice(h->device));
    const int HW = h->H * h->W;
    const size_t total = static_cast<size_t>(B) * HW;
    CK(cudaMemsetAsync(h->amax, 0, sizeof(uint32_t) * B, st));
    pack_input_kernel<<<static_cast<unsigned>((total + 255) / 256), 256, 0, st>>>(
        xt_and_cond, static_cast<size_t>(8) * HW, xt_and_cond + static_cast<size_t>(4) * HW, static_cast<size_t>(8) * HW,
        t_idx, 0.f, 1.f, nullptr, h->xin16, h->amax, HW, total);
    CK(cudaGetLastError());
    ++h->launches;
    HeadParams io;
    memset(&io, 0, sizeof(io));
    io.pred_out = out;
    io.step.kind = STEP_NONE;
    return run_network(h, B, io, h->amax, nullptr, st, err);
}

static int sample_impl(s1s2_handle* h, const s1s2_step* steps, int n_steps, const float* cond, const float* x_init,
                       float init_scale, const float* step_noise, float* out, float* tap_pred, float* tap_x, int B,
                       void* stream, cudaEvent_t after_pack) {
    int rc = check_batch(h, B);
    if (rc != S1S2_OK) return rc;
    std::string* err = &h->err;
    if (steps == nullptr || n_steps < 1 || cond == nullptr || x_init == nullptr || out == nullptr) {
        h->err = "s1s2_sample: null pointer / empty step list";
        return S1S2_ERR_INVALID;
    }
    for (int i = 0; i < n_steps; ++i) {
        if (steps[i].t < 0 || steps[i].t > 2048) {
            set_err(err, "step %d: timestep %d outside [0, 2048] (fp16-exact range of the time planes)", i, steps[i].t);
            return S1S2_ERR_INVALID;
        }
        if (steps[i].kind < S1S2_STEP_EPS_DDIM || steps[i].kind > S1S2_STEP_V_DDPM) {
            set_err(err, "step %d: unknown scheduler kind %d", i, steps[i].kind);
            return S1S2_ERR_INVALID;
        }
        if ((steps[i].flags & S1S2_STEP_NOISE) && (step_noise == nullptr || steps[i].noise_index < 0)) {
            set_err(err, "step %d asks for noise but step_noise is NULL / noise_index < 0", i);
            return S1S2_ERR_INVALID;
        }
        if ((steps[i].flags & S1S2_STEP_PHILOX) && ((steps[i].flags & S1S2_STEP_NOISE) || steps[i].noise_index < 0 ||
                                                    !kernel_table()[h->layers.back().kid].px)) {
            set_err(err, "step %d: in-kernel noise needs noise_index >= 0, no S1S2_STEP_NOISE, and the default head kernel", i);
            return S1S2_ERR_INVALID;
        }
    }
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    CK(cudaSetDevice(h->device));
    const int HW = h->H * h->W;
    const size_t total = static_cast<size_t>(B) * HW;
    const size_t img = static_cast<size_t>(B) * 4 * HW;
    CK(cudaMemsetAsync(h->amax, 0, sizeof(uint32_t) * B, st));
    pack_input_kernel<<<static_cast<unsigned>((total + 255) / 256), 256, 0, st>>>(
        x_init, static_cast<size_t>(4) * HW, cond, static_cast<size_t>(4) * HW, nullptr, static_cast<float>(steps[0].t),
        init_scale, out, h->xin16, h->amax, HW, total);
    CK(cudaGetLastError());
    ++h->launches;
    if (after_pack != nullptr) CK(cudaEventRecord(after_pack, st));       // cond / x_init are not read past this point
    for (int i = 0; i < n_steps; ++i) {
        uint32_t* amax_cur = h->amax + static_cast<size_t>(i & 1) * h->max_batch;
        uint32_t* amax_nxt = h->amax + static_cast<size_t>((i + 1) & 1) * h->max_batch;
        HeadParams io;
        memset(&io, 0, sizeof(io));
        io.amax_out = amax_nxt;
        io.x_t = out;
        io.pred_out = tap_pred != nullptr ? tap_pred + static_cast<size_t>(i) * img : nullptr;
        io.noise = (steps[i].flags & S1S2_STEP_NOISE) ? step_noise + static_cast<size_t>(steps[i].noise_index) * img : nullptr;
        io.seed_lo = static_cast<uint32_t>(h->noise_seed);
        io.seed_hi = static_cast<uint32_t>(h->noise_seed >> 32);
        io.noise_stream = static_cast<uint32_t>(steps[i].noise_index < 0 ? 0 : steps[i].noise_index);
        io.patch_base = h->patch_base;
        io.xin16 = h->xin16;
        io.step.c0 = steps[i].c0;
        io.step.c1 = steps[i].c1;
        io.step.c2 = steps[i].c2;
        io.step.c3 = steps[i].c3;
        io.step.c4 = steps[i].c4;
        io.step.t_next = i + 1 < n_steps ? static_cast<float>(steps[i + 1].t) : 0.f;
        io.step.kind = steps[i].kind;
        io.step.flags = steps[i].flags;
        rc = run_network(h, B, io, amax_cur, amax_nxt, st, err);
        if (rc != S1S2_OK) return rc;
        if (tap_x != nullptr)
            CK(cudaMemcpyAsync(tap_x + static_cast<size_t>(i) * img, out, img * sizeof(float), cudaMemcpyDeviceToDevice, st));
    }
    return S1S2_OK;
}

int s1s2_sample(s1s2_handle* h, const s1s2_step* steps, int n_steps, const float* cond, const float* x_init,
                float init_scale, const float* step_noise, float* out, float* tap_pred, float* tap_x, int B,
                void* stream) {
    return sample_impl(h, steps, n_steps, cond, x_init, init_scale, step_noise, out, tap_pred, tap_x, B, stream, nullptr);
}

int s1s2_set_noise_seed(s1s2_handle* h, uint64_t seed, uint32_t patch_base) {
    if (h == nullptr) return S1S2_ERR_INVALID;
    h->noise_seed = seed;
    h->patch_base = patch_base;
    return S1S2_OK;
}

int s1s2_sample_host(s1s2_handle* h, const s1s2_step* steps, int n_steps, const float* cond_host,
                     const float* x_init_host, float init_scale, float* out_host, int B, void* stream) {
    return s1s2_sample_host_stream(h, steps, n_steps, cond_host, x_init_host, init_scale, out_host, B, B, stream);
}

int s1s2_sample_host_stream(s1s2_handle* h, const s1s2_step* steps, int n_steps, const float* cond_host,
                            const float* x_init_host, float init_scale, float* out_host, int N, int batch, void* stream) {
    int rc = check_batch(h, batch);
    if (rc != S1S2_OK) return rc;
    std::string* err = &h->err;
    if (cond_host == nullptr || x_init_host == nullptr || out_host == nullptr || N < 0) {
        h->err = "s1s2_sample_host_stream: null pointer argument / negative patch count";
        return S1S2_ERR_INVALID;
    }
    if (N == 0) return S1S2_OK;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    CK(cudaSetDevice(h->device));
    const size_t patch = static_cast<size_t>(4) * h->H * h->W;          // floats per patch
    const int nb = (N + batch - 1) / batch;
    // Three streams: uploads of batch i+1 (s_in) and the download of batch i-1 (s_out) run under the model calls of batch
    // i (`stream`); two staging sets.  Set k is free for upload once the pack kernel of its previous occupant has run
    // (cond / x_init are read nowhere else), and free for compute once that occupant's result has been downloaded.
    CK(cudaEventRecord(h->ev_entry, st));
    CK(cudaStreamWaitEvent(h->s_in, h->ev_entry, 0));
    CK(cudaStreamWaitEvent(h->s_out, h->ev_entry, 0));
    for (int i = 0; i < nb; ++i) {
        const int k = i & 1;
        const size_t lo = static_cast<size_t>(i) * batch;
        const int Bi = static_cast<int>(static_cast<size_t>(N) - lo < static_cast<size_t>(batch) ? N - lo : batch);
        const size_t bytes = static_cast<size_t>(Bi) * patch * sizeof(float);
        if (i >= 2) CK(cudaStreamWaitEvent(h->s_in, h->ev_packed[k], 0));
        CK(cudaMemcpyAsync(h->st_cond[k], cond_host + lo * patch, bytes, cudaMemcpyHostToDevice, h->s_in));
        CK(cudaMemcpyAsync(h->st_x[k], x_init_host + lo * patch, bytes, cudaMemcpyHostToDevice, h->s_in));
        CK(cudaEventRecord(h->ev_in[k], h->s_in));
        CK(cudaStreamWaitEvent(st, h->ev_in[k], 0));
        if (i >= 2) CK(cudaStreamWaitEvent(st, h->ev_out[k], 0));
        rc = sample_impl(h, steps, n_steps, h->st_cond[k], h->st_x[k], init_scale, nullptr, h->st_out[k], nullptr, nullptr, Bi,
                         stream, h->ev_packed[k]);
        if (rc != S1S2_OK) break;
        CK(cudaEventRecord(h->ev_done[k], st));
        CK(cudaStreamWaitEvent(h->s_out, h->ev_done[k], 0));
        CK(cudaMemcpyAsync(out_host + lo * patch, h->st_out[k], bytes, cudaMemcpyDeviceToHost, h->s_out));
        CK(cudaEventRecord(h->ev_out[k], h->s_out));
    }
    // whatever happened above, leave no copy in flight on the side streams when returning
    cudaError_t e1 = cudaStreamSynchronize(h->s_in), e2 = cudaStreamSynchronize(st), e3 = cudaStreamSynchronize(h->s_out);
    if (rc != S1S2_OK) return rc;
    for (cudaError_t e : {e1, e2, e3})
        if (e != cudaSuccess) {
            set_err(err, "s1s2_sample_host_stream: %s", cudaGetErrorString(e));
            return S1S2_ERR_CUDA;
        }
    return S1S2_OK;
}

int s1s2_debug_activation(s1s2_handle* h, const char* name, float* out_nchw, int B, int* C, int* Hout, int* Wout,
                          void* stream) {
    if (h == nullptr || name == nullptr) return S1S2_ERR_INVALID;
    std::string* err = &h->err;
    for (const View& v : h->views) {
        if (strcmp(v.name, name) == 0) {
            const int Hl = h->H >> v.level, Wl = h->W >> v.level;
            if (C != nullptr) *C = v.C;
            if (Hout != nullptr) *Hout = Hl;
            if (Wout != nullptr) *Wout = Wl;
            if (out_nchw == nullptr) return S1S2_OK;   // shape query
            if (B < 1 || B > h->max_batch) return S1S2_ERR_INVALID;
            CK(cudaSetDevice(h->device));
            const size_t total = static_cast<size_t>(B) * v.C * Hl * Wl;
            unpack_activation_kernel<<<static_cast<unsigned>((total + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
                v.ptr, v.pitch, v.C, Hl * Wl, out_nchw, total);
            CK(cudaGetLastError());
            ++h->launches;
            return S1S2_OK;
        }
    }
    set_err(err, "unknown activation '%s'", name);
    return S1S2_ERR_INVALID;
}

int s1s2_debug_tile_width(s1s2_handle* h, int layer, int B) {
    if (h == nullptr || layer < 0 || layer >= static_cast<int>(h->layers.size()) || B < 1 || B > h->max_batch) return -1;
    const Layer& L = h->layers[layer];
    const Layer::Alt* alt = pick_variant(h, L, B);
    return kernel_table()[alt != nullptr ? alt->kid : L.kid].block_n;
}

const char* s1s2_layer_name(const s1s2_handle* h, int i) {
    if (h == nullptr || i < 0 || i >= static_cast<int>(h->layers.size())) return nullptr;
    return h->layers[i].name;
}

int s1s2_debug_loop_layer(s1s2_handle* h, int B, int layer, int reps, int perf_mode, float* ms_out, void* stream) {
    int rc = check_batch(h, B);
    if (rc != S1S2_OK) return rc;
    std::string* err = &h->err;
    if (layer < 0 || layer >= static_cast<int>(h->layers.size()) || reps < 1 || ms_out == nullptr) {
        set_err(err, "s1s2_debug_loop_layer: layer %d, reps %d", layer, reps);
        return S1S2_ERR_INVALID;
    }
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    CK(cudaSetDevice(h->device));
    Layer& L = h->layers[layer];
    if (kernel_table()[L.kid].mode == MODE_HEAD) {
        HeadParams& hp = L.p.head;
        memcpy(hp.w, h->head_w, sizeof(hp.w));
        memcpy(hp.b, h->head_b, sizeof(hp.b));
        hp.x_t = nullptr; hp.pred_out = nullptr; hp.noise = nullptr; hp.xin16 = nullptr; hp.amax_out = nullptr;
        hp.step.kind = STEP_NONE;
    }
    EventSet ev;                                   // destroyed on every return path
    if ((rc = ev.create(2, err)) != S1S2_OK) return rc;
    L.p.perf_mode = perf_mode;
    rc = launch_layer(h, L, B, nullptr, nullptr, st, err);      // warm-up
    cudaError_t ce = cudaEventRecord(ev[0], st);
    for (int r = 0; r < reps && rc == S1S2_OK; ++r) rc = launch_layer(h, L, B, nullptr, nullptr, st, err);
    if (ce == cudaSuccess) ce = cudaEventRecord(ev[1], st);
    L.p.perf_mode = 0;
    if (rc != S1S2_OK) return rc;
    CK(ce);
    CK(cudaStreamSynchronize(st));
    float ms = 0.f;
    CK(cudaEventElapsedTime(&ms, ev[0], ev[1]));
    *ms_out = ms / reps;
    return S1S2_OK;
}

int s1s2_profile_layers(s1s2_handle* h, int B, int reps, float* ms_out, int n_out, int* n_layers, void* stream) {
    int rc = check_batch(h, B);
    if (rc != S1S2_OK) return rc;
    std::string* err = &h->err;
    const int nl = static_cast<int>(h->layers.size());
    if (n_layers != nullptr) *n_layers = nl;
    if (ms_out == nullptr) return S1S2_OK;
    if (reps < 1 || n_out < nl) {
        set_err(err, "s1s2_profile_layers: reps %d, n_out %d (need >= %d)", reps, n_out, nl);
        return S1S2_ERR_INVALID;
    }
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    CK(cudaSetDevice(h->device));
    EventSet ev;                                   // destroyed on every return path
    if ((rc = ev.create(reps * (nl + 1), err)) != S1S2_OK) return rc;
    HeadParams io;
    memset(&io, 0, sizeof(io));
    io.step.kind = STEP_NONE;
    for (int r = 0; r < reps; ++r) {
        for (int i = 0; i < nl; ++i) {
            Layer& L = h->layers[i];
            if (kernel_table()[L.kid].mode == MODE_HEAD) {
                HeadParams& hp = L.p.head;
                memcpy(hp.w, h->head_w, sizeof(hp.w));
                memcpy(hp.b, h->head_b, sizeof(hp.b));
                hp.x_t = nullptr; hp.pred_out = nullptr; hp.noise = nullptr; hp.xin16 = nullptr; hp.amax_out = nullptr;
                hp.step = io.step;
            }
            CK(cudaEventRecord(ev[r * (nl + 1) + i], st));
            rc = launch_layer(h, L, B, nullptr, nullptr, st, err);
            if (rc != S1S2_OK) return rc;
        }
        CK(cudaEventRecord(ev[r * (nl + 1) + nl], st));
    }
    CK(cudaStreamSynchronize(st));
    for (int i = 0; i < nl; ++i) {
        double acc = 0.0;
        for (int r = 0; r < reps; ++r) {
            float ms = 0.f;
            CK(cudaEventElapsedTime(&ms, ev[r * (nl + 1) + i], ev[r * (nl + 1) + i + 1]));
            acc += ms;
        }
        ms_out[i] = static_cast<float>(acc / reps);
    }
    return S1S2_OK;
}

int s1s2_tile_extract(int device, const float* scene, const uint8_t* vmask, int SH, int SW, const int32_t* origins, int N,
                      int ps, float* cond, uint8_t* mask, float* valid_ratio, void* stream) {
    std::string* err = &g_error;
    if (N == 0) return S1S2_OK;
    if (scene == nullptr || origins == nullptr || cond == nullptr || mask == nullptr || N < 0 || ps < 1 || ps > SH || ps > SW) {
        set_err(err, "s1s2_tile_extract: bad argument (N %d, ps %d, scene %d x %d)", N, ps, SH, SW);
        return S1S2_ERR_INVALID;
    }
    if (N == 0) return S1S2_OK;
    CK(cudaSetDevice(device));
    const int allow_vec = (SW % 4 == 0 && ps % 4 == 0 && aligned_to(scene, 16) && aligned_to(cond, 16) && aligned_to(mask, 4) &&
                           (vmask == nullptr || aligned_to(vmask, 4))) ? 1 : 0;
    tile_extract_kernel<<<N, kExtractThreads, 0, static_cast<cudaStream_t>(stream)>>>(scene, vmask, SH, SW, origins, ps, cond,
                                                                                     mask, valid_ratio, allow_vec);
    CK(cudaGetLastError());
    return S1S2_OK;
}

int s1s2_tile_filter(int device, const float* scene, int Ci, const float* target, const uint8_t* colloc, int SH, int SW,
                     const int32_t* origins, int N, int ps, const float* thresholds, float* stats, void* stream) {
    std::string* err = &g_error;
    if (N == 0) return S1S2_OK;
    if (scene == nullptr || target == nullptr || origins == nullptr || thresholds == nullptr || stats == nullptr || N < 0 ||
        Ci < 1 || Ci > kFilterMaxCi || ps < 1 || ps > SH || ps > SW) {
        set_err(err, "s1s2_tile_filter: bad argument (N %d, Ci %d, ps %d, scene %d x %d)", N, Ci, ps, SH, SW);
        return S1S2_ERR_INVALID;
    }
    CK(cudaSetDevice(device));
    FilterThresholds th{thresholds[0], thresholds[1], thresholds[2], thresholds[3], thresholds[4]};
    const int allow_vec = (SW % 4 == 0 && ps % 4 == 0 && aligned_to(scene, 16) && aligned_to(target, 16) &&
                           (colloc == nullptr || aligned_to(colloc, 4))) ? 1 : 0;
    tile_filter_kernel<<<N, kFilterThreads, 0, static_cast<cudaStream_t>(stream)>>>(scene, Ci, target, colloc, SH, SW, origins, ps,
                                                                                   th, stats, allow_vec);
    CK(cudaGetLastError());
    return S1S2_OK;
}

int s1s2_patch_metrics(int device, const float* pred, const float* gt, const uint8_t* mask, int N, int C, int HW,
                       double* out, void* stream) {
    std::string* err = &g_error;
    if (N == 0) return S1S2_OK;
    if (pred == nullptr || gt == nullptr || out == nullptr || N < 0 || C < 1 || C > kMetricsMaxC || HW < 1) {
        set_err(err, "s1s2_patch_metrics: bad argument (N %d, C %d, HW %d)", N, C, HW);
        return S1S2_ERR_INVALID;
    }
    CK(cudaSetDevice(device));
    const bool vec = HW % 4 == 0 && aligned_to(pred, 16) && aligned_to(gt, 16) && (mask == nullptr || aligned_to(mask, 4));
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const bool stream_ok = vec && C <= kMsC && HW % kMsChunk == 0 && (mask == nullptr || aligned_to(mask, 16)) &&
                           getenv("S1S2_METRICS_REGPATH") == nullptr;
    if (stream_ok) {                   // bulk-copy ring + 16 consumer warps (the usual 256 x 256 x 4 geometry)
        {
            DeviceOnce& once = device_once();
            std::lock_guard<std::mutex> lock(once.mu);
            if (device < 0 || device >= 64 || !once.metrics_attr[device]) {
                CK(cudaFuncSetAttribute(reinterpret_cast<const void*>(patch_metrics_stream_kernel),
                                        cudaFuncAttributeMaxDynamicSharedMemorySize, kMsSmemBytes));
                if (device >= 0 && device < 64) once.metrics_attr[device] = true;
            }
        }
        patch_metrics_stream_kernel<<<N, kMsThreads, kMsSmemBytes, st>>>(pred, gt, mask, C, HW, out);
    } else if (vec && C <= 4) patch_metrics_kernel<4, 4><<<N, kMetricsThreads, 0, st>>>(pred, gt, mask, C, HW, out);
    else if (vec) patch_metrics_kernel<4, kMetricsMaxC><<<N, kMetricsThreads, 0, st>>>(pred, gt, mask, C, HW, out);
    else patch_metrics_kernel<1, kMetricsMaxC><<<N, kMetricsThreads, 0, st>>>(pred, gt, mask, C, HW, out);
    CK(cudaGetLastError());
    return S1S2_OK;
}

int s1s2_stitch(int device, const float* preds, const int32_t* origins, int N, int C, int ps, int stride, int SH, int SW,
                float* canvas, uint8_t* cover, void* stream) {
    return s1s2_stitch_weighted(device, preds, origins, N, C, ps, stride, SH, SW, nullptr, canvas, cover, stream);
}

int s1s2_stitch_weighted(int device, const float* preds, const int32_t* origins, int N, int C, int ps, int stride, int SH,
                         int SW, const float* window, float* canvas, uint8_t* cover, void* stream) {
    std::string* err = &g_error;
    if (preds == nullptr || origins == nullptr || canvas == nullptr || cover == nullptr || N < 0 || C < 1 || C > kStitchMaxC ||
        ps < 1 || stride < 1 || ps > SH || ps > SW) {
        set_err(err, "s1s2_stitch: bad argument (N %d, C %d, ps %d, stride %d, scene %d x %d)", N, C, ps, stride, SH, SW);
        return S1S2_ERR_INVALID;
    }
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    CK(cudaSetDevice(device));
    const int nrows = (SH - ps) / stride + 1, ncols = (SW - ps) / stride + 1;
    {   // keep the stream-ordered pool's memory across synchronisations (the default threshold of 0 hands it back to the
        // driver at every sync, which turns the small scratch allocation below into a ~0.5 ms driver call per stitch)
        DeviceOnce& once = device_once();
        std::lock_guard<std::mutex> lock(once.mu);
        if (device >= 0 && device < 64 && !once.pool_keep[device]) {
            cudaMemPool_t pool;
            if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess) {
                uint64_t keep = 64ull << 20;
                cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
            }
            once.pool_keep[device] = true;
        }
    }
    int32_t* grid_map = nullptr;
    CK(cudaMallocAsync(reinterpret_cast<void**>(&grid_map), sizeof(int32_t) * nrows * ncols, st));
    cudaError_t e = cudaMemsetAsync(grid_map, 0xFF, sizeof(int32_t) * nrows * ncols, st);
    if (e == cudaSuccess && N > 0) {
        stitch_map_kernel<<<(N + 255) / 256, 256, 0, st>>>(origins, N, stride, nrows, ncols, grid_map);
        e = cudaGetLastError();
    }
    if (e == cudaSuccess) {
        const unsigned gy = static_cast<unsigned>(SH < 65535 ? SH : 65535);       // rows beyond that: the kernel strides over y
        if (SW % 4 == 0 && ps % 4 == 0 && stride % 4 == 0 && aligned_to(preds, 16) && aligned_to(canvas, 16) && aligned_to(cover, 4)) {
            dim3 grid((SW / 4 + 127) / 128, gy);
            if (C <= 4) stitch_gather_kernel<4, 4><<<grid, 128, 0, st>>>(preds, grid_map, C, ps, stride, nrows, ncols, SH, SW, canvas, cover, window);
            else stitch_gather_kernel<4, kStitchMaxC><<<grid, 128, 0, st>>>(preds, grid_map, C, ps, stride, nrows, ncols, SH, SW, canvas, cover, window);
        } else {
            dim3 grid((SW + 127) / 128, gy);
            stitch_gather_kernel<1, kStitchMaxC><<<grid, 128, 0, st>>>(preds, grid_map, C, ps, stride, nrows, ncols, SH, SW, canvas, cover, window);
        }
        e = cudaGetLastError();
    }
    const cudaError_t ef = cudaFreeAsync(grid_map, st);        // on every path: stream-ordered, after whatever was enqueued
    CK(e);
    CK(ef);
    return S1S2_OK;
}

int s1s2_patch_noise(int device, uint64_t seed, const int64_t* patch_ids, int N, int64_t elems_per_patch, float* out,
                     void* stream) {
    std::string* err = &g_error;
    if (N == 0) return S1S2_OK;
    if (patch_ids == nullptr || out == nullptr || N < 0 || N > 65535 || elems_per_patch < 4 || elems_per_patch % 4 != 0 ||
        !aligned_to(out, 16)) {
        set_err(err, "s1s2_patch_noise: bad argument (N %d (<= 65535 per call), %lld elements per patch (multiple of 4), "
                     "out 16-byte aligned)", N, static_cast<long long>(elems_per_patch));
        return S1S2_ERR_INVALID;
    }
    CK(cudaSetDevice(device));
    const size_t e4 = static_cast<size_t>(elems_per_patch / 4);
    dim3 grid(static_cast<unsigned>((e4 + 255) / 256), static_cast<unsigned>(N));
    patch_noise_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(patch_ids, static_cast<uint32_t>(seed),
                                                                            static_cast<uint32_t>(seed >> 32), e4,
                                                                            reinterpret_cast<float4*>(out));
    CK(cudaGetLastError());
    return S1S2_OK;
}

int s1s2_debug_saturation_count(s1s2_handle* h, int B, uint64_t* counts, int n_out, int* n_views, void* stream) {
    int rc = check_batch(h, B);
    if (rc != S1S2_OK) return rc;
    std::string* err = &h->err;
    const int nv = static_cast<int>(h->views.size());
    if (n_views != nullptr) *n_views = nv;
    if (counts == nullptr) return S1S2_OK;
    if (n_out < nv || nv > 32) {
        set_err(err, "s1s2_debug_saturation_count: n_out %d (need >= %d)", n_out, nv);
        return S1S2_ERR_INVALID;
    }
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    CK(cudaSetDevice(h->device));
    CK(cudaMemsetAsync(h->sat_counts, 0, sizeof(unsigned long long) * nv, st));
    for (int i = 0; i < nv; ++i) {
        const View& v = h->views[i];
        const size_t npix = static_cast<size_t>(B) * (h->H >> v.level) * (h->W >> v.level);
        saturation_count_kernel<<<h->num_sms * 4, 256, 0, st>>>(v.ptr, v.pitch, v.C, npix, h->sat_counts + i);
        CK(cudaGetLastError());
    }
    static_assert(sizeof(unsigned long long) == sizeof(uint64_t), "counter width");
    CK(cudaMemcpyAsync(counts, h->sat_counts, sizeof(uint64_t) * nv, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    return S1S2_OK;
}

#ifdef S1S2_TIMELINE
// Timeline build only (tools/timeline.py; not declared in the public header): runs `calls` model calls back to back
// exactly like s1s2_sample's loop (programmatic dependent launches, no events in between) and returns the %globaltimer
// stamps of the LAST call, out[layers][16] (host).
int s1s2_debug_timeline(s1s2_handle* h, int B, int calls, uint64_t* out, void* stream) {
    int rc = check_batch(h, B);
    if (rc != S1S2_OK) return rc;
    std::string* err = &h->err;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    CK(cudaSetDevice(h->device));
    const size_t n = h->layers.size() * 16;
    unsigned long long* buf = nullptr;
    CK(cudaMalloc(&buf, n * sizeof(unsigned long long)));
    std::vector<unsigned long long> init(n, 0ull);
    for (size_t l = 0; l < h->layers.size(); ++l) init[l * 16 + 12] = ~0ull;
    HeadParams io;
    memset(&io, 0, sizeof(io));
    io.step.kind = STEP_NONE;
    for (int c = 0; c < calls && rc == S1S2_OK; ++c) {
        if (c == calls - 1) {
            cudaMemcpyAsync(buf, init.data(), n * sizeof(unsigned long long), cudaMemcpyHostToDevice, st);
            cudaStreamSynchronize(st);
            h->tl_dbg = buf;
        }
        rc = run_network(h, B, io, nullptr, nullptr, st, err);
    }
    h->tl_dbg = nullptr;
    cudaStreamSynchronize(st);
    if (rc == S1S2_OK) cudaMemcpy(out, buf, n * sizeof(unsigned long long), cudaMemcpyDeviceToHost);
    cudaFree(buf);
    return rc;
}
#endif

const char* s1s2_view_name(const s1s2_handle* h, int i) {
    if (h == nullptr || i < 0 || i >= static_cast<int>(h->views.size())) return nullptr;
    return h->views[i].name;
}

}  // extern "C"
