// Implicit-GEMM convolution for the UNetSmall denoiser on sm_100a: TMA-fed, tcgen05.mma.cta_group::2 with the
// accumulator in TMEM, warp-specialised and persistent (one CTA per SM, CTAs paired into clusters of 2).
//
//   D[pixel, cout] = sum_{tap, cin} X[pixel shifted by tap, cin] * Wt[cout, tap, cin]
//
// Reference semantics being reproduced (Evaluation/DDIM_Multi-step.py:19-53): nn.Conv2d(k=3, padding=1) + ReLU,
// nn.MaxPool2d(2) after the second conv of each encoder block, nn.ConvTranspose2d(k=2, stride=2) (no activation),
// torch.cat skips (realised by writing producers into halves of one NHWC buffer), and the 1x1 `outc` head, whose
// epilogue also applies the DDIM/DDPM scheduler update of the samplers (DDIM_Multi-step.py:129-133 etc.).
//
// Data layout: activations NHWC fp16; a 128-pixel M tile is a TN x TH x TW block of pixels fetched by ONE 4-D TMA
// box per (tap, channel chunk) at coordinates shifted by (kx-1, ky-1) -- out-of-image rows/columns are zero-filled
// by the TMA unit, which is exactly the conv's zero padding.  Weights are [cout][tap][cin] fp16 (K-major), one 2-D
// TMA box per chunk.  Both operands land in the canonical K-major swizzled layout UMMA reads.
//
// CTA pair: the two CTAs of a cluster work on two consecutive M tiles and the same N tile.  Each CTA loads its own
// 128 activation rows and HALF of the weight rows; one thread of the leader CTA issues M=256 MMAs that read both
// CTAs' shared memory and write both CTAs' TMEM.  Halving the weight traffic per SM is what lifts the kernel off
// the shared-memory bandwidth bound the single-CTA version sat on (profiles/r1a_*: 62-66 % tensor-pipe active).
//
// Warp roles (256 threads per CTA, 384 with a second epilogue warpgroup): warp 0 = TMA producer, warp 1 = MMA issuer
// (leader CTA only), warp 2 = TMEM allocator, warps 4..7 [and 8..11] = epilogue (TMEM lane quadrant = warp_idx % 4).
// Two TMEM accumulators so the epilogue of tile i overlaps the MMAs of tile i+1.
//
// Halo mode (every 3x3 layer): the activations of a (tile, channel chunk) are ONE halo tile that all nine taps read
// through shifted UMMA descriptors; the ring then streams weight tiles only, TPS taps per stage (see the kernel's
// template notes for TPS / WRES / EPIWG / HSLOTS and the measurements behind them).
//
// Dynamic range: activations are fp16, the reference is fp32.  Every model call carries a per-patch power-of-two
// scale s = 2^k derived from max|x_t| (k = 0 while max|x_t| < 16): the first layer stores relu(.)/s, every later
// layer adds bias/s, the head multiplies its output by s.  ReLU, max-pool and the convolutions are positively
// homogeneous, so this is the same network; with k = 0 it is bit-identical to the unscaled arithmetic.
#pragma once
#include <type_traits>

#include "ptx_sm100.cuh"

namespace s1s2 {

enum : int { MODE_STORE = 0, MODE_POOL = 1, MODE_CONVT = 2, MODE_HEAD = 3 };

enum : int { STEP_NONE = 0, STEP_EPS_DDIM = 1, STEP_V_DDIM = 2, STEP_EPS_DDPM = 3, STEP_V_DDPM = 4 };
enum : int { STEP_FLAG_FINAL = 1, STEP_FLAG_NOISE = 2, STEP_FLAG_PHILOX = 4 };
enum : int { LAYER_FLAG_FIRST = 1 };   // first layer: apply the range scale in the epilogue (inputs are unscaled)

constexpr float kXSplit = 4096.f;      // x_t travels as an fp16 pair: x = kXSplit * hi + lo
constexpr uint32_t kAmaxPoison = 0x7FC00000u;   // amax word of a patch whose call is invalid (timestep outside the fp16-exact
                                                // range): above every finite float's bits, so atomicMax keeps it; the head
                                                // then writes NaN for that patch

// One scheduler update; travels by value in the kernel parameters, so a sampling loop is enqueued (or captured in
// a CUDA graph) without device-side bookkeeping or host round trips.
//  EPS_DDIM: x0 = (x - c0*e)/c1            ; xn = c2*x0 + c3*e  (+ c4*z)
//  V_DDIM  : x0 = c0*x - c1*v, e = c1*x + c0*v ; xn = c2*x0 + c3*e  (+ c4*z)
//  EPS_DDPM: xn = c2*(x - c3*e) (+ c4*z)   ; V_DDPM: e = c1*x + c0*v first
//  FINAL   : result = clamp(x0, 0, 1) for DDIM kinds, clamp(xn, 0, 1) for DDPM kinds
struct StepCoef {
    float c0, c1, c2, c3, c4;
    float t_next;   // timestep planted into the next call's time planes
    int kind;
    int flags;
};

constexpr int kHeadIn = 96;   // conv1.2 output channels == outc input channels: base_ch (96, or 64 with the unused tail zero)
constexpr int kHeadOut = 4;

struct HeadParams {
    float w[kHeadOut * kHeadIn];  // outc.weight [4][base_ch], contiguous
    float b[kHeadOut];
    float* x_t;                   // f32 NCHW [B,4,H,W], updated in place (nullptr for a plain forward)
    float* pred_out;              // f32 NCHW eps / v (nullable)
    const float* noise;           // f32 NCHW per-step z (nullable)
    __half* xin16;                // next call's input planes, NHWC16 fp16 (nullable)
    uint32_t* amax_out;           // [B] float bits of max|x_{next}| per patch (nullable)
    uint32_t seed_lo, seed_hi;    // STEP_FLAG_PHILOX: Philox4x32-10 key
    uint32_t noise_stream;        // ... counter word 2 (the step's noise index)
    uint32_t patch_base;          // ... counter word 1 = patch_base + patch slot (global patch id of slot 0)
    StepCoef step;                // by value: kind == STEP_NONE for a plain forward
};

// Division by a launch-constant divisor without the ~40-instruction integer-division sequence: the tile -> (n_tile, tx, ty,
// image) decomposition runs once per tile in every role, and for the short-K layers (first layer, transposed convs) the
// epilogue's instruction count per tile is what bounds the kernel (ncu source view: ~190 of ~520 epilogue instructions
// per tile of inc.0 were these divisions).  mul == 0: power of two (shift only); else q = umulhi(n, mul) >> shr, exact for
// n < 2^31 (mul = ceil(2^(31 + ceil_log2 d) / d), the classic round-up multiplier).
struct FastDiv {
    uint32_t mul, shr, d;
    __host__ static FastDiv make(uint32_t d) {
        FastDiv f;
        f.d = d;
        uint32_t lg = 0;
        while ((1u << lg) < d) ++lg;                       // ceil_log2
        if ((1u << lg) == d) { f.mul = 0; f.shr = lg; return f; }
        const unsigned p = 31 + lg;
        f.mul = static_cast<uint32_t>(((1ull << p) + d - 1) / d);
        f.shr = p - 32;
        return f;
    }
    __device__ __forceinline__ uint32_t div(uint32_t n) const { return mul != 0 ? (__umulhi(n, mul) >> shr) : (n >> shr); }
};

struct ConvParams {
    CUtensorMap tmap_a;
    CUtensorMap tmap_b;
    CUtensorMap tmap_out;         // MODE_STORE / MODE_POOL: NHWC destination, box = 32 channels x the (pooled) pixel tile
    CUtensorMap tmap_out2;        // MODE_CONVT: tmap_out / tmap_out2 = output rows 2y / 2y+1 as 5-D (c, kx, x, y, n) views
    const float* bias;            // [num_n_tiles * BLOCK_N]
    const uint32_t* amax_in;      // [B] float bits of max|x_t| per patch for this call (nullptr: scale 1)
    uint32_t* amax_zero;          // first layer only, nullable: [B] words cleared for this call's head to reduce max|x_next| into
    const uint8_t* next_w;        // weights of the NEXT launch of the chain (nullable) and their size: prefetched into L2
    uint32_t next_w_bytes;        //   while this kernel runs (see prefetch_next_weights)
    __half* out;                  // NHWC fp16 destination (channel offset already applied)
    int out_cpitch;               // elements between consecutive destination pixels
    int H, W, B;                  // input image size, live batch
    int tw_log2, th_log2;         // M tile = TN x TH x TW = 128 pixels
    int tiles_x, tiles_y;
    int num_m_tiles, num_n_tiles;
    FastDiv fd_n, fd_tx, fd_ty;   // divisions by num_n_tiles, tiles_x, tiles_y (conv_umma_kernel's tile decomposition)
    int taps_w;                   // 3 -> 3x3 pad 1 ; 1 -> 1x1 / transposed-conv GEMM
    int chunks;                   // K chunks of KBOX channels per tap
    int tap_kstride;              // conv_px_kernel: K columns between consecutive taps in the weight rows (= Cin)
    int cout;                     // real channels per output pixel (CONVT: per tap)
    int flags;                    // LAYER_FLAG_*
    int perf_mode;                // measurement aid (s1s2_debug_loop_layer): bit 0 / bit 1 = stop re-loading A / B
                                  // once every ring slot has been filled (results are garbage, timing is not)
    HeadParams head;              // MODE_HEAD only
#ifdef S1S2_TIMELINE
    unsigned long long* dbg;      // timeline build only (tools/timeline.py): 16 globaltimer stamps of this launch
#endif
};

// Timeline build (-DS1S2_TIMELINE, never the shipped library): CTA 0 stamps %globaltimer at the hand-over points of its
// roles, every CTA folds its entry / exit time into slots 12 / 13.  Slots: 0 entry, 1 prologue done, 2 past
// griddepcontrol.wait, 3 first loads requested, 4 first activation tile landed, 5 first weight stage landed, 6 last MMA
// issued, 7 first accumulator complete, 8 last tile's stores issued, 9 stores drained, 10 CTA 0 done.
#ifdef S1S2_TIMELINE
#define S1S2_TL(slot)                                                                               \
    do {                                                                                            \
        if (p.dbg != nullptr && blockIdx.x == 0 && (threadIdx.x & 31) == 0) p.dbg[slot] = globaltimer_ns(); \
    } while (0)
#define S1S2_TL_ONCE(flag, slot) do { if (flag) { S1S2_TL(slot); flag = false; } } while (0)
#define S1S2_TL_GRID(slot, op)                                                                      \
    do {                                                                                            \
        if (p.dbg != nullptr && threadIdx.x == 0) op(p.dbg + slot, static_cast<unsigned long long>(globaltimer_ns())); \
    } while (0)
#else
#define S1S2_TL(slot) do { } while (0)
#define S1S2_TL_ONCE(flag, slot) do { } while (0)
#define S1S2_TL_GRID(slot, op) do { } while (0)
#endif

// group tile index -> N tile and the M tile's position (tile column / row inside an image, image index) of CTA `rank`
struct TileCoord { int n_tile, tx, ty, tn; };
template <int CTAS>
__device__ __forceinline__ TileCoord tile_coord(const ConvParams& p, int tile, uint32_t rank) {
    TileCoord c;
    const uint32_t mq = p.fd_n.div(static_cast<uint32_t>(tile));
    c.n_tile = tile - static_cast<int>(mq) * p.num_n_tiles;
    const uint32_t m_tile = CTAS * mq + rank;
    const uint32_t row = p.fd_tx.div(m_tile);
    c.tx = static_cast<int>(m_tile - row * static_cast<uint32_t>(p.tiles_x));
    const uint32_t img = p.fd_ty.div(row);
    c.ty = static_cast<int>(row - img * static_cast<uint32_t>(p.tiles_y));
    c.tn = static_cast<int>(img);
    return c;
}

// Halo mode (3x3 layers with Cin % 64 == 0): the M tile is 8 wide x 16 tall and its activations are fetched ONCE per
// 64-channel chunk as an 18 x 10 pixel halo tile; the nine taps read it through UMMA descriptors whose start address
// is shifted by (ky*10 + kx) 128-byte rows and whose 8-row-group stride is 10 rows (the hardware applies the 128B
// swizzle to absolute shared-memory address bits, tools/umma_halo_test.cu).  Activation traffic drops 6.4x.
constexpr int kHaloW = 10, kHaloH = 18;
constexpr int kHaloSlotsDefault = 3;
template <int KBOX>
struct Halo {                                                   // KBOX channels per pixel row: 128 / 64 / 32-byte rows
    static constexpr int kBytes = kHaloW * kHaloH * KBOX * 2;   // 23040 at KBOX = 64
    static constexpr int kSlot = (kBytes + 1023) / 1024 * 1024; // 1024-aligned ring slot
};

// Counter-based normal noise for the stochastic samplers (DDPM ancestral steps, DDIM eta > 0) when the caller does not
// supply z: Philox4x32-10 keyed by the chain seed, counter = (pixel, global patch id, step noise index, 0) -> four
// uniforms -> two Box-Muller pairs = the four channels of one pixel.  Reproducible whatever the batch composition.
__device__ __forceinline__ void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1,
                                              uint32_t (&out)[4]) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
        const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
        c0 = hi1 ^ c1 ^ k0; c1 = lo1; c2 = hi0 ^ c3 ^ k1; c3 = lo0;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}
__device__ __forceinline__ void philox_normal4(const HeadParams& hp, uint32_t pixel, uint32_t patch_slot, float (&z)[4]) {
    uint32_t u[4];
    philox4x32_10(pixel, hp.patch_base + patch_slot, hp.noise_stream, 0u, hp.seed_lo, hp.seed_hi, u);
#pragma unroll
    for (int i = 0; i < 2; ++i) {
        const float a = (static_cast<float>(u[2 * i]) + 0.5f) * 2.3283064365386963e-10f;        // (0, 1)
        const float b = (static_cast<float>(u[2 * i + 1]) + 0.5f) * 2.3283064365386963e-10f;
        const float r = sqrtf(-2.f * __logf(a));
        float sn, cs;
        __sincosf(6.283185307179586f * b, &sn, &cs);
        z[2 * i] = r * cs;
        z[2 * i + 1] = r * sn;
    }
}

// At small batches the activations of one model call (~123 MB per patch) push the 34.5 MB of weights out of L2, so every
// kernel would stream its weights from HBM: ~1.5 us of latency under a weight ring that covers ~2 us -- measured as 70-80 %
// MMA-phase efficiency at batch 1 against 91 % with L2-resident weights (tools/timeline.py).  Each kernel therefore asks
// L2 for the NEXT kernel's weights (at most 10.6 MB) while it computes: CTA b takes slice b, one idle warp issues it.
__device__ __forceinline__ void prefetch_next_weights(const uint8_t* w, uint32_t bytes, int lane) {
    if (w == nullptr) return;
    const uint32_t per_cta = ((bytes + gridDim.x - 1) / gridDim.x + 4095u) & ~4095u;
    const uint32_t lo = blockIdx.x * per_cta;
    if (lo >= bytes) return;
    const uint32_t n = min(per_cta, bytes - lo);                 // (weight tensors are multiples of 16 bytes)
    const uint32_t per_lane = ((n + 31u) / 32u + 127u) & ~127u;
    const uint32_t off = lane * per_lane;
    if (off < n) bulk_prefetch_l2(w + lo + off, min(per_lane, n - off));
}

template <int BLOCK_N, int KBOX, int BOXES, int STAGES, int MODE, int CTAS, bool HALO = false, int SBUF = 1, int TPS = 1,
          int HSLOTS = kHaloSlotsDefault, int SUBC = 32>
struct ConvSmem {
    static constexpr int kABox = 128 * KBOX * 2;
    static constexpr int kBBox = (BLOCK_N / CTAS) * KBOX * 2;   // this CTA's share of the weight rows
    static constexpr int kARing = HALO ? HSLOTS * Halo<KBOX>::kSlot : 0;
    static constexpr int kStage = HALO ? TPS * kBBox : BOXES * (kABox + kBBox);   // halo mode: the ring holds weight tiles only
                                                                                  // (TPS taps of one channel chunk per stage)
    // output staging for the TMA store: [rows][SUBC channels] fp16 sub-tiles -- SUBC = 32: 64-byte rows, 64B swizzle, one
    // sub-tile per 32-column accumulator chunk; SUBC = 64: 128-byte rows, 128B swizzle, one per two chunks (half as many
    // rows for the TMA unit to write; needs Cout % 64 == 0); rows = 128 pixels, or the 32 pooled pixels of the tile
    static constexpr int kSubRows = MODE == MODE_POOL ? 32 : 128;
    static constexpr int kSubBytes = kSubRows * SUBC * 2;
    static constexpr int kStagingBuf = (MODE != MODE_HEAD) ? (BLOCK_N / SUBC) * kSubBytes : 0;
    static_assert(SUBC == 32 || (SUBC == 64 && BLOCK_N % 64 == 0 && MODE != MODE_HEAD), "staging sub-tile width");
    static constexpr int kStaging = SBUF * kStagingBuf;       // SBUF > 1: short-K layers, where a tile is shorter than a
                                                              // TMA store's smem-read latency
    static constexpr int kBias = 1536 * 4;
    static constexpr int kBytes = 1024 /*align slack*/ + kARing + STAGES * kStage + kStaging + kBias + 256 /*barriers*/;
    static_assert(kStage % 512 == 0 && kBBox % 512 == 0, "stage alignment");
    static_assert(!HALO || (BOXES == 1 && STAGES <= 10), "halo mode");
    static_assert(TPS == 1 || (HALO && (TPS == 3 || TPS == 9)), "taps per stage");
    static_assert((2 * STAGES + 4 + 2 * HSLOTS) * 8 + 4 <= 256, "barrier block");
    static_assert(kBytes <= 232448, "shared memory budget");
};

// {lo: a, hi: b} as fp16 with round-to-nearest and saturation to +-65504 in ONE instruction
__device__ __forceinline__ uint32_t pack_half2_sat(float a, float b) {
    uint32_t r;
    asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(b), "f"(a));
    return r;
}
__device__ __forceinline__ uint32_t hmax2_u32(uint32_t a, uint32_t b) {
    __half2 r = __hmax2(*reinterpret_cast<__half2*>(&a), *reinterpret_cast<__half2*>(&b));
    return *reinterpret_cast<uint32_t*>(&r);
}
// x = kXSplit * hi + lo with hi, lo fp16: ~22 significant bits and a range of 65504 * 4096.
__device__ __forceinline__ void split_x(float x, float& hi, float& lo) {
    hi = __half2float(__float2half_rn(fminf(fmaxf(x * (1.f / kXSplit), -65504.f), 65504.f)));
    lo = x - kXSplit * hi;
}
// Range scale of a patch from the float bits of max|x_t|: exponent k = max(0, floor(log2(amax)) - 3), i.e.
// amax / 2^k < 16.  Returns 2^-k (and 2^k through `s`).
__device__ __forceinline__ float range_scale(uint32_t amax_bits, float& s) {
    int k = static_cast<int>((amax_bits >> 23) & 0xFFu) - 127 - 3;
    k = k < 0 ? 0 : (k > 100 ? 100 : k);
    s = __uint_as_float(static_cast<uint32_t>(127 + k) << 23);
    return __uint_as_float(static_cast<uint32_t>(127 - k) << 23);
}

// CTAS = 2: CTA pair (cta_group::2, M = 256 per MMA, weight rows split across the pair).  CTAS = 1: single-CTA MMAs
// (M = 128), kept for A/B measurements of the pairing itself.
// TPS (halo mode): weight tiles of TPS consecutive taps (one kernel row, or all nine) of a channel chunk share one ring
// stage and one barrier round trip, and the MMA warp issues them from a fully unrolled block with compile-time
// descriptor offsets.  The issue loop costs ~300 cycles per stage whatever it carries (ncu source view, r1s profiles), and an
// MMA N/2 cycles, so stages of few or narrow MMAs (KBOX = 16 / 32: one / two MMAs per tap; N <= 128) are issue-bound at TPS = 1.
// WRES (halo mode, one chunk per tap, TPS = 9, one stage): the nine weight tiles are loaded once per CTA and stay
// resident; the ring then only carries activation halo tiles (inc.0).
// EPIWG: epilogue warpgroups (warps 4..7 [, 8..11]); with two, each drains half of the accumulator's columns (the first
// layer and the transposed convs: their tiles are 9 / 12 ... 48 MMAs long, so the epilogue's instruction count per tile is
// the long pole; 96 columns are split 48 / 48; a third group measured neutral).
// HSLOTS: depth of the activation halo ring (3; deeper for inc.0, whose 6 KB halo tiles are pure TMA latency).
template <int BLOCK_N, int KBOX, int BOXES, int STAGES, int MODE, int CTAS, bool HALO = false, bool WRES = false, int SBUF = 1,
          int TPS = 1, int EPIWG = 1, int HSLOTS = kHaloSlotsDefault, int SUBC = 32>
__global__ void __cluster_dims__(CTAS, 1, 1) __launch_bounds__(128 + 128 * EPIWG, 1)
conv_umma_kernel(const __grid_constant__ ConvParams p) {
    static_assert(!WRES || (HALO && TPS == 9 && STAGES == 1), "resident weights: halo mode, all nine taps in one stage");
    static_assert(EPIWG == 1 || (EPIWG == 2 && MODE != MODE_HEAD), "epilogue warpgroups");
    using L = ConvSmem<BLOCK_N, KBOX, BOXES, STAGES, MODE, CTAS, HALO, SBUF, TPS, HSLOTS, SUBC>;
    constexpr int kHaloSlots = HSLOTS;            // activation halo ring depth (halo mode)
    constexpr bool kPair = CTAS == 2;
    constexpr bool kTmaStore = (MODE != MODE_HEAD);
    static_assert(BLOCK_N % 32 == 0 && BLOCK_N <= 256, "BLOCK_N");
    static_assert(KBOX == 16 || KBOX == 32 || KBOX == 64, "KBOX");
    constexpr int kRowBytes = KBOX * 2;
    constexpr int kAccStride = 256;            // TMEM columns between the two accumulators
    constexpr uint32_t kIdesc = umma_idesc_f16(128 * CTAS, BLOCK_N);

    extern __shared__ uint8_t smem_raw[];
    // 1024-byte alignment by pointer + offset (an integer round trip would lose the shared address space and turn every
    // staging store / load into a generic ST.E / LD.E)
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    uint8_t* a_ring = smem;                       // halo mode: kHaloSlots activation halo tiles
    uint8_t* stage_base = smem + L::kARing;
    uint8_t* sout0 = stage_base + STAGES * L::kStage;
    float* sbias = reinterpret_cast<float*>(sout0 + L::kStaging);
    uint64_t* bars = reinterpret_cast<uint64_t*>(sout0 + L::kStaging + L::kBias);
    uint64_t* full_bar = bars;                    // [STAGES]   (leader's copy is the live one)
    uint64_t* empty_bar = bars + STAGES;          // [STAGES]   (one per CTA, signalled by multicast commits)
    uint64_t* tfull_bar = bars + 2 * STAGES;      // [2]        (one per CTA, multicast commits)
    uint64_t* tempty_bar = bars + 2 * STAGES + 2; // [2]        (leader's copy is the live one)
    uint64_t* afull_bar = bars + 2 * STAGES + 4;  // [kHaloSlots] halo mode (leader's copy is the live one)
    uint64_t* aempty_bar = afull_bar + kHaloSlots;  // [kHaloSlots] halo mode
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(aempty_bar + kHaloSlots);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    if (warp == 0) S1S2_TL(0);
    S1S2_TL_GRID(12, atomicMin);
    [[maybe_unused]] bool tl_a = true, tl_b = true, tl_c = true;
    const uint32_t rank = kPair ? cluster_ctarank() : 0u;
    const int cluster_id = blockIdx.x / CTAS;
    const int num_clusters = gridDim.x / CTAS;
    const int m_pairs = (p.num_m_tiles + CTAS - 1) / CTAS;
    const int num_tiles = m_pairs * p.num_n_tiles;          // group tiles: CTAS M tiles x 1 N tile
    const int k_iters = (p.taps_w * p.taps_w * p.chunks) / BOXES;
    const int pad = p.taps_w >> 1;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&p.tmap_a);
        tma_prefetch_desc(&p.tmap_b);
        if constexpr (kTmaStore) tma_prefetch_desc(&p.tmap_out);
        if constexpr (MODE == MODE_CONVT) tma_prefetch_desc(&p.tmap_out2);
    }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(&full_bar[s], 1);         // leader's producer: arrive.expect_tx for both CTAs' bytes
            mbar_init(&empty_bar[s], 1);        // one multicast commit
        }
        for (int a = 0; a < 2; ++a) {
            mbar_init(&tfull_bar[a], 1);        // one multicast commit
            mbar_init(&tempty_bar[a], 4 * CTAS * EPIWG); // one arrive per epilogue warp of every CTA of the group
        }
        if constexpr (HALO) {
            for (int a = 0; a < kHaloSlots; ++a) {
                mbar_init(&afull_bar[a], 1);
                mbar_init(&aempty_bar[a], 1);
            }
        }
        mbar_fence_init();
    }
    if (warp == 2) { if constexpr (kPair) tmem_alloc_pair<512>(tmem_slot); else tmem_alloc<512>(tmem_slot); }
    {   // bias for every N tile of this layer, staged once
        const int nb = p.num_n_tiles * BLOCK_N;
        for (int i = threadIdx.x; i < nb; i += blockDim.x) sbias[i] = p.bias[i];
    }
    tc_fence_before();
    __syncthreads();
    if constexpr (kPair) cluster_sync_all();    // peer's barriers are initialised before anything targets them
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    if (warp == 0) S1S2_TL(1);
    // Everything above touched only shared memory, TMEM and constant data (tensor maps, bias): under programmatic dependent
    // launch it overlaps the previous kernel's tail.  Activations, amax and the sampler state are read / written below.
    pdl_launch_dependents();
    if (!(HALO && warp == 0)) pdl_wait();       // the halo-mode producer waits after it has requested its first weight tiles
    if (warp == 3) {                                                    // (warp 3 has no other role)
        if (blockIdx.x == 0 && p.amax_zero != nullptr)
            for (int b = lane; b < p.B; b += 32) p.amax_zero[b] = 0u;
        prefetch_next_weights(p.next_w, p.next_w_bytes, lane);
    }

    // Both issue loops below run WARP-UNIFORM (all 32 lanes wait on the barriers and keep the loop state; one
    // elected lane issues TMA / MMA / commit).  Uniform control flow lets ptxas keep descriptors and addresses in
    // uniform registers; a `lane == 0` branch around the loops instead costs an R2UR + ELECT + BRA.U.ANY sequence
    // per instruction and made the single issuing thread the bottleneck (~600 cycles per 4-MMA stage).
    if (HALO && warp == 0) {
        // ================================================================= TMA producer, halo mode (both CTAs)
        // flat walk over (tile, chunk) positions: halo tiles are requested ahead of the weight tiles (see kAhead below)
        int s = 0, sa = 0;
        uint32_t ph = 0, pha = 0;
        auto load_halo = [&](int tile, int chunk) {
            const TileCoord tc = tile_coord<CTAS>(p, tile, rank);
            const int tx = tc.tx, ty = tc.ty, tn = tc.tn;          // tn past the batch for a phantom tile: zero fill
            mbar_wait(&aempty_bar[sa], pha ^ 1);
            const uint32_t bar = kPair ? mapa_shared(smem_u32(&afull_bar[sa]), 0) : smem_u32(&afull_bar[sa]);
            if (elect_one()) {
                if (rank == 0) mbar_expect_tx(&afull_bar[sa], CTAS * Halo<KBOX>::kBytes);
                tma_load_4d_g<kPair>(a_ring + sa * Halo<KBOX>::kSlot, &p.tmap_a, bar, chunk * KBOX, (tx << 3) - 1, (ty << 4) - 1, tn);
            }
            __syncwarp();
            if (++sa == kHaloSlots) { sa = 0; pha ^= 1; }
        };
        int tile = cluster_id, chunk = 0;
        if constexpr (WRES) {                     // one N tile, one chunk per tap: all weights of the layer, once
            const uint32_t bar = kPair ? mapa_shared(smem_u32(&full_bar[0]), 0) : smem_u32(&full_bar[0]);
            if (elect_one()) {
                if (rank == 0) mbar_expect_tx(&full_bar[0], CTAS * 9 * L::kBBox);
#pragma unroll
                for (int tap = 0; tap < 9; ++tap)
                    tma_load_2d_g<kPair>(stage_base + tap * L::kBBox, &p.tmap_b, bar, tap * p.tap_kstride,
                                         static_cast<int>(rank) * (BLOCK_N / CTAS));
            }
            __syncwarp();
        }
        // Weight tiles of the first (tile, chunk) position are requested BEFORE the wait on the previous kernel: weights do
        // not depend on it, and at small batches they come from HBM (the activations of a model call push them out of L2),
        // ~1.3 us that would otherwise sit between the dependency release and the first MMA (tools/timeline.py).  The
        // ring slots are free at kernel start, so no empty-barrier wait is needed here.
        // One ring stage of weight tiles (TPS taps of one channel chunk); `wait` = the slot may still be in use.
        auto weight_stage = [&](int b_row0, int kcol, bool wait) {
            if (wait) mbar_wait(&empty_bar[s], ph ^ 1);
            const uint32_t bar = kPair ? mapa_shared(smem_u32(&full_bar[s]), 0) : smem_u32(&full_bar[s]);
            if (elect_one()) {
                if (rank == 0) mbar_expect_tx(&full_bar[s], CTAS * L::kStage);
#pragma unroll
                for (int j = 0; j < TPS; ++j)
                    tma_load_2d_g<kPair>(stage_base + s * L::kStage + j * L::kBBox, &p.tmap_b, bar, kcol + j * p.tap_kstride, b_row0);
            }
            __syncwarp();
            if (++s == STAGES) { s = 0; ph ^= 1; }
        };
        [[maybe_unused]] int tap0 = 0;
        [[maybe_unused]] const int b_row_first = tile_coord<CTAS>(p, tile, rank).n_tile * BLOCK_N + static_cast<int>(rank) * (BLOCK_N / CTAS);
        if constexpr (!WRES) {
            if (tile < num_tiles)
                for (; tap0 < 9 && tap0 < STAGES * TPS; tap0 += TPS) weight_stage(b_row_first, tap0 * p.tap_kstride, false);
        }
        pdl_wait();
        S1S2_TL(2);
        // the halo cursor runs kAhead = HSLOTS - 2 positions ahead of the weight stream: the slot it targets was freed two
        // positions earlier, so the request never blocks on the MMA warp while the weight ring still has work queued
        constexpr int kAhead = kHaloSlots - 2;
        static_assert(kAhead >= 1, "halo ring depth");
        int htile = cluster_id, hchunk = 0;
        auto advance = [&](int& t, int& c) { if (++c == p.chunks) { c = 0; t += num_clusters; } };
        for (int d = 0; d < kAhead; ++d)
            if (htile < num_tiles) { load_halo(htile, hchunk); advance(htile, hchunk); }
        S1S2_TL(3);
        // Order inside a position.  64-channel chunks: next halo tile first, then this chunk's weights (a chunk is 9 x 4 MMAs,
        // the halo request is far ahead either way).  32-channel chunks (down1.0.0): WEIGHTS FIRST -- a chunk is only 1728 MMA
        // cycles, the weight ring holds two chunks, and a halo request that has to wait for its slot (freed two chunks back)
        // would hold this chunk's weights back with it: ncu's source view showed the MMA warp spending 42 % of its time
        // waiting for the first weight stage of each chunk, and none on the second and third.
        constexpr bool kWeightsFirst = KBOX == 32;
        if constexpr (!WRES) {                    // the first position, peeled: its first weight stages are already in flight
            if (tile < num_tiles) {
                if (!kWeightsFirst && htile < num_tiles) { load_halo(htile, hchunk); advance(htile, hchunk); }
                for (; tap0 < 9; tap0 += TPS) weight_stage(b_row_first, tap0 * p.tap_kstride, true);
                if (kWeightsFirst && htile < num_tiles) { load_halo(htile, hchunk); advance(htile, hchunk); }
                advance(tile, chunk);
            }
        }
        while (tile < num_tiles) {
            if (!kWeightsFirst && htile < num_tiles) { load_halo(htile, hchunk); advance(htile, hchunk); }
            const int b_row0 = tile_coord<CTAS>(p, tile, rank).n_tile * BLOCK_N + static_cast<int>(rank) * (BLOCK_N / CTAS);
            int kcol = chunk * KBOX;
            for (int tap = 0; tap < (WRES ? 0 : 9); tap += TPS, kcol += TPS * p.tap_kstride) {
                mbar_wait(&empty_bar[s], ph ^ 1);
                const uint32_t bar = kPair ? mapa_shared(smem_u32(&full_bar[s]), 0) : smem_u32(&full_bar[s]);
                if (elect_one()) {
                    if (rank == 0) mbar_expect_tx(&full_bar[s], CTAS * L::kStage);
#pragma unroll
                    for (int j = 0; j < TPS; ++j)
                        tma_load_2d_g<kPair>(stage_base + s * L::kStage + j * L::kBBox, &p.tmap_b, bar,
                                             kcol + j * p.tap_kstride, b_row0);
                }
                __syncwarp();
                if (++s == STAGES) { s = 0; ph ^= 1; }
            }
            if (kWeightsFirst && htile < num_tiles) { load_halo(htile, hchunk); advance(htile, hchunk); }
            advance(tile, chunk);
        }
    } else if (HALO && warp == 1) {
        // ================================================================= MMA issuer, halo mode (leader CTA)
        if (rank == 0) {
            int s = 0, sa = 0;
            uint32_t ph = 0, pha = 0;
            int acc = 0;
            uint32_t acc_ph = 0;
            [[maybe_unused]] bool wres_ready = false;
            const uint32_t b0 = smem_u32(stage_base), a0 = smem_u32(a_ring);
            for (int tile = cluster_id; tile < num_tiles; tile += num_clusters) {
                mbar_wait(&tempty_bar[acc], acc_ph ^ 1);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + acc * kAccStride;
                for (int chunk = 0; chunk < p.chunks; ++chunk) {
                    mbar_wait(&afull_bar[sa], pha);
                    S1S2_TL_ONCE(tl_a, 4);
                    const uint32_t a_addr = a0 + sa * Halo<KBOX>::kSlot;
                    // halo descriptor: rows kRowBytes apart, 8-row groups (= image rows of the 8-wide tile) 10 rows apart
                    uint64_t adesc0 = static_cast<uint64_t>(1) << 16;
                    adesc0 |= static_cast<uint64_t>((kHaloW * kRowBytes) >> 4) << 32;
                    adesc0 |= static_cast<uint64_t>(1) << 46;
                    adesc0 |= (kRowBytes == 128 ? 2ull : (kRowBytes == 64 ? 4ull : 6ull)) << 61;
                    if constexpr (TPS == 1) {
                        for (int tap = 0; tap < 9; ++tap) {
                            mbar_wait(&full_bar[s], ph);
                            S1S2_TL_ONCE(tl_b, 5);
                            tc_fence_after();
                            const int ky = tap / 3, kx = tap - 3 * ky;
                            const uint32_t a_tap = a_addr + (ky * kHaloW + kx) * kRowBytes;
                            if (elect_one()) {
                                const uint64_t adesc = adesc0 | static_cast<uint64_t>((a_tap & 0x3FFFFu) >> 4);
                                const uint64_t bdesc = umma_smem_desc<kRowBytes>(b0 + s * L::kStage);
#pragma unroll
                                for (int k = 0; k < KBOX / 16; ++k)
                                    umma_f16_g<kPair>(d_tmem, adesc + 2 * k, bdesc + 2 * k, kIdesc, (chunk | tap | k) != 0 ? 1u : 0u);
                                umma_commit_g<kPair>(&empty_bar[s]);
                                if (tap == 8) {
                                    umma_commit_g<kPair>(&aempty_bar[sa]);
                                    if (chunk == p.chunks - 1) umma_commit_g<kPair>(&tfull_bar[acc]);
                                }
                            }
                            __syncwarp();
                            if (++s == STAGES) { s = 0; ph ^= 1; }
                        }
                    } else {
                        // TPS taps per stage: one barrier round trip, then a straight-line block of TPS * KBOX/16 MMAs whose
                        // descriptors differ from the stage's base descriptors by compile-time constants
                        const uint64_t adesc_c = adesc0 | static_cast<uint64_t>((a_addr & 0x3FFFFu) >> 4);
#pragma unroll
                        for (int t0 = 0; t0 < 9; t0 += TPS) {
                            if constexpr (WRES) {
                                if (!wres_ready) { mbar_wait(&full_bar[0], 0); wres_ready = true; }
                            } else {
                                mbar_wait(&full_bar[s], ph);
                            }
                            S1S2_TL_ONCE(tl_b, 5);
                            tc_fence_after();
                            if (elect_one()) {
                                const uint64_t bdesc_s = umma_smem_desc<kRowBytes>(b0 + (WRES ? 0 : s) * L::kStage);
#pragma unroll
                                for (int j = 0; j < TPS; ++j) {
                                    const int tap = t0 + j, ky = tap / 3, kx = tap - 3 * ky;
                                    const uint64_t adesc = adesc_c + static_cast<uint64_t>(((ky * kHaloW + kx) * kRowBytes) >> 4);
                                    const uint64_t bdesc = bdesc_s + static_cast<uint64_t>((j * L::kBBox) >> 4);
#pragma unroll
                                    for (int k = 0; k < KBOX / 16; ++k)
                                        umma_f16_g<kPair>(d_tmem, adesc + 2 * k, bdesc + 2 * k, kIdesc,
                                                          (tap | k) != 0 ? 1u : (chunk != 0 ? 1u : 0u));
                                }
                                if constexpr (!WRES) umma_commit_g<kPair>(&empty_bar[s]);
                                if (t0 + TPS == 9) {
                                    umma_commit_g<kPair>(&aempty_bar[sa]);
                                    if (chunk == p.chunks - 1) umma_commit_g<kPair>(&tfull_bar[acc]);
                                }
                            }
                            __syncwarp();
                            if constexpr (!WRES) { if (++s == STAGES) { s = 0; ph ^= 1; } }
                        }
                    }
                    if (++sa == kHaloSlots) { sa = 0; pha ^= 1; }
                }
                acc ^= 1;
                if (acc == 0) acc_ph ^= 1;
            }
            S1S2_TL(6);
        }
    } else if (warp == 0) {
        // ================================================================= TMA producer (both CTAs)
        int s = 0;
        uint32_t ph = 0;
        int issued = 0;
        S1S2_TL(2);
        S1S2_TL(3);
        for (int tile = cluster_id; tile < num_tiles; tile += num_clusters) {
            const TileCoord tc = tile_coord<CTAS>(p, tile, rank);
            const int n_tile = tc.n_tile, tx = tc.tx, ty = tc.ty, tn = tc.tn;
            const int x0 = (tx << p.tw_log2) - pad;
            const int y0 = (ty << p.th_log2) - pad;
            const int n0 = tn << (7 - p.tw_log2 - p.th_log2);    // past the batch for a phantom tile: zero fill
            const int b_row0 = n_tile * BLOCK_N + static_cast<int>(rank) * (BLOCK_N / CTAS);
            int chunk = 0, kx = 0, ky = 0, kcol = 0;             // running (tap, chunk) position, no divisions
            for (int it = 0; it < k_iters; ++it, ++issued) {
                mbar_wait(&empty_bar[s], ph ^ 1);
                const bool load_a = !((p.perf_mode & 1) && issued >= STAGES);
                const bool load_b = !((p.perf_mode & 2) && issued >= STAGES);
                uint8_t* a_dst = stage_base + s * L::kStage;
                uint8_t* b_dst = a_dst + BOXES * L::kABox;
                const uint32_t full_leader = kPair ? mapa_shared(smem_u32(&full_bar[s]), 0) : smem_u32(&full_bar[s]);
                if (elect_one()) {
                    if (rank == 0)
                        mbar_expect_tx(&full_bar[s], CTAS * BOXES * ((load_a ? L::kABox : 0) + (load_b ? L::kBBox : 0)));
#pragma unroll
                    for (int b = 0; b < BOXES; ++b) {
                        int c_ = chunk + b, kx_ = kx, ky_ = ky;                  // BOXES <= chunks or chunks == 1
                        if (p.chunks == 1) { kx_ += b; if (kx_ >= p.taps_w) { kx_ -= p.taps_w; ++ky_; } c_ = 0; }
                        if (load_a)
                            tma_load_4d_g<kPair>(a_dst + b * L::kABox, &p.tmap_a, full_leader, c_ * KBOX, x0 + kx_, y0 + ky_, n0);
                        if (load_b)
                            tma_load_2d_g<kPair>(b_dst + b * L::kBBox, &p.tmap_b, full_leader, kcol + b * KBOX, b_row0);
                    }
                }
                __syncwarp();
                kcol += BOXES * KBOX;
                if (p.chunks == 1) {
                    kx += BOXES;
                    while (kx >= p.taps_w) { kx -= p.taps_w; ++ky; }
                } else {
                    chunk += BOXES;
                    if (chunk >= p.chunks) { chunk = 0; if (++kx == p.taps_w) { kx = 0; ++ky; } }
                }
                if (++s == STAGES) { s = 0; ph ^= 1; }
            }
        }
    } else if (warp == 1) {
        // ================================================================= MMA issuer (leader CTA)
        if (rank == 0) {
            int s = 0;
            uint32_t ph = 0;
            int acc = 0;
            uint32_t acc_ph = 0;
            const uint32_t stage0 = smem_u32(stage_base);
            for (int tile = cluster_id; tile < num_tiles; tile += num_clusters) {
                mbar_wait(&tempty_bar[acc], acc_ph ^ 1);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + acc * kAccStride;
                for (int it = 0; it < k_iters; ++it) {
                    mbar_wait(&full_bar[s], ph);
                    S1S2_TL_ONCE(tl_b, 5);
                    tc_fence_after();
                    const uint32_t a_addr = stage0 + s * L::kStage;
                    const uint32_t b_addr = a_addr + BOXES * L::kABox;
                    if (elect_one()) {
#pragma unroll
                        for (int b = 0; b < BOXES; ++b) {
                            const uint64_t adesc = umma_smem_desc<kRowBytes>(a_addr + b * L::kABox);
                            const uint64_t bdesc = umma_smem_desc<kRowBytes>(b_addr + b * L::kBBox);
#pragma unroll
                            for (int k = 0; k < KBOX / 16; ++k) {
                                // advance 16 K-elements = 32 bytes inside the swizzle span: +2 in the (addr >> 4) field
                                umma_f16_g<kPair>(d_tmem, adesc + 2 * k, bdesc + 2 * k, kIdesc, (it | b | k) != 0 ? 1u : 0u);
                            }
                        }
                        umma_commit_g<kPair>(&empty_bar[s]);    // (both CTAs') slots reusable once these MMAs retire
                        if (it == k_iters - 1) umma_commit_g<kPair>(&tfull_bar[acc]);
                    }
                    __syncwarp();
                    if (++s == STAGES) { s = 0; ph ^= 1; }
                }
                acc ^= 1;
                if (acc == 0) acc_ph ^= 1;
            }
            S1S2_TL(6);
        }
    } else if (warp >= 4) {
        // ================================================================= epilogue (1 pixel per thread; with two
        // warpgroups, group g drains the column chunks [g * kChunksWg, ...) of the same accumulator)
        constexpr int kChunks = BLOCK_N / 32;
        constexpr int kChunksWg = (kChunks + EPIWG - 1) / EPIWG;
        const int c_lo = EPIWG == 1 ? 0 : ((warp - 4) >> 2) * kChunksWg;
        const int q = warp & 3;
        const int m = q * 32 + lane;
        const int tw = 1 << p.tw_log2;
        const int lx = m & (tw - 1);
        const int ly = (m >> p.tw_log2) & ((1 << p.th_log2) - 1);
        const int ln = m >> (p.tw_log2 + p.th_log2);
        int acc = 0;
        uint32_t acc_ph = 0;
        int sbuf = 0;
        for (int tile = cluster_id; tile < num_tiles; tile += num_clusters) {
            uint8_t* sout = sout0 + sbuf * L::kStagingBuf;
            if (++sbuf == SBUF) sbuf = 0;
            const TileCoord tc = tile_coord<CTAS>(p, tile, rank);
            const int n_tile = tc.n_tile, tx = tc.tx, ty = tc.ty, tn = tc.tn;
            const int x = (tx << p.tw_log2) + lx;
            const int y = (ty << p.th_log2) + ly;
            const int n = (tn << (7 - p.tw_log2 - p.th_log2)) + ln;
            const bool live = (n < p.B) && (y < p.H) && (x < p.W);
            float s_up = 1.f, s_dn = 1.f;
            [[maybe_unused]] bool poisoned = false;
            if (live && p.amax_in != nullptr) {
                const uint32_t bits = __ldg(p.amax_in + n);
                s_dn = range_scale(bits, s_up);
                poisoned = bits > 0x7F800000u;
            }

            mbar_wait(&tfull_bar[acc], acc_ph);
            if (warp == 4) S1S2_TL_ONCE(tl_c, 7);
            tc_fence_after();
            const uint32_t taddr = tmem_base + acc * kAccStride + (static_cast<uint32_t>(q * 32) << 16);
            const float* bias_t = sbias + n_tile * BLOCK_N;
            const uint32_t tempty_leader = kPair ? mapa_shared(smem_u32(&tempty_bar[acc]), 0) : smem_u32(&tempty_bar[acc]);

            if constexpr (MODE == MODE_HEAD) {
                static_assert(MODE != MODE_HEAD || BLOCK_N == kHeadIn, "head wants all channels of a pixel");
                float o[kHeadOut];
#pragma unroll
                for (int k = 0; k < kHeadOut; ++k) o[k] = p.head.b[k] * s_dn;
#pragma unroll
                for (int c = 0; c < BLOCK_N / 32; ++c) {
                    uint32_t r[32];
                    tmem_ld32(taddr + c * 32, r);
                    tmem_ld_wait();
#pragma unroll
                    for (int j = 0; j < 32; ++j) {
                        const float h = fmaxf(fmaf(bias_t[c * 32 + j], s_dn, __uint_as_float(r[j])), 0.f);
#pragma unroll
                        for (int k = 0; k < kHeadOut; ++k) o[k] = fmaf(h, p.head.w[k * kHeadIn + c * 32 + j], o[k]);
                    }
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive_cluster(tempty_leader);  // accumulator drained: MMAs of tile+2 may start

                const StepCoef& sc = p.head.step;
                float amax = 0.f;
                if (live) {
                    const size_t plane = static_cast<size_t>(p.H) * p.W;
                    const size_t pix = static_cast<size_t>(y) * p.W + x;
                    float res[kHeadOut];
#pragma unroll
                    for (int k = 0; k < kHeadOut; ++k) {
                        const size_t idx = (static_cast<size_t>(n) * kHeadOut + k) * plane + pix;
                        const float pr = poisoned ? __uint_as_float(kAmaxPoison) : o[k] * s_up;
                        if (p.head.pred_out != nullptr) p.head.pred_out[idx] = pr;
                        res[k] = pr;
                        if (sc.kind != STEP_NONE) {
                            const float xt = p.head.x_t[idx];
                            float x0 = 0.f, e = pr, xn;
                            if (sc.kind == STEP_EPS_DDIM) {
                                x0 = __fdiv_rn(__fsub_rn(xt, __fmul_rn(sc.c0, pr)), sc.c1);
                            } else if (sc.kind == STEP_V_DDIM || sc.kind == STEP_V_DDPM) {
                                x0 = __fsub_rn(__fmul_rn(sc.c0, xt), __fmul_rn(sc.c1, pr));
                                e = __fadd_rn(__fmul_rn(sc.c1, xt), __fmul_rn(sc.c0, pr));
                            }
                            const bool ddpm = sc.kind == STEP_EPS_DDPM || sc.kind == STEP_V_DDPM;
                            if (ddpm) xn = __fmul_rn(sc.c2, __fsub_rn(xt, __fmul_rn(sc.c3, e)));
                            else      xn = __fadd_rn(__fmul_rn(sc.c2, x0), __fmul_rn(sc.c3, e));
                            if (sc.flags & STEP_FLAG_NOISE) xn = __fadd_rn(xn, __fmul_rn(sc.c4, p.head.noise[idx]));
                            if (sc.flags & STEP_FLAG_FINAL) xn = fminf(fmaxf(ddpm ? xn : x0, 0.f), 1.f);
                            p.head.x_t[idx] = xn;
                            res[k] = xn;
                            amax = fmaxf(amax, fabsf(xn));
                        }
                    }
                    if (sc.kind != STEP_NONE && p.head.xin16 != nullptr) {
                        float hi[kHeadOut], lo[kHeadOut];
#pragma unroll
                        for (int k = 0; k < kHeadOut; ++k) split_x(res[k], hi[k], lo[k]);
                        __half* rec = p.head.xin16 + ((static_cast<size_t>(n) * p.H + y) * p.W + x) * 16;
                        uint4 v;
                        v.x = pack_half2_sat(lo[0], lo[1]);
                        v.y = pack_half2_sat(lo[2], lo[3]);
                        v.z = pack_half2_sat(sc.t_next, sc.t_next);
                        v.w = 0u;
                        *reinterpret_cast<uint4*>(rec) = v;                               // slots 0..7
                        uint2 u;
                        u.x = pack_half2_sat(hi[0], hi[1]);
                        u.y = pack_half2_sat(hi[2], hi[3]);
                        *reinterpret_cast<uint2*>(rec + 12) = u;                          // slots 12..15
                    }
                }
                if (sc.kind != STEP_NONE && p.head.amax_out != nullptr) {
                    // per-patch max|x_next| for the next call's range scale (non-negative floats order like uints)
                    const uint32_t key = live ? static_cast<uint32_t>(n) : 0xFFFFFFFFu;
                    const uint32_t key0 = __shfl_sync(0xffffffffu, key, 0);
                    const uint32_t bits = __float_as_uint(amax);
                    if (__all_sync(0xffffffffu, key == key0)) {
                        const uint32_t mx = __reduce_max_sync(0xffffffffu, bits);
                        if (lane == 0 && key0 != 0xFFFFFFFFu) atomicMax(p.head.amax_out + key0, mx);
                    } else if (live) {
                        atomicMax(p.head.amax_out + n, bits);
                    }
                }
            } else {
                // The tile is staged in shared memory ([row][32 ch] sub-tiles, 64B swizzle) and leaves through TMA
                // stores (full-sector, asynchronous, clipped at the image / batch border).  CONVT: the 32 columns of a
                // sub-tile belong to one tap (ky, kx); its store walks the output with element strides (1, 2, 2, 1)
                // from (co, 2*x0 + kx, 2*y0 + ky), i.e. the pixel shuffle is done by the TMA unit.
                int srow = m;                                  // staging row of this thread's pixel
                bool writer = true;
                if constexpr (MODE == MODE_POOL) {
                    writer = ((lx | ly) & 1) == 0;
                    srow = ((ln << (p.th_log2 - 1)) + (ly >> 1)) * (tw >> 1) + (lx >> 1);
                }
                if constexpr (kTmaStore) {
                    if (warp == 4 && lane == 0) bulk_wait_read<SBUF - 1>();   // this buffer's previous stores have read it
                    named_bar_sync(1, 128 * EPIWG);
                }
                const bool first = (p.flags & LAYER_FLAG_FIRST) != 0;
                const float s_bias = first ? 1.f : s_dn;      // first layer: unscaled inputs, scale the result
                const float s_post = first ? s_dn : 1.f;
                // One segment of WD accumulator columns starting at col0: TMEM -> bias / ReLU / scale -> fp16 (-> 2x2 max) -> staging.
                auto segment = [&](auto wd_tag, const int col0) {
                    constexpr int WD = decltype(wd_tag)::value;
                    uint32_t r[WD];
                    if constexpr (WD == 32) tmem_ld32(taddr + col0, r); else tmem_ld16(taddr + col0, r);
                    tmem_ld_wait();
                    uint32_t h[WD / 2];
#pragma unroll
                    for (int j = 0; j < WD / 2; ++j) {
                        float a = fmaf(bias_t[col0 + 2 * j], s_bias, __uint_as_float(r[2 * j]));
                        float b = fmaf(bias_t[col0 + 2 * j + 1], s_bias, __uint_as_float(r[2 * j + 1]));
                        if constexpr (MODE != MODE_CONVT) { a = fmaxf(a, 0.f); b = fmaxf(b, 0.f); }
                        if constexpr (MODE == MODE_STORE) { a *= s_post; b *= s_post; }
                        h[j] = pack_half2_sat(a, b);
                    }
                    if constexpr (MODE == MODE_POOL) {
#pragma unroll
                        for (int j = 0; j < WD / 2; ++j) {
                            h[j] = hmax2_u32(h[j], __shfl_xor_sync(0xffffffffu, h[j], 1));
                            h[j] = hmax2_u32(h[j], __shfl_xor_sync(0xffffffffu, h[j], tw));
                        }
                    }
                    if constexpr (kTmaStore) {
                        if (writer) {
                            // staging row of this pixel inside the SUBC-channel sub-tile the segment belongs to; 16-byte
                            // chunk j0 + i of the row, XOR-swizzled like the store tensor map (64B: row[2:1], 128B: row[2:0])
                            uint8_t* row = sout + (col0 / SUBC) * L::kSubBytes + srow * (SUBC * 2);
                            const int sw = SUBC == 32 ? ((srow >> 1) & 3) : (srow & 7);
                            const int j0 = (col0 % SUBC) >> 3;
#pragma unroll
                            for (int i = 0; i < WD / 8; ++i)
                                *reinterpret_cast<uint4*>(row + (((j0 + i) ^ sw) << 4)) =
                                    make_uint4(h[4 * i], h[4 * i + 1], h[4 * i + 2], h[4 * i + 3]);
                        }
                    }
                };
                using W32 = std::integral_constant<int, 32>;
                using W16 = std::integral_constant<int, 16>;
                if constexpr (BLOCK_N == 96 && EPIWG == 2) {
                    // 96 columns over two warpgroups: 48 each (32 + 16 / 16 + 32) instead of two chunks against one
                    if (c_lo == 0) { segment(W32{}, 0); segment(W16{}, 32); }
                    else { segment(W16{}, 48); segment(W32{}, 64); }
                } else {
#pragma unroll
                    for (int cc = 0; cc < kChunksWg; ++cc) {
                        const int c = c_lo + cc;
                        if (EPIWG > 1 && c >= kChunks) break;
                        segment(W32{}, c * 32);
                    }
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive_cluster(tempty_leader);
                if constexpr (kTmaStore) {
                    fence_proxy_async_smem();
                    named_bar_sync(1, 128 * EPIWG);
                    if (warp == 4 && lane == 0) {
                        const int on0 = tn << (7 - p.tw_log2 - p.th_log2);
                        if constexpr (MODE == MODE_CONVT) {
                            // the 32 columns of a sub-tile belong to one tap (ky, kx): output pixels (2x + kx, 2y + ky) of the
                            // tile = a plain box of the 5-D view (c, kx, x, y, n) of the output rows of parity ky
                            const int ix0 = tx << p.tw_log2, iy0 = ty << p.th_log2;
#pragma unroll
                            for (int c = 0; c < BLOCK_N / SUBC; ++c) {
                                const int ng = n_tile * BLOCK_N + c * SUBC;
                                const int tap = ng / p.cout;
                                tma_store_5d((tap >> 1) ? &p.tmap_out2 : &p.tmap_out, sout + c * L::kSubBytes, ng - tap * p.cout,
                                             tap & 1, ix0, iy0, on0);
                            }
                        } else {
                            const int sh = MODE == MODE_POOL ? 1 : 0;
                            const int ox0 = (tx << p.tw_log2) >> sh, oy0 = (ty << p.th_log2) >> sh;
#pragma unroll
                            for (int c = 0; c < BLOCK_N / SUBC; ++c)
                                tma_store_4d(&p.tmap_out, sout + c * L::kSubBytes, n_tile * BLOCK_N + c * SUBC, ox0, oy0, on0);
                        }
                        bulk_commit();
                    }
                }
            }
            acc ^= 1;
            if (acc == 0) acc_ph ^= 1;
        }
        if (warp == 4) S1S2_TL(8);
    }

    if constexpr (kTmaStore) {
        // the staging buffers may be released once the outstanding stores have READ them; completion of the writes is
        // ordered by grid completion (what the next kernel's griddepcontrol.wait / stream order waits for)
        if (warp == 4 && lane == 0) bulk_wait_read0();
    }
    if (warp == 4) S1S2_TL(9);
    tc_fence_before();
    __syncthreads();
    if constexpr (kPair) cluster_sync_exec();   // both CTAs are done with each other's shared memory and TMEM
    if (warp == 2) {
        tc_fence_after();
        if constexpr (kPair) tmem_dealloc_pair<512>(tmem_base); else tmem_dealloc<512>(tmem_base);
    }
    if (warp == 0) S1S2_TL(10);
    S1S2_TL_GRID(13, atomicMax);
}

}  // namespace s1s2
