// Pixels-on-N convolution kernel for the head layer: conv1.2 (+ outc + scheduler), Cout = 96 (or 64).
//
// (Round 1 built it for both Cout = 96 layers on the reading that one tcgen05.mma costs ~128 cycles for any N; that
// experiment was issue-bound -- an MMA takes N/2 cycles -- and conv1.0 now runs 8 % faster on conv_umma_kernel with
// 96-column tiles and three taps per stage.  For the head the pixels-on-N form still wins: its two epilogue warpgroups
// work on different tiles, while the Cout-on-N head has one group doing the 96 x 4 head FMAs of every pixel)
// Here the GEMM is transposed:
//
//   D^T[cout, pixel] = sum_{tap, cin} Wt[cout, tap, cin] * X[pixel shifted by tap, cin]
//
// M = 128 weight rows (96 real; the TMA box overhangs the 96-row weight tensor and the overhang is zero-filled), N = 256
// pixels = one 8 wide x 32 tall spatial tile of one image, K chunks of KBOX channels per tap.  Halo mode like
// conv_umma_kernel: the pixel operand of a channel chunk is ONE 34 x 10 halo tile that all nine taps read through
// shifted descriptors (8-row groups = image rows, 10 rows apart); only the weight tiles stream per tap.  Same swizzles
// and UMMA descriptors as conv_umma_kernel with the operand roles swapped; single CTA (cta_group::1).
//
// Epilogue: a thread owns one output channel (TMEM lane) and sees the tile's 256 pixels as columns.  bias + ReLU in
// fp32, then the tile is transposed through shared memory, half a tile (8 image rows) at a time to keep the staging at
// 24 KB and the operand ring deep, into the [pixel][32 ch] 64B-swizzled sub-tiles that
//   * MODE_STORE: TMA stores write to the NHWC destination (conv1.0);
//   * MODE_HEAD:  a second pass reads back pixel-major - one thread per pixel - to apply the 1x1 outc (96 -> 4, fp32
//                 accumulation over the fp16 hidden row), the scheduler update, and the writes of x_{t-1}, the next
//                 input record, the eps/v tap and max|x_{t-1}| (same arithmetic as conv_umma_kernel's MODE_HEAD).
#pragma once
#include "conv_umma.cuh"

namespace s1s2 {

constexpr int kPxHaloW = 10, kPxHaloH = 34;

// TPS: weight tiles (taps) per ring stage, HSLOTS: halo ring depth (see conv_umma_kernel's TPS note: the issue loop is
// ~300 cycles per stage, so KBOX = 32 stages of one tap = two MMAs were issue-bound).
// NQ: 32-channel quarters of the output (3 for Cout = 96, 2 for Cout = 64): sub-tiles of the staging buffer, TMEM lane
// quadrants drained in pass 1, and the length NQ * 32 of the hidden row the 1x1 head contracts.
template <int KBOX, int STAGES, int TPS = 1, int HSLOTS = 2, int NQ = 3>
struct PxSmem {
    static constexpr int kABox = 128 * KBOX * 2;        // weights: 128 rows (cout, zero-padded past the real rows)
    static constexpr int kHaloBytes = kPxHaloW * kPxHaloH * KBOX * 2;          // pixels: 8 x 32 tile + 1-pixel ring
    static constexpr int kHaloSlot = (kHaloBytes + 1023) / 1024 * 1024;
    static constexpr int kHaloRing = HSLOTS * kHaloSlot;
    static constexpr int kStage = TPS * kABox;          // the ring streams weight tiles only
    static constexpr int kSubBytes = 128 * 64;          // [128 pixels = half a tile][32 ch] fp16, 64B swizzle
    static constexpr int kStagingBuf = NQ * kSubBytes;
    static constexpr int kStaging = 2 * kStagingBuf;    // one buffer per epilogue warpgroup
    static constexpr int kBias = 128 * 4;
    static constexpr int kBytes = 1024 + kHaloRing + STAGES * kStage + kStaging + kBias + 256;
    static_assert(STAGES >= 2 && STAGES <= 10, "weight ring depth");
    static_assert(TPS == 1 || TPS == 3, "taps per stage");
    static_assert(HSLOTS >= 2 && HSLOTS <= 4, "halo ring depth");
    static_assert(kBytes <= 232448, "shared memory budget");
};

// 384 threads: warp 0 TMA producer, warp 1 MMA issuer, warp 2 TMEM allocator, warps 4-7 and 8-11 two epilogue
// warpgroups.  The epilogue is a long pole of these layers (transposition through shared memory, and for the head a
// second pass with global loads), so the two TMEM accumulators are drained by different warpgroups: group g owns
// accumulator g, staging buffer g and named barrier 1+g and handles every other tile of the CTA.
constexpr int kPxThreads = 384;

template <int KBOX, int STAGES, int MODE, int TPS = 1, int HSLOTS = 2, int NQ = 3>
__global__ void __launch_bounds__(kPxThreads, 1) conv_px_kernel(const __grid_constant__ ConvParams p) {
    using L = PxSmem<KBOX, STAGES, TPS, HSLOTS, NQ>;
    static_assert(NQ == 2 || NQ == 3, "Cout = 64 or 96");
    constexpr int kHidden = NQ * 32;            // hidden channels = Cout of this layer = input channels of the 1x1 head
    constexpr int kPxHaloSlots = HSLOTS;
    static_assert(MODE == MODE_STORE || MODE == MODE_HEAD, "conv_px_kernel modes");
    constexpr int kRowBytes = KBOX * 2;
    constexpr int kAccStride = 256;
    constexpr uint32_t kIdesc = umma_idesc_f16(128, 256);

    extern __shared__ uint8_t smem_raw[];
    // 1024-byte alignment by pointer + offset (an integer round trip would lose the shared address space and turn every
    // staging store / load into a generic ST.E / LD.E)
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    uint8_t* halo_ring = smem;
    uint8_t* stage_base = smem + L::kHaloRing;
    uint8_t* sout0 = stage_base + STAGES * L::kStage;
    float* sbias = reinterpret_cast<float*>(sout0 + L::kStaging);
    uint64_t* bars = reinterpret_cast<uint64_t*>(sout0 + L::kStaging + L::kBias);
    uint64_t* full_bar = bars;
    uint64_t* empty_bar = bars + STAGES;
    uint64_t* tfull_bar = bars + 2 * STAGES;
    uint64_t* tempty_bar = bars + 2 * STAGES + 2;
    uint64_t* hfull_bar = bars + 2 * STAGES + 4;      // [kPxHaloSlots]
    uint64_t* hempty_bar = hfull_bar + kPxHaloSlots;  // [kPxHaloSlots]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(hempty_bar + kPxHaloSlots);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    if (warp == 0) S1S2_TL(0);
    S1S2_TL_GRID(12, atomicMin);
    [[maybe_unused]] bool tl_a = true, tl_b = true, tl_c = true;
    const int tiles_x = p.W >> 3, tiles_y = (p.H + 31) >> 5;
    const int num_tiles = tiles_x * tiles_y * p.B;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&p.tmap_a);
        tma_prefetch_desc(&p.tmap_b);
        if constexpr (MODE == MODE_STORE) tma_prefetch_desc(&p.tmap_out);
    }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(&full_bar[s], 1);
            mbar_init(&empty_bar[s], 1);
        }
        for (int a = 0; a < 2; ++a) {
            mbar_init(&tfull_bar[a], 1);
            mbar_init(&tempty_bar[a], 4);
        }
        for (int a = 0; a < kPxHaloSlots; ++a) {
            mbar_init(&hfull_bar[a], 1);
            mbar_init(&hempty_bar[a], 1);
        }
        mbar_fence_init();
    }
    if (warp == 2) tmem_alloc<512>(tmem_slot);
    if (threadIdx.x < 128) sbias[threadIdx.x] = threadIdx.x < p.cout ? p.bias[threadIdx.x] : 0.f;
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    if (warp == 0) S1S2_TL(1);
    pdl_launch_dependents();                    // see conv_umma_kernel: the prologue above overlaps the previous kernel's tail
    if (warp != 0) pdl_wait();                  // the producer waits after it has requested its first weight tiles

    if (warp == 3) prefetch_next_weights(p.next_w, p.next_w_bytes, lane);      // (warp 3 has no other role)
    if (warp == 0) {
        // ================================================================= TMA producer
        // flat walk over (tile, chunk).  With two halo slots the NEXT halo tile is requested after weight tile kIssueTap
        // of the current chunk: by then the MMA warp (at most STAGES weight tiles behind) is done with the slot's
        // previous occupant, so the request never blocks the weight stream.
        // With three halo slots the next halo tile is simply requested before the current chunk's weights (its slot was
        // freed two chunks ago).
        constexpr int kStagesPerChunk = 9 / TPS;
        [[maybe_unused]] constexpr int kIssueStage = STAGES - 1 < kStagesPerChunk - 1 ? STAGES - 1 : kStagesPerChunk - 1;
        int s = 0, sh = 0;
        uint32_t ph = 0, phh = 0;
        auto load_halo = [&](int tile, int chunk) {
            const int tx = tile % tiles_x;
            const int ty = (tile / tiles_x) % tiles_y;
            const int n = tile / (tiles_x * tiles_y);
            mbar_wait(&hempty_bar[sh], phh ^ 1);
            if (elect_one()) {
                mbar_expect_tx(&hfull_bar[sh], L::kHaloBytes);
                tma_load_4d(halo_ring + sh * L::kHaloSlot, &p.tmap_a, &hfull_bar[sh], chunk * KBOX, (tx << 3) - 1, (ty << 5) - 1, n);
            }
            __syncwarp();
            if (++sh == kPxHaloSlots) { sh = 0; phh ^= 1; }
        };
        int tile = blockIdx.x, chunk = 0;
        auto advance = [&](int& t, int& c) { if (++c == p.chunks) { c = 0; t += gridDim.x; } };
        // One ring stage of weight tiles (TPS taps of one channel chunk); `wait` = the slot may still be in use.
        auto weight_stage = [&](int kcol, bool wait) {
            if (wait) mbar_wait(&empty_bar[s], ph ^ 1);
            if (elect_one()) {
                mbar_expect_tx(&full_bar[s], L::kStage);
#pragma unroll
                for (int j = 0; j < TPS; ++j)
                    tma_load_2d(stage_base + s * L::kStage + j * L::kABox, &p.tmap_b, &full_bar[s], kcol + j * p.tap_kstride, 0);
            }
            __syncwarp();
            if (++s == STAGES) { s = 0; ph ^= 1; }
        };
        // weight tiles of the first (tile, chunk) position go out before the wait on the previous kernel (conv_umma_kernel)
        int st0 = 0;
        if (HSLOTS >= 3 && tile < num_tiles)             // (the two-slot walk below interleaves halo requests with the stages)
            for (; st0 < kStagesPerChunk && st0 < STAGES; ++st0) weight_stage(st0 * TPS * p.tap_kstride, false);
        pdl_wait();
        S1S2_TL(2);
        if constexpr (HSLOTS >= 3) {
            // the halo cursor runs HSLOTS - 2 positions ahead of the weight stream (its slot was freed two positions earlier)
            constexpr int kAhead = HSLOTS - 2;
            int htile = tile, hchunk = 0;
            for (int d = 0; d < kAhead; ++d)
                if (htile < num_tiles) { load_halo(htile, hchunk); advance(htile, hchunk); }
            if (tile < num_tiles) {                      // the first position, peeled: its first weight stages are in flight
                if (htile < num_tiles) { load_halo(htile, hchunk); advance(htile, hchunk); }
                for (; st0 < kStagesPerChunk; ++st0) weight_stage(st0 * TPS * p.tap_kstride, true);
                advance(tile, chunk);
            }
            while (tile < num_tiles) {
                if (htile < num_tiles) { load_halo(htile, hchunk); advance(htile, hchunk); }
                int kcol = chunk * KBOX;
                for (int st = 0; st < kStagesPerChunk; ++st, kcol += TPS * p.tap_kstride) {
                    mbar_wait(&empty_bar[s], ph ^ 1);
                    if (elect_one()) {
                        mbar_expect_tx(&full_bar[s], L::kStage);
#pragma unroll
                        for (int j = 0; j < TPS; ++j)
                            tma_load_2d(stage_base + s * L::kStage + j * L::kABox, &p.tmap_b, &full_bar[s], kcol + j * p.tap_kstride, 0);
                    }
                    __syncwarp();
                    if (++s == STAGES) { s = 0; ph ^= 1; }
                }
                advance(tile, chunk);
            }
        } else {
            if (tile < num_tiles) load_halo(tile, 0);
            while (tile < num_tiles) {
                int ntile = tile, nchunk = chunk;
                advance(ntile, nchunk);
                int kcol = chunk * KBOX;
                for (int st = 0; st < kStagesPerChunk; ++st, kcol += TPS * p.tap_kstride) {
                    mbar_wait(&empty_bar[s], ph ^ 1);
                    if (elect_one()) {
                        mbar_expect_tx(&full_bar[s], L::kStage);
#pragma unroll
                        for (int j = 0; j < TPS; ++j)
                            tma_load_2d(stage_base + s * L::kStage + j * L::kABox, &p.tmap_b, &full_bar[s], kcol + j * p.tap_kstride, 0);
                    }
                    __syncwarp();
                    if (++s == STAGES) { s = 0; ph ^= 1; }
                    if (st == kIssueStage && ntile < num_tiles) load_halo(ntile, nchunk);
                }
                tile = ntile;
                chunk = nchunk;
            }
        }
    } else if (warp == 1) {
        // ================================================================= MMA issuer
        int s = 0, sh = 0;
        uint32_t ph = 0, phh = 0;
        int acc = 0;
        uint32_t acc_ph = 0;
        const uint32_t w0 = smem_u32(stage_base), h0 = smem_u32(halo_ring);
        uint64_t bdesc0 = static_cast<uint64_t>(1) << 16;            // pixel operand: rows kRowBytes apart, groups 10 rows apart
        bdesc0 |= static_cast<uint64_t>((kPxHaloW * kRowBytes) >> 4) << 32;
        bdesc0 |= static_cast<uint64_t>(1) << 46;
        bdesc0 |= (kRowBytes == 128 ? 2ull : (kRowBytes == 64 ? 4ull : 6ull)) << 61;
        for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
            mbar_wait(&tempty_bar[acc], acc_ph ^ 1);
            tc_fence_after();
            const uint32_t d_tmem = tmem_base + acc * kAccStride;
            for (int chunk = 0; chunk < p.chunks; ++chunk) {
                mbar_wait(&hfull_bar[sh], phh);
                S1S2_TL_ONCE(tl_a, 4);
                const uint32_t h_addr = h0 + sh * L::kHaloSlot;
                if constexpr (TPS == 1) {
                    for (int tap = 0; tap < 9; ++tap) {
                        mbar_wait(&full_bar[s], ph);
                        tc_fence_after();
                        const int ky = tap / 3, kx = tap - 3 * ky;
                        const uint32_t h_tap = h_addr + (ky * kPxHaloW + kx) * kRowBytes;
                        if (elect_one()) {
                            const uint64_t adesc = umma_smem_desc<kRowBytes>(w0 + s * L::kStage);                    // M side: weights
                            const uint64_t bdesc = bdesc0 | static_cast<uint64_t>((h_tap & 0x3FFFFu) >> 4);       // N side: pixels
#pragma unroll
                            for (int k = 0; k < KBOX / 16; ++k)
                                umma_f16(d_tmem, adesc + 2 * k, bdesc + 2 * k, kIdesc, (chunk | tap | k) != 0 ? 1u : 0u);
                            umma_commit(&empty_bar[s]);
                            if (tap == 8) {
                                umma_commit(&hempty_bar[sh]);
                                if (chunk == p.chunks - 1) umma_commit(&tfull_bar[acc]);
                            }
                        }
                        __syncwarp();
                        if (++s == STAGES) { s = 0; ph ^= 1; }
                    }
                } else {
                    // one kernel row (three taps) per stage: one barrier round trip, then a straight-line block of
                    // 3 * KBOX/16 MMAs with compile-time descriptor offsets
                    const uint64_t bdesc_c = bdesc0 | static_cast<uint64_t>((h_addr & 0x3FFFFu) >> 4);
#pragma unroll
                    for (int t0 = 0; t0 < 9; t0 += TPS) {
                        mbar_wait(&full_bar[s], ph);
                        S1S2_TL_ONCE(tl_b, 5);
                        tc_fence_after();
                        if (elect_one()) {
                            const uint64_t adesc_s = umma_smem_desc<kRowBytes>(w0 + s * L::kStage);
#pragma unroll
                            for (int j = 0; j < TPS; ++j) {
                                const int tap = t0 + j, ky = tap / 3, kx = tap - 3 * ky;
                                const uint64_t adesc = adesc_s + static_cast<uint64_t>((j * L::kABox) >> 4);
                                const uint64_t bdesc = bdesc_c + static_cast<uint64_t>(((ky * kPxHaloW + kx) * kRowBytes) >> 4);
#pragma unroll
                                for (int k = 0; k < KBOX / 16; ++k)
                                    umma_f16(d_tmem, adesc + 2 * k, bdesc + 2 * k, kIdesc, (tap | k) != 0 ? 1u : (chunk != 0 ? 1u : 0u));
                            }
                            umma_commit(&empty_bar[s]);
                            if (t0 + TPS == 9) {
                                umma_commit(&hempty_bar[sh]);
                                if (chunk == p.chunks - 1) umma_commit(&tfull_bar[acc]);
                            }
                        }
                        __syncwarp();
                        if (++s == STAGES) { s = 0; ph ^= 1; }
                    }
                }
                if (++sh == kPxHaloSlots) { sh = 0; phh ^= 1; }
            }
            acc ^= 1;
            if (acc == 0) acc_ph ^= 1;
        }
        S1S2_TL(6);
    } else if (warp >= 4) {
        // ================================================================= epilogue
        const int q = warp & 3;                     // TMEM lane quadrant = 32 output channels
        const int grp = (warp - 4) >> 2;            // epilogue warpgroup = accumulator = staging buffer
        const int et = q * 32 + lane;               // thread index 0..127 inside the group (= output channel in pass 1)
        const int acc = grp;
        uint32_t acc_ph = 0;
        uint8_t* sout = sout0 + grp * L::kStagingBuf;
        const int bar_id = 1 + grp;
        const bool issuer = (q == 0 && lane == 0);  // the group's TMA-store thread
        for (int tile = blockIdx.x + grp * gridDim.x; tile < num_tiles; tile += 2 * gridDim.x) {
            const int tx = tile % tiles_x;
            const int ty = (tile / tiles_x) % tiles_y;
            const int n = tile / (tiles_x * tiles_y);
            float s_up = 1.f, s_dn = 1.f;
            [[maybe_unused]] bool poisoned = false;      // invalid call for this patch (pack_input_kernel): the head writes NaN
            if (p.amax_in != nullptr) {
                const uint32_t bits = __ldg(p.amax_in + n);
                s_dn = range_scale(bits, s_up);
                poisoned = bits > 0x7F800000u;
            }

            mbar_wait(&tfull_bar[acc], acc_ph);
            if (warp == 4) S1S2_TL_ONCE(tl_c, 7);
            tc_fence_after();
            const StepCoef& sc = p.head.step;
            const size_t plane = static_cast<size_t>(p.H) * p.W;
            float amax = 0.f;
#pragma unroll 1
            for (int half = 0; half < 2; ++half) {
                if constexpr (MODE == MODE_STORE) {
                    if (issuer) bulk_wait_read0();                   // the group's previous stores have read the staging
                }
                named_bar_sync(bar_id, 128);                         // (HEAD: the previous pass 2 is done)
                // HEAD: this thread's pixel of pass 2; its state values are requested now so the global-load latency
                // hides behind pass 1
                const int px2 = et;
                const int x2 = (tx << 3) + (px2 & 7), y2 = (ty << 5) + half * 16 + (px2 >> 3);
                const bool inside = y2 < p.H;                     // H % 32 == 16: the tile's lower half is outside the image
                const size_t pix = static_cast<size_t>(y2) * p.W + x2;
                float xt4[kHeadOut] = {0.f, 0.f, 0.f, 0.f}, z4[kHeadOut] = {0.f, 0.f, 0.f, 0.f};
                if constexpr (MODE == MODE_HEAD) {
                    if (inside && sc.kind != STEP_NONE) {
#pragma unroll
                        for (int k = 0; k < kHeadOut; ++k) {
                            const size_t idx = (static_cast<size_t>(n) * kHeadOut + k) * plane + pix;
                            xt4[k] = p.head.x_t[idx];
                            if (sc.flags & STEP_FLAG_NOISE) z4[k] = p.head.noise[idx];
                        }
                        if (sc.flags & STEP_FLAG_PHILOX) philox_normal4(p.head, static_cast<uint32_t>(pix), static_cast<uint32_t>(n), z4);
                    }
                }
                if (q < NQ) {
                    // pass 1: this thread's channel, 128 pixels of the half tile -> staging[pixel][channel]
                    const uint32_t taddr = tmem_base + acc * kAccStride + (static_cast<uint32_t>(q * 32) << 16) + half * 128;
                    const float b = sbias[et] * s_dn;
                    uint8_t* sub = sout + q * L::kSubBytes + (lane & 7) * 2;
                    const int cchunk = lane >> 3;                    // 16-byte chunk of this channel inside a 64-byte row
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        uint32_t r[32];
                        tmem_ld32(taddr + j * 32, r);
                        tmem_ld_wait();
#pragma unroll
                        for (int i = 0; i < 32; ++i) {
                            const int px = j * 32 + i;
                            const float v = fmaxf(__uint_as_float(r[i]) + b, 0.f);
                            uint16_t hv;
                            asm("cvt.rn.satfinite.f16.f32 %0, %1;" : "=h"(hv) : "f"(v));
                            *reinterpret_cast<uint16_t*>(sub + px * 64 + ((cchunk ^ ((px >> 1) & 3)) << 4)) = hv;
                        }
                    }
                }
                if (half == 1) {                                     // accumulator drained: MMAs of tile+2 may start
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&tempty_bar[acc]);
                }
                if constexpr (MODE == MODE_STORE) {
                    fence_proxy_async_smem();
                    named_bar_sync(bar_id, 128);
                    if (issuer) {
#pragma unroll
                        for (int c = 0; c < NQ; ++c)
                            tma_store_4d(&p.tmap_out, sout + c * L::kSubBytes, c * 32, tx << 3, (ty << 5) + half * 16, n);
                        bulk_commit();
                    }
                } else {
                    named_bar_sync(bar_id, 128);                     // hidden half tile complete in shared memory
                    // pass 2: this thread's pixel: outc over the 96 hidden channels, scheduler update, writes
                    const int px = px2, x = x2, y = y2;
                    float o[kHeadOut];
#pragma unroll
                    for (int k = 0; k < kHeadOut; ++k) o[k] = p.head.b[k] * s_dn;
                    const int sw = (px >> 1) & 3;
#pragma unroll
                    for (int c = 0; c < NQ; ++c) {
#pragma unroll
                        for (int g = 0; g < 4; ++g) {
                            const uint4 v = *reinterpret_cast<const uint4*>(sout + c * L::kSubBytes + px * 64 + ((g ^ sw) << 4));
                            const uint32_t w4[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
                            for (int e = 0; e < 4; ++e) {
                                const float2 hh = __half22float2(*reinterpret_cast<const __half2*>(&w4[e]));
                                const int ch = c * 32 + g * 8 + e * 2;
#pragma unroll
                                for (int k = 0; k < kHeadOut; ++k) {
                                    o[k] = fmaf(hh.x, p.head.w[k * kHidden + ch], o[k]);
                                    o[k] = fmaf(hh.y, p.head.w[k * kHidden + ch + 1], o[k]);
                                }
                            }
                        }
                    }
                    float res[kHeadOut];
                    if (inside) {
#pragma unroll
                    for (int k = 0; k < kHeadOut; ++k) {
                        const size_t idx = (static_cast<size_t>(n) * kHeadOut + k) * plane + pix;
                        const float pr = poisoned ? __uint_as_float(kAmaxPoison) : o[k] * s_up;
                        if (p.head.pred_out != nullptr) p.head.pred_out[idx] = pr;
                        res[k] = pr;
                        if (sc.kind != STEP_NONE) {
                            const float xt = xt4[k];
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
                            if (sc.flags & (STEP_FLAG_NOISE | STEP_FLAG_PHILOX)) xn = __fadd_rn(xn, __fmul_rn(sc.c4, z4[k]));
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
                        *reinterpret_cast<uint4*>(rec) = v;
                        uint2 u;
                        u.x = pack_half2_sat(hi[0], hi[1]);
                        u.y = pack_half2_sat(hi[2], hi[3]);
                        *reinterpret_cast<uint2*>(rec + 12) = u;
                    }
                    }
                }
            }
            if constexpr (MODE == MODE_HEAD) {
                if (sc.kind != STEP_NONE && p.head.amax_out != nullptr) {
                    const uint32_t mx = __reduce_max_sync(0xffffffffu, __float_as_uint(amax));   // one image per tile
                    if (lane == 0) atomicMax(p.head.amax_out + n, mx);
                }
            }
            acc_ph ^= 1;                            // this group's accumulator is used once per two tiles
        }
        if (warp == 4) S1S2_TL(8);
        if constexpr (MODE == MODE_STORE) {
            if (issuer) bulk_wait_read0();       // see conv_umma_kernel: writes are ordered by grid completion
        }
        if (warp == 4) S1S2_TL(9);
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        tmem_dealloc<512>(tmem_base);
    }
    if (warp == 0) S1S2_TL(10);
    S1S2_TL_GRID(13, atomicMax);
}

}  // namespace s1s2
