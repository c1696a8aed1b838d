/* libs1s2_b200.so -- C ABI of the B200-native S1->S2 diffusion sampling path.
 *
 * The reference (ChenghanXia/S1-to-S2_Super-Resolution_Project-Code) is pure Python/PyTorch and has no FFI; its
 * "operator interface" for this path is the Python callable `model(torch.cat([x_t, x_cond], 1), t_idx)` plus free
 * sampler functions taking `model`.  Each entry point below names the reference call site
 * it replaces; the Python mirror of those call sites (`s1s2_b200.UNetSmallB200`, `s1s2_b200.samplers`) binds these
 * symbols with ctypes (INTEGRATION.md shows the stub).
 *
 * Conventions
 *   - plain pointers and sizes only; every device pointer is caller-owned (e.g. torch tensor .data_ptr()).
 *   - all image tensors are float32, contiguous NCHW, exactly like the reference's tensors.
 *   - `stream` is a cudaStream_t passed as void* (0 = legacy default stream).  Work is stream-ordered and
 *     asynchronous unless stated; there is no CPU fallback: without a B200-class device every call fails.
 *   - return value 0 = success, non-zero = error (message via s1s2_last_error / s1s2_global_error);
 *     no C++ exception crosses the boundary.
 *   - a handle is bound to one device and is not thread-safe.
 */
#ifndef S1S2_B200_H
#define S1S2_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct s1s2_handle s1s2_handle;

enum {
    S1S2_OK = 0,
    S1S2_ERR_INVALID = 1,   /* bad argument / unsupported geometry */
    S1S2_ERR_CUDA = 2,      /* CUDA runtime / driver error */
    S1S2_ERR_STATE = 3      /* e.g. weights not loaded */
};

/* Scheduler-update kinds evaluated in the epilogue of the network's last convolution. */
enum {
    S1S2_STEP_NONE = 0,      /* plain model call */
    S1S2_STEP_EPS_DDIM = 1,  /* x0=(x-c0*e)/c1; x'=c2*x0+c3*e (+c4*z)   Evaluation/DDIM_Multi-step.py:129-133 */
    S1S2_STEP_V_DDIM = 2,    /* x0=c0*x-c1*v; e=c1*x+c0*v; x' as above  DDIM_Multi-step_v_Prediction.py:59-65,161-174 */
    S1S2_STEP_EPS_DDPM = 3,  /* x'=c2*(x-c3*e) (+c4*z)                  Evaluation/Limitation_Test.py:213-223 */
    S1S2_STEP_V_DDPM = 4     /* e=c1*x+c0*v first                        Limitation_Test_v_Prediction.py:214-225 */
};
enum {
    S1S2_STEP_FINAL = 1,     /* result = clamp(x0,0,1) (DDIM kinds) / clamp(x',0,1) (DDPM kinds) */
    S1S2_STEP_NOISE = 2,     /* add c4 * z, z taken from the step_noise argument */
    S1S2_STEP_PHILOX = 4     /* add c4 * z, z ~ N(0,1) generated in the kernel: Philox4x32-10 keyed by the seed of
                                s1s2_set_noise_seed, counter (pixel, patch_base + patch slot, noise_index).  Replaces the
                                per-step torch.randn_like of Limitation_Test.py:222 / DDIM_Multi-step_v_Prediction.py:172
                                when the caller supplies no z (a DDPM-1000 chain at batch 64 would need 64 GB of it);
                                statistically, not bitwise, equivalent to the reference's generator */
};

/* One model call + scheduler update.  The host computes the coefficients exactly as the reference would
 * (float32 torch arithmetic on the float32 alpha_bar table) so the fused update is bit-identical to the
 * reference's elementwise code given the same network output. */
typedef struct s1s2_step {
    int32_t t;          /* timestep fed to the network for this call */
    int32_t kind;       /* S1S2_STEP_* */
    int32_t flags;      /* S1S2_STEP_FINAL | S1S2_STEP_NOISE */
    int32_t noise_index;/* index into step_noise[...] when S1S2_STEP_NOISE is set */
    float c0, c1, c2, c3, c4;
} s1s2_step;

/* Replaces `UNetSmall(in_ch, out_ch, base_ch).to(device)` (Evaluation/DDIM_Multi-step.py:19-41; call site
 * DDIM_Multi-step_v_Prediction.py:263).  Allocates the fp16 NHWC activation arena for `max_batch` patches of
 * H x W on `device`.  Supported: in_ch=8, out_ch=4, base_ch=96, H and W multiples of 16. */
int s1s2_create(s1s2_handle** out, int device, int in_ch, int out_ch, int base_ch, int H, int W, int max_batch);

/* Replaces `model.load_state_dict(state, strict=True)` (DDIM_Multi-step_v_Prediction.py:264-271).  `names[i]` are
 * the reference's 34 state_dict keys, `ptrs[i]` float32 DEVICE pointers in the reference's layouts (Conv2d OIHW,
 * ConvTranspose2d IOHW), `numel[i]` their element counts.  Strict: every key must be present with the right size.
 * Weights are repacked to fp16 [cout][tap][cin]; the time plane is folded into two fp16 input planes. */
int s1s2_load_weights(s1s2_handle* h, int n, const char* const* names, const float* const* ptrs,
                      const int64_t* numel, void* stream);

/* Replaces `model(torch.cat([x_t, x_cond], dim=1), t_idx)` (Evaluation/DDIM_Multi-step.py:43-53,131).
 * xt_and_cond: f32[B,8,H,W]; t_idx: int64[B]; out: f32[B,4,H,W] (all device).  Timesteps must lie in [0, 2048] (the
 * fp16-exact integer range of the time planes; the reference uses T = 1000): a patch whose t is outside gets NaN output
 * instead of a silently rounded timestep (t_idx lives on the device, so the check cannot be a host error code). */
int s1s2_forward(s1s2_handle* h, const float* xt_and_cond, const int64_t* t_idx, float* out, int B, void* stream);

/* Replaces the whole sampling loop of ddpm_ddim_generate / ddim_multistep_eval[_v] / ddim_sample / sample_ddim_v /
 * ddpm_sample / sample_ddpm_v (call sites listed in SURVEY.md section 8 a3-a10): n_steps model calls with the
 * scheduler update fused into the last convolution's epilogue; no host synchronisation inside.
 *   steps      host array [n_steps]
 *   cond       f32[B,4,H,W] device
 *   x_init     f32[B,4,H,W] device; the state starts as x_init * init_scale
 *   step_noise f32[n_noise,B,4,H,W] device or NULL
 *   out        f32[B,4,H,W] device: the sampler state, holds the result when the stream drains
 *   tap_pred   f32[n_steps,B,4,H,W] device or NULL: network output (eps or v) of every call
 *   tap_x      f32[n_steps,B,4,H,W] device or NULL: state after every update */
int s1s2_sample(s1s2_handle* h, const s1s2_step* steps, int n_steps, const float* cond, const float* x_init,
                float init_scale, const float* step_noise, float* out, float* tap_pred, float* tap_x, int B,
                void* stream);

/* Seed of the in-kernel noise (S1S2_STEP_PHILOX) and the global patch id of batch slot 0, so that a patch draws the
 * same noise wherever it sits in a batch or on whichever GPU it lands. */
int s1s2_set_noise_seed(s1s2_handle* h, uint64_t seed, uint32_t patch_base);

/* Same as s1s2_sample with HOST buffers (pinned memory recommended): copies cond / x_init to the device, samples,
 * copies the result back and synchronises the stream.  This is the end-to-end entry a host-side caller of the
 * reference scripts would bind (they load .npz patches on the CPU, DDIM_Multi-step.py:104-111). */
int s1s2_sample_host(s1s2_handle* h, const s1s2_step* steps, int n_steps, const float* cond_host,
                     const float* x_init_host, float init_scale, float* out_host, int B, void* stream);

/* The same for N >= 1 patches held in host memory, processed in batches of `batch` (<= max_batch) through a three-stream
 * pipeline: the upload of batch i+1 and the download of batch i-1 run under the model calls of batch i (two staging sets on
 * the device).  This is the per-file loop of the reference drivers (DDIM_Multi-step.py:219-246: load .npz -> .to(device) ->
 * sample -> .cpu()) for a whole list of files.  cond_host / x_init_host / out_host: f32[N,4,H,W]; pinned memory recommended
 * (pageable memory serialises the copies).  Synchronises `stream` and its two internal copy streams before returning. */
int s1s2_sample_host_stream(s1s2_handle* h, const s1s2_step* steps, int n_steps, const float* cond_host,
                            const float* x_init_host, float init_scale, float* out_host, int N, int batch, void* stream);

/* Unit-normal initial noise for N patches keyed by GLOBAL patch id: Philox4x32-10, key = seed, counter = (element / 4,
 * patch id, tag), Box-Muller.  A patch draws the same noise on whichever rank / batch slot it lands, so a sharded scene
 * equals the single-GPU scene bit for bit.  (The reference seeds torch's generator per file, DDIM_Sweep.py:193,404; the
 * drivers keep doing that -- this entry serves the whole-scene path, which has no reference counterpart.)
 *   patch_ids int64[N] device; out f32[N, elems_per_patch] device, 16-byte aligned; elems_per_patch % 4 == 0; N <= 65535 */
int s1s2_patch_noise(int device, uint64_t seed, const int64_t* patch_ids, int N, int64_t elems_per_patch, float* out,
                     void* stream);

/* Tile extraction + per-patch normalisation (Patch.py:80-84,201-209,226-239) for the windows listed in `origins`.
 *   scene   f32[4,SH,SW] device (HH dB, HV dB, incidence deg, elevation m; may hold NaN/Inf)
 *   vmask   u8[SH,SW] device or NULL: extra validity (Patch.py:41-49's target / collocation terms); a pixel is
 *           valid when all four scene channels are finite and vmask (if given) is non-zero
 *   origins int32[N,2] device (row, col); a window that does not lie inside the scene yields an all-invalid patch
 *           (cond 0, mask 0, valid_ratio 0)
 *   cond    f32[N,4,ps,ps] device out; mask u8[N,ps,ps] device out; valid_ratio f32[N] device out (nullable) */
int s1s2_tile_extract(int device, const float* scene, const uint8_t* vmask, int SH, int SW, const int32_t* origins,
                      int N, int ps, float* cond, uint8_t* mask, float* valid_ratio, void* stream);

/* Patch.py's four window quality tests on the target (Patch.py:88-114,205-224) for the windows in `origins`.
 *   scene f32[Ci,SH,SW], target f32[4,SH,SW] (B2,B3,B4,B8 reflectance), colloc u8[SH,SW] or NULL: device
 *   thresholds[5] host: valid_ratio_threshold, variance_threshold, dark_thr, dark_max_ratio, texture_thr
 *   stats f32[N,8] device out: valid_ratio, var[0..3], dark_fraction, laplacian_var, code (0 keep, 1..4 = failed test) */
int s1s2_tile_filter(int device, const float* scene, int Ci, const float* target, const uint8_t* colloc, int SH, int SW,
                     const int32_t* origins, int N, int ps, const float* thresholds, float* stats, void* stream);

/* Uniform-weight overlap blend (not in the reference; definition in DESIGN.md): gather formulation, deterministic.
 *   preds f32[N,C,ps,ps] device; origins int32[N,2] device, sorted in Patch.py iteration order on a regular
 *   `stride` grid; canvas f32[C,SH,SW] and cover u8[SH,SW] device out. */
int s1s2_stitch(int device, const float* preds, const int32_t* origins, int N, int C, int ps, int stride, int SH,
                int SW, float* canvas, uint8_t* cover, void* stream);

/* The same with a separable per-pixel weight window[ly] * window[lx] (f32[ps] device, e.g. a Hann window that is positive
 * everywhere): canvas = sum(w * pred) / sum(w), accumulated in the same order.  window == NULL is s1s2_stitch. */
int s1s2_stitch_weighted(int device, const float* preds, const int32_t* origins, int N, int C, int ps, int stride, int SH,
                         int SW, const float* window, float* canvas, uint8_t* cover, void* stream);

/* Evaluation metrics of N predicted patches in one pass each (the drivers' per-file metric calls:
 * masked_mae / masked_mse / psnr / ssim_simple, Evaluation/DDIM_Multi-step.py:72-101; sam / ergas,
 * Evaluation_Updated/Evaluation_Pure_Generation.py:229-254) plus the per-channel error sums behind the dataset-level
 * pixel-weighted aggregation of the batched evaluators (channelwise_error_sums, Evaluation/Limitation_Test.py:118-133).
 *   pred, gt f32[N,C,HW] device; mask u8[N,HW] device or NULL (all valid); C <= 8
 *   out f64[N,24] device: [0..7] mae, mse, psnr, ssim_simple, sam, ergas, valid-pixel count, 0;
 *                         [8+c] sum |pred-gt| and [16+c] sum (pred-gt)^2 of channel c over the valid pixels (0 for c >= C) */
int s1s2_patch_metrics(int device, const float* pred, const float* gt, const uint8_t* mask, int N, int C, int HW,
                       double* out, void* stream);

/* Per-layer parity tap: converts the named activation of the LAST model call (fp16 NHWC arena view) to float32
 * NCHW.  Names follow the network's blocks: "inc", "down1", "down2", "down3", "up3", "conv3", "up2", "conv2",
 * "up1", plus the mid-block tensors "down1.0", "down2.0", "down3.0", "conv3.0", "conv2.0", "conv1.0" and the
 * packed input record "xin16".  With out_nchw == NULL only the shape (C, H, W) is returned. */
int s1s2_debug_activation(s1s2_handle* h, const char* name, float* out_nchw, int B, int* C, int* H, int* W,
                          void* stream);

/* Debug aid for the fp16 activation arena ("fp16 overflow"): for every activation view of the LAST
 * model call (order and names: s1s2_view_name) the number of stored fp16 elements at the saturation value of the conv
 * epilogues' cvt.rn.satfinite (|v| = 65504) or non-finite -- i.e. how many outputs a layer clamped.  Host-synchronous;
 * counts is a HOST array [n_out]; with counts == NULL only *n_views is returned. */
int s1s2_debug_saturation_count(s1s2_handle* h, int B, uint64_t* counts, int n_out, int* n_views, void* stream);
const char* s1s2_view_name(const s1s2_handle* h, int i);

/* Measurement aid for bench.py's roofline block: runs the denoiser `reps` times at batch B on whatever the arena
 * holds, with a CUDA event pair around every launch on `stream`, and returns the mean duration of each launch in
 * milliseconds (ms_out[i], i < n_out; launch order = the network's execution order, see s1s2_layer_name).  Host-
 * synchronous.  Returns the number of launches per model call through *n_layers. */
int s1s2_profile_layers(s1s2_handle* h, int B, int reps, float* ms_out, int n_out, int* n_layers, void* stream);
/* Measurement aid: launches layer `layer` alone `reps` times back to back at batch B (event-timed on `stream`, host-
 * synchronous) and returns the mean milliseconds per launch.  perf_mode bit 0 / bit 1 stop the kernel from re-loading
 * the activation / weight operand once its shared-memory ring is full, to attribute time to data movement; the
 * arena contents are garbage afterwards when perf_mode != 0. */
int s1s2_debug_loop_layer(s1s2_handle* h, int B, int layer, int reps, int perf_mode, float* ms_out, void* stream);
/* Measurement / test aid: the GEMM-N tile width (output channels per tile; 4 x Cout columns for the transposed convs) the
 * i-th launch of a model call uses at batch B -- narrower tiles are picked when the wide ones cannot fill the GPU (small
 * batches); every choice computes bit-identical results.  -1 for a bad argument. */
int s1s2_debug_tile_width(s1s2_handle* h, int layer, int B);
/* state_dict prefix of the i-th launch of a model call ("inc.0", "down1.0.0", ... "conv1.2"), NULL past the end. */
const char* s1s2_layer_name(const s1s2_handle* h, int i);

/* Number of kernels this library launched on behalf of `h` since creation (bench.py's gpu_launches). */
int64_t s1s2_launch_count(const s1s2_handle* h);

const char* s1s2_last_error(const s1s2_handle* h);
const char* s1s2_global_error(void);   /* for the handle-less entry points and failed s1s2_create */
void s1s2_destroy(s1s2_handle* h);
int s1s2_abi_version(void);

#ifdef __cplusplus
}
#endif
#endif /* S1S2_B200_H */
