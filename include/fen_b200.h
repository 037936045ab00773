/* fen_b200.h - C ABI of the B200-native FaceEnhanceNet forward path.
 *
 * The reference (tomasz-pres/face-super-resolution) has no FFI layer: its boundary for this path is
 * the Python nn.Module API.  Every entry point below replaces one piece of that boundary; the
 * Python host (face-super-resolution_b200/model.py, data.py) binds them with ctypes and keeps the
 * reference's names, arguments and error behaviour.  INTEGRATION.md shows the binding a maintainer
 * of the reference would add.
 *
 * Conventions: plain pointers and sizes only (all pointers are DEVICE pointers unless named
 * host_*), `stream` is a cudaStream_t passed as void*, every function returns 0 on success or a
 * negative FEN_E* code, never throws and never allocates device memory (the caller passes a
 * workspace sized by fen_forward_workspace_bytes).  fen_last_error() returns a thread-local
 * message for the last failure.  There is no CPU fallback: without a sm_100 device every compute
 * entry point returns FEN_ENODEV.
 */
#ifndef FEN_B200_H
#define FEN_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FEN_ABI_VERSION 1

#define FEN_OK 0
#define FEN_EINVAL (-1)   /* bad argument / unsupported configuration (e.g. channels != 64) */
#define FEN_ENODEV (-2)   /* no sm_100 CUDA device */
#define FEN_ENOMEM (-3)   /* workspace too small */
#define FEN_ECUDA (-4)    /* CUDA runtime / driver error, see fen_last_error() */

/* Model hyper-parameters; mirrors FaceEnhanceNetConfig (reference src/models/custom.py:22-43).
 * Supported by the kernels: num_channels == 64, kernel_size == 3, scale_factor == 4,
 * in_channels == out_channels == 3, SE hidden width max(64 / reduction_ratio, 8) <= 64. */
typedef struct fen_config {
  int32_t num_channels;
  int32_t num_groups;
  int32_t blocks_per_group;
  int32_t reduction_ratio;
  int32_t scale_factor;
  float res_scale;
} fen_config;

int fen_abi_version(void);
const char* fen_last_error(void);

/* Number of fp32 elements of the flat parameter vector: the state_dict tensors of
 * FaceEnhanceNet concatenated in registration order (custom.py:88-124). */
int64_t fen_param_count(const fen_config* cfg);

/* Bytes of the packed (kernel-ready) weight blob. */
int64_t fen_packed_bytes(const fen_config* cfg);

/* Replaces nothing in the reference (weights are consumed in place by cuDNN there): turns the fp32
 * OIHW state_dict tensors (flat vector, see fen_param_count) into the layout the kernels read -
 * bf16 [tap][Cout][Cin] K-major conv weights (upsample convs permuted to PixelShuffle sub-pixel
 * groups, conv_last padded to 16 outputs), fp32 biases / PReLU slopes / SE matrices.
 * Call after construction, load_state_dict and every optimiser step. */
int fen_pack_weights(const fen_config* cfg, const float* params, void* packed, void* stream);

/* Workspace bytes fen_forward needs for a batch of B images of H x W (LR size). */
int64_t fen_forward_workspace_bytes(const fen_config* cfg, int B, int H, int W);

/* Replaces FaceEnhanceNet.forward (reference src/models/custom.py:147-190).
 *   x        [B,3,H,W]   fp32 NCHW in [0,1]
 *   out      [B,3,4H,4W] fp32 NCHW; clamped to [0,1] unless `training` (custom.py:187-188)
 *   se_out   optional [B, num_groups*blocks_per_group, 64] fp32: the channel-attention scales of
 *            every RCAB (what get_attention_maps, custom.py:192-230, returns); may be NULL
 * Any H, W >= 1 (the network is fully convolutional; the demo feeds sizes up to 128 x 128, app/demo.py:247-251).
 * 64-column inputs (the benchmark's 64 x 64) run the whole residual body in one persistent kernel; other widths
 * take one convolution launch per layer. */
int fen_forward(const fen_config* cfg, const void* packed, const float* x, float* out, int B, int H,
                int W, int training, void* workspace, int64_t workspace_bytes, float* se_out,
                void* stream);

/* fen_forward in eval mode followed by the evaluation scripts' to_numpy (scripts/test_model.py:176-190,
 * scripts/compare_two_models.py:150-179): out_u8 [B,4H,4W,3] uint8 HWC = trunc(clip(sr * 255, 0, 255)), channel order
 * B,G,R when `bgr`.  The conversion runs in the epilogue of conv_last: the fp32 SR tensor is never written.
 * Bit-identical to fen_sr_to_u8(fen_forward(x)). */
int fen_forward_u8(const fen_config* cfg, const void* packed, const float* x, uint8_t* out_u8, int bgr, int B, int H,
                   int W, void* workspace, int64_t workspace_bytes, void* stream);

/* Debug / parity taps: copies of intermediate NHWC bf16 feature maps of the LAST fen_forward on this
 * workspace.  which: 0 = conv_first output, 1 = body output (after conv_after_body + long skip),
 * 2 = upsample stage 0 output, 3 = upsample stage 1 output, 4 = residual stream after group g
 * (g = `index`).  Returns the tap's byte size, or a negative error. */
int64_t fen_forward_tap(const fen_config* cfg, const void* workspace, int B, int H, int W, int which,
                        int index, const void** ptr);

/* Replaces the LR generator of src/data: cv2.resize(hr, (W/4, H/4), interpolation=cv2.INTER_CUBIC)
 * (reference src/data/dataset.py:292-296, src/data/prepare_data.py:36-39) followed, optionally, by
 * to_tensor (src/data/transforms.py:260-279).  Integer arithmetic, bit-exact.
 *   hr       [B,H,W,C] uint8 HWC, H and W multiples of 4, C in 1..4
 *   lr_u8    optional [B,H/4,W/4,C] uint8 HWC
 *   lr_f32   optional [B,C,H/4,W/4] fp32 CHW, value / 255.0f */
int fen_lr_from_hr_u8(const uint8_t* hr, uint8_t* lr_u8, float* lr_f32, int B, int H, int W, int C,
                      void* stream);

/* Replaces the float LR generation of the trainer and of the evaluation scripts:
 * F.interpolate(hr, scale_factor=0.25, mode='bicubic', align_corners=False) (reference
 * src/training/trainer.py:416-421, also :568-573, :798-803) and generate_lr of scripts/test_model.py:139-156
 * (the same taps, then clip(v * 255, 0, 255) truncated to uint8).  At the exact /4 ratio the 16 taps are
 * w_i w_j, w = [-3, 19, 19, -3] / 32, on pixels 4y .. 4y+3 x 4x .. 4x+3; fp32, no rounding, no clamp.
 *   hr      [B,C,H,W] fp32 NCHW, H and W multiples of 4
 *   lr_f32  optional [B,C,H/4,W/4] fp32 NCHW (the trainer's tensor; tolerance vs PyTorch 2.4e-7)
 *   lr_u8   optional [B,H/4,W/4,C] uint8 HWC, trunc(clip(v * 255, 0, 255)); channel order reversed
 *           (RGB -> BGR, cv2 convention of the scripts) when bgr != 0 */
int fen_lr_from_hr_f32(const float* hr, float* lr_f32, uint8_t* lr_u8, int B, int C, int H, int W, int bgr,
                       void* stream);

/* Replaces to_numpy of the evaluation scripts (reference scripts/test_model.py:176-190,
 * scripts/compare_two_models.py:150-179, app/demo.py:180-222): network output fp32 NCHW in [0,1] ->
 * uint8 HWC, trunc(clip(v * 255, 0, 255)), channel order reversed when bgr != 0.
 *   sr   [B,C,H,W] fp32      out  [B,H,W,C] uint8 */
int fen_sr_to_u8(const float* sr, uint8_t* out, int B, int C, int H, int W, int bgr, void* stream);

/* ---- Stage-1 training step, loss / optimiser side.  The backward pass of the network
 * is fen_forward_train / fen_backward below; these are the pieces of Trainer._train_epoch that follow it.
 *
 * Scratch for the deterministic two-stage reductions below. */
int64_t fen_train_workspace_bytes(int64_t n);

/* Replaces nn.L1Loss(reduction='mean') (reference src/losses/combined.py:38-47, summed into total_loss at
 * :148-177 with weight 1) and its backward: loss[0] = mean |sr - hr| over n elements (device scalar),
 * dsr[i] = sign(sr[i] - hr[i]) / n (optional; what loss.backward() hands to the network output). */
int fen_l1_loss(const float* sr, const float* hr, int64_t n, float* loss, float* dsr, void* workspace,
                int64_t workspace_bytes, void* stream);

/* Replaces the validation metric Trainer._compute_psnr (reference src/training/trainer.py:621-628) and
 * psnr() of src/evaluation/metrics.py:17-34: psnr[0] = 10 log10(data_range^2 / mean((pred - target)^2)) over
 * ALL n elements of the batch (+inf when the mean squared error is 0); device scalar, deterministic. */
int fen_psnr(const float* pred, const float* target, int64_t n, float data_range, float* psnr, void* workspace,
             int64_t workspace_bytes, void* stream);

/* Global L2 norm of the flat gradient (what clip_grad_norm_ computes, reference src/training/trainer.py:490-496);
 * norm_out is a device scalar.  After a data-parallel all-reduce every rank holds the same gradient, so the norm
 * needs no second collective. */
int fen_grad_norm(const float* grads, int64_t n, float* norm_out, void* workspace, int64_t workspace_bytes,
                  void* stream);

/* Replaces torch.nn.utils.clip_grad_norm_(params, max_norm) followed by torch.optim.AdamW.step()
 * (reference src/training/trainer.py:217-221, 490-503): g *= min(1, max_norm / (total_norm + 1e-6)) (no
 * clipping when max_norm <= 0), then the decoupled-weight-decay Adam update with bias correction for
 * `step` (1-based).  total_norm is the device scalar of fen_grad_norm; one fused pass over the 4 arrays. */
int fen_clip_adamw_step(float* params, const float* grads, float* exp_avg, float* exp_avg_sq, int64_t n,
                        const float* total_norm, float max_norm, float lr, float beta1, float beta2, float eps,
                        float weight_decay, int step, void* stream);

/* Replaces ssim() / SSIMLoss (reference src/losses/ssim_loss.py:44-98, 166-226; callers Trainer._compute_ssim,
 * src/training/trainer.py:630-634, evaluation/metrics.py:77, CombinedLoss src/losses/combined.py:134-138): SSIM of
 * pred vs target, both [B,C,H,W] fp32, Gaussian window g g^T (window_1d: the ws normalised 1-D weights, HOST pointer;
 * ws odd, <= 11), zero padding ws / 2, C1 = (K1 range)^2, C2 = (K2 range)^2.
 *   per_image  optional [B]: mean of the SSIM map per image (size_average = False)
 *   mean       optional [1]: mean over everything            (size_average = True)
 *   grad_pred  optional [B,C,H,W]: d mean / d pred (the loss 1 - ssim has the opposite sign; per-image means: times B)
 * One pass over pred and target (the five filtered maps stay in shared memory), a second pass for the gradient. */
int64_t fen_ssim_workspace_bytes(int B, int C, int H, int W, int want_grad);
int fen_ssim(const float* pred, const float* target, int B, int C, int H, int W, const float* window_1d, int window_size,
             float c1, float c2, float* per_image, float* mean, float* grad_pred, void* workspace,
             int64_t workspace_bytes, void* stream);

/* ---- Stage-1 training step, network side (BASELINE config 5): what sr = model(lr) in train()
 * mode and loss.backward() do in Trainer._train_epoch (reference src/training/trainer.py:458-505) over
 * FaceEnhanceNet.forward (src/models/custom.py:147-190) and its blocks (src/models/blocks.py:75-263).
 *
 * Bytes of the transposed / tap-flipped weight blob the data-gradient convolutions read. */
int64_t fen_packed_bwd_bytes(const fen_config* cfg);

/* fp32 state_dict tensors (flat vector, see fen_param_count) -> bf16 [tap][Cin][Cout] weights with the taps
 * flipped (every 64-channel convolution; the PixelShuffle convolutions as four sub-pixel groups) and the fp32
 * conv_last weights in the layout of its data-gradient kernel.  Call after every parameter update. */
int fen_pack_weights_bwd(const fen_config* cfg, const float* params, void* packed_bwd, void* stream);

/* Bytes of the step workspace for a batch of B LR images of H x W: the activations fen_forward_train keeps
 * (per RCAB its input, h and o, the group outputs, both upsample stages) plus the gradient buffers. */
int64_t fen_step_workspace_bytes(const fen_config* cfg, int B, int H, int W);

/* FaceEnhanceNet.forward in train() mode (no clamp, custom.py:187-188) that keeps what fen_backward needs in
 * `step_workspace`.  Same arguments as fen_forward otherwise; runs one tcgen05 convolution launch per layer. */
int fen_forward_train(const fen_config* cfg, const void* packed, const float* x, float* out, int B, int H, int W,
                      void* step_workspace, int64_t step_workspace_bytes, void* stream);

/* loss.backward() through the network: dout [B,3,4H,4W] fp32 = d loss / d output of the LAST fen_forward_train on
 * this workspace (same x, B, H, W) -> grads, fp32, flat, in the order of fen_param_count (what .grad of every
 * state_dict tensor holds after backward(); overwritten, not accumulated).  PReLU slopes must be > 0 (the saved
 * activations are post-PReLU).  Data gradients and saved activations are bf16, parameter gradients fp32. */
int fen_backward(const fen_config* cfg, const void* packed, const void* packed_bwd, const float* x, const float* dout,
                 float* grads, int B, int H, int W, void* step_workspace, int64_t step_workspace_bytes, void* stream);

/* fen_backward cut into stages, in the order the backward runs, for a data-parallel trainer that all-reduces finished
 * slices of the flat gradient while later stages still compute (the reference has no multi-GPU code to
 * cite).  fen_backward_num_stages = num_groups + 2:
 *   stage 0: conv_last, both upsample stages, data gradient of conv_after_body; stage 1 + k: residual group G - 1 - k;
 *   stage G + 1: long skip + conv_first.
 * fen_backward_stage_range gives the slice [begin, begin + count) of `grads` that is COMPLETE after a stage.  count may
 * be 0: the weight gradients of the 64 -> 64 convolutions are computed in batched launches after some groups only
 * (after the upper half, after group 1, after group 0), which then complete the slices of several groups at once.
 * Stages must run in order on one stream; stage 0 zeroes `grads`. */
int fen_backward_num_stages(const fen_config* cfg);
int fen_backward_stage_range(const fen_config* cfg, int stage, int64_t* begin, int64_t* count);
int fen_backward_stages(const fen_config* cfg, const void* packed, const void* packed_bwd, const float* x,
                        const float* dout, float* grads, int B, int H, int W, void* step_workspace,
                        int64_t step_workspace_bytes, int stage_begin, int stage_end, void* stream);

/* One 3x3 / pad-1 convolution with 64 input and 64 output channels on NHWC bf16 tensors
 * (the RCAB building block, reference src/models/blocks.py:122-130): out = epilogue(conv(x) + bias).
 *   w_packed [9][64][64] bf16 (tap, cout, cin), as produced by fen_pack_conv3x3
 *   epilogue 0: prelu(slope)  1: none + channel sums into sums[B][64]  2: + residual  5: none
 * Used by the single-layer parity tests and the RCAB micro-benchmark. */
int fen_conv3x3_c64(const void* x, const void* w_packed, const float* bias, const float* slope,
                    const void* residual, float* sums, void* out, int B, int H, int W, int epilogue,
                    void* stream);

/* fp32 OIHW [cout][cin=64][3][3] -> bf16 [tap][cout_pad][64]; rows >= cout are zero. */
int fen_pack_conv3x3(const float* w_oihw, int cout, int cout_pad, void* w_packed, void* stream);

/* Kernels launched by the last fen_forward / fen_conv3x3_c64 / fen_lr_from_hr_u8 call of this thread. */
int fen_last_launch_count(void);

/* Measurement hooks for bench.py's roofline: fen_profile_body(1) makes fen_forward record CUDA events
 * on its launch stream around the persistent body kernel (all 64->64 convolutions of the residual
 * body); fen_last_body_ms() waits for the last one and returns its duration (-1 if none). */
int fen_profile_body(int enable);
float fen_last_body_ms(void);

#ifdef __cplusplus
}
#endif
#endif /* FEN_B200_H */
