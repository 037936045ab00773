// Shared declarations for the FaceEnhanceNet sm_100a kernels.
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <cuda.h>
#include <stdint.h>

namespace fen {

typedef __nv_bfloat16 bf16;

// ------------------------------------------------------------------ geometry of the conv kernel
constexpr int kC = 64;                               // feature channels (K per tap)
constexpr int kTileM = 128;                          // output pixels per UMMA tile (TMEM lanes)
constexpr int kStripW = 64;                          // image columns handled by one strip
constexpr int kPitch = kStripW + 2;                  // strip row pitch in shared memory (halo cols)
constexpr int kBoxRows = 4;                          // image rows per TMA box / ring slot
constexpr int kBoxPx = kBoxRows * kPitch;            // 264 pixels
constexpr int kSlotBytes = kBoxPx * kC * 2;          // 33792 B (= 33 * 1024: swizzle-atom aligned)
constexpr int kRingSlots = 3;                        // + 1 mirror slot after the last one
constexpr int kRingBytes = (kRingSlots + 1) * kSlotBytes;
constexpr int kMaxShift = 2 * kPitch + 2;            // largest tap offset in the strip-linear space

// Fixed-point accumulation of the squeeze-and-excitation pool sums: 64-bit integer atomics are associative, so the sums
// (and with them every output of the forward pass) do not depend on the order in which warps and CTAs arrive.
// Scale 2^24: |sum| < 2^39 fits, resolution 6e-8 (sums of bf16 activations over <= 2^24 pixels).
constexpr float kHsScale = 16777216.f;
__device__ __forceinline__ void hs_add(long long* p, float v) {
  atomicAdd(reinterpret_cast<unsigned long long*>(p), static_cast<unsigned long long>(__float2ll_rn(v * kHsScale)));
}
__device__ __forceinline__ float hs_to_float(long long v) { return __ll2float_rn(v) * (1.f / kHsScale); }
// The same for the small reductions of the BACKWARD pass (PReLU slope, SE matrix and SE dot-product gradients: sums of many
// tiny terms from many CTAs).  Scale 2^44: resolution 6e-14, |sum| < 2^19; integer atomics make every parameter gradient
// independent of the order of arrival (the reference trains with cudnn.deterministic = True, scripts/train.py:52-53).
constexpr float kGsScale = 17592186044416.f;
__device__ __forceinline__ void gs_add(long long* p, float v) {
  atomicAdd(reinterpret_cast<unsigned long long*>(p), static_cast<unsigned long long>(__float2ll_rn(v * kGsScale)));
}
__device__ __forceinline__ float gs_to_float(long long v) { return float(__ll2double_rn(v) * (1.0 / 17592186044416.0)); }

enum EpilogueKind : int {
  kEpiPrelu = 0,     // out = prelu(acc + bias)                       (RCAB conv1)
  kEpiSum = 1,       // out = acc + bias ; per-image channel sums     (RCAB conv2 -> SE pool)
  kEpiResidual = 2,  // out = acc + bias + residual                   (group conv, conv_after_body)
  kEpiShuffle = 3,   // out[2y+i,2x+j,c] = prelu(acc + bias), sub-pixel = blockIdx.y (upsample conv)
  kEpiLast = 4,      // out_f32 NCHW = [clamp](acc + bias + bicubic_x4(lr))   (conv_last)
  kEpiBias = 5,      // out = acc + bias
  // backward pass (fen_step_host.cuh): data-gradient convolutions with the element-wise backward that follows fused in
  kEpiGate = 6,      // PReLU backward: out = acc * (z > 0 ? 1 : slope), sums[c] += acc * min(z, 0)
                     //   (z > 0 comes from `mask_in`, the sign bits saved by the forward; min(z, 0) = a / slope with
                     //    a = `residual` = the saved post-PReLU activation; sums = the [64] slope gradient)
  kEpiDot = 7,       // out = acc (+ residual if given); sums[n][c] += out * aux   (SE backward: sum dx' * o per image)
};

struct ConvParams {
  int B, H, W;            // input spatial size (W a multiple of 64)
  int strips;             // W / 64
  int tiles_per_seg;      // ceil(H * 66 / 128); a segment = one strip of one image
  int total_tiles;        // B * strips * tiles_per_seg
  int tiles_per_cta;
  int epi;                // EpilogueKind
  int training;           // kEpiLast: no clamp when non-zero
  const float* bias;      // [gridDim.y][N]
  const float* slope;     // [64] PReLU slopes or nullptr
  const bf16* residual;   // NHWC, same shape as out (kEpiResidual; kEpiGate: saved activation; kEpiDot: optional)
  const bf16* aux;        // NHWC, same shape as out (kEpiDot)
  bf16* out;              // NHWC bf16 output
  float* sums;            // [B][64] fp32, accumulated with atomics (kEpiSum, kEpiDot); [64] for kEpiGate
  long long* sums64;      // optional: accumulate there in fixed point instead of `sums`: deterministic (kEpiSum: hs_add, the
                          //   forward's scale; kEpiGate / kEpiDot: gs_add, the gradient scale)
  uint32_t* mask_out;     // kEpiPrelu / kEpiShuffle, optional: bit c of word [2 * output pixel + column half] = (pre-activation
                          //   of channel 32 * half + c > 0) - what the PReLU backward needs for slopes of any sign
  const uint32_t* mask_in;  // kEpiGate: those words of the activation being differentiated
  const float* lr;        // [B][3][H/4][W/4] fp32 network input (kEpiLast)
  float* out_f32;         // [B][3][H][W] fp32 (kEpiLast), optional when out_u8 is given
  uint8_t* out_u8;        // [B][H][W][3] uint8 = trunc(clip(out * 255, 0, 255)) (kEpiLast), optional
  int bgr;                // out_u8 channel order B, G, R
  int unshuffle;          // kEpiGate, != 0: `out` is written as the four PixelShuffle sub-pixel planes [4][B][H/2][W/2][64]
                          //   (plane 2 (y & 1) + (x & 1)): the gradient of the upsample convolution's pre-shuffle output
  long long* dbg;         // optional [gridDim.x][8] cycle counters (developer builds), else nullptr
};

}  // namespace fen
