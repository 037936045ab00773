// Backward pass of FaceEnhanceNet (the network side of the reference's Stage-1 step: loss.backward() in
// src/training/trainer.py:458-505 over src/models/custom.py:147-190 / blocks.py:75-263).
//
// Data gradients (dgrad) of every 64-channel convolution run on the SAME tcgen05 implicit-GEMM kernel as the
// forward (conv3x3_umma.cuh) with transposed + tap-flipped weights, with the PReLU backward (kEpiGate) and the
// squeeze-and-excitation dot product (kEpiDot) fused into its epilogue.  The kernels in this file are what that
// kernel cannot do: weight gradients (a contraction over PIXELS, K = B*H*W; the product kernel is the tcgen05 one in
// wgrad_umma.cuh - here are its two predecessors, kept for A/B runs: wgrad_c64_mma_kernel on the warp-level tensor
// cores and wgrad_c64_kernel with fp32 FMAs), the two 3-channel ends of the
// network, the PixelShuffle stages and the squeeze-and-excitation chain.  Activations and data gradients are NHWC
// bf16, parameter gradients fp32 (accumulated with atomics into the flat gradient vector, which the caller zeroes).
#pragma once
#include "conv3x3_umma.cuh"

namespace fen {

__device__ __forceinline__ void unpack4(uint2 v, float (&f)[4]) {
  f[0] = bf16lo(v.x); f[1] = bf16hi(v.x); f[2] = bf16lo(v.y); f[3] = bf16hi(v.y);
}
__device__ __forceinline__ void unpack8(uint4 v, float (&f)[8]) {
  f[0] = bf16lo(v.x); f[1] = bf16hi(v.x); f[2] = bf16lo(v.y); f[3] = bf16hi(v.y);
  f[4] = bf16lo(v.z); f[5] = bf16hi(v.z); f[6] = bf16lo(v.w); f[7] = bf16hi(v.w);
}
__device__ __forceinline__ uint4 pack8(const float (&f)[8]) {
  uint4 r;
  r.x = pack_bf16(f[0], f[1]); r.y = pack_bf16(f[2], f[3]); r.z = pack_bf16(f[4], f[5]); r.w = pack_bf16(f[6], f[7]);
  return r;
}

constexpr int kWgOutPitch = kC * 9 + 4;   // floats per staged output row of a weight-gradient flush (padded)
__device__ __forceinline__ void red_add_v4(float* dst, float4 v) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w)
               : "memory");
}

#ifdef FEN_DEV   // the fp32-FMA and mma.sync generations of the 64 -> 64 weight gradient: developer builds only
// --------------------------------------------------------------------------------------------------------------
// Weight gradient of a 64 -> 64 3x3 / pad-1 convolution:
//   dW[co][ci][ky][kx] += sum_{b,y,x} dY[b,y,x,co] * X[b,y+ky-1,x+kx-1,ci]        db[co] += sum dY[b,y,x,co]
// One CTA walks over "bands" (one image row of one 64-column strip): dY row (8 KB) and the three X rows with a
// one-pixel halo (25 KB) are staged in shared memory; thread (cb, ib) owns the 4 x 4 x 9 block
// co = 4 cb .. 4 cb + 3, ci = 4 ib .. 4 ib + 3, all taps (144 fp32 accumulators) for the whole kernel and
// flushes it once with atomics.  Output rows are co * co_mul + co_off (the PixelShuffle convolutions are four
// interleaved 64-row groups).
constexpr int kWgPx = kStripW;
__global__ void __launch_bounds__(256, 1)
wgrad_c64_kernel(const bf16* __restrict__ dY, const bf16* __restrict__ X, float* __restrict__ dW,
                 float* __restrict__ dB, int B, int H, int W, int co_mul, int co_off) {
  __shared__ __align__(16) bf16 sY[kWgPx * kC];
  __shared__ __align__(16) bf16 sX[3 * (kWgPx + 2) * kC];
  const int tid = threadIdx.x;
  const int cb = tid >> 4, ib = tid & 15;
  float acc[9][4][4];
  float bsum[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
  for (int t = 0; t < 9; ++t)
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
      for (int c = 0; c < 4; ++c) acc[t][a][c] = 0.f;
  const int strips = W / kWgPx;
  const int bands = B * H * strips;
  for (int band = blockIdx.x; band < bands; band += gridDim.x) {
    const int sidx = band % strips;
    const int y = (band / strips) % H;
    const int b = band / (strips * H);
    const int x0 = sidx * kWgPx;
    __syncthreads();   // previous band fully consumed
    {
      const uint4* src = reinterpret_cast<const uint4*>(dY + ((size_t(b) * H + y) * W + x0) * kC);
      uint4* dst = reinterpret_cast<uint4*>(sY);
      dst[tid] = __ldg(src + tid);
      dst[tid + 256] = __ldg(src + tid + 256);
      uint4* dx = reinterpret_cast<uint4*>(sX);
      for (int i = tid; i < 3 * (kWgPx + 2) * 8; i += 256) {
        const int chunk = i & 7, col = (i >> 3) % (kWgPx + 2), r = i / (8 * (kWgPx + 2));
        const int yy = y - 1 + r, xx = x0 - 1 + col;
        uint4 v = make_uint4(0u, 0u, 0u, 0u);
        if (yy >= 0 && yy < H && xx >= 0 && xx < W)
          v = __ldg(reinterpret_cast<const uint4*>(X + ((size_t(b) * H + yy) * W + xx) * kC) + chunk);
        dx[i] = v;
      }
    }
    __syncthreads();
#pragma unroll 1
    for (int px = 0; px < kWgPx; ++px) {
      float d[4];
      unpack4(*reinterpret_cast<const uint2*>(sY + px * kC + 4 * cb), d);
#pragma unroll
      for (int a = 0; a < 4; ++a) bsum[a] += d[a];
#pragma unroll
      for (int r = 0; r < 3; ++r)
#pragma unroll
        for (int kx = 0; kx < 3; ++kx) {
          float xv[4];
          unpack4(*reinterpret_cast<const uint2*>(sX + (r * (kWgPx + 2) + px + kx) * kC + 4 * ib), xv);
#pragma unroll
          for (int a = 0; a < 4; ++a)
#pragma unroll
            for (int c = 0; c < 4; ++c) acc[r * 3 + kx][a][c] = fmaf(d[a], xv[c], acc[r * 3 + kx][a][c]);
        }
    }
  }
#pragma unroll
  for (int a = 0; a < 4; ++a) {
    const int co = (4 * cb + a) * co_mul + co_off;
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      float* dst = dW + (size_t(co) * kC + 4 * ib + c) * 9;
#pragma unroll
      for (int t = 0; t < 9; ++t) atomicAdd(dst + t, acc[t][a][c]);
    }
    if (ib == 0) atomicAdd(dB + co, bsum[a]);
  }
}

// --------------------------------------------------------------------------------------------------------------
// Second generation of the same weight gradient on the warp-level tensor-core path (mma.sync m16n8k16, bf16 in,
// fp32 accumulate): per band D_tap[co][ci] += dY^T[co][px] * X_tap[px][ci], K = the 64 pixels of the band.  Both
// operands sit pixel-major in shared memory (the contraction index is the slow one), so both fragments come from
// ldmatrix.trans; rows are padded to 144 B to keep the 8 row addresses of an 8x8 matrix on distinct banks.
// Warp (mb, nb) owns co 32 mb .. +32, ci 16 nb .. +16 for all 9 taps: 36 m16n8 accumulator tiles = 144 registers.
// The dY fragments (2 ldmatrix.x4) are shared by the 9 taps, each tap needs ONE ldmatrix.x4 of X: 11 shared-memory
// fragment loads per 36 MMAs (a 16 x 32 warp tile needs 19).
#ifndef FEN_WG_UNROLL
#define FEN_WG_UNROLL 4
#endif
constexpr int kWgUnroll = FEN_WG_UNROLL;   // k-steps of a band unrolled together (developer knob)
constexpr int kWgPitch = kC + 8;   // elements per padded shared-memory row
__device__ __forceinline__ void ldmatrix_x4_trans(uint32_t addr, uint32_t (&r)[4]) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0, %1, %2, %3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void mma_bf16_16816(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, "
               "{%0, %1, %2, %3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
// Loads are double-buffered with cp.async (zero-fill outside the image); the accumulators leave through shared
// memory, 16 output rows at a time, as coalesced 128-bit reductions (red.global.add.v4.f32): a quarter of the
// instructions and an eighth of the L2 sectors of one scalar atomic per element.
constexpr int kWgStageElems = (kWgPx + 3 * (kWgPx + 2)) * kWgPitch;        // one buffer: dY row + 3 X rows
constexpr int kWgDynBytes = 2 * kWgStageElems * 2;                         // 75 456 B
static_assert(16 * kWgOutPitch * 4 <= kWgDynBytes, "output staging must fit the operand buffers");
__device__ __forceinline__ void cp_async_16(uint32_t dst, const void* src, int src_bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}
__global__ void __launch_bounds__(256, 1)
wgrad_c64_mma_kernel(const bf16* __restrict__ dY, const bf16* __restrict__ X, float* __restrict__ dW,
                     float* __restrict__ dB, int B, int H, int W, int co_mul, int co_off) {
  extern __shared__ __align__(16) uint8_t wg_smem[];
  const uint32_t smem_u32 = uint32_t(__cvta_generic_to_shared(wg_smem));
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int mb = warp >> 2, nb = warp & 3;
  float acc[9][4][4];   // [tap][2 * m16 tile + n8 tile][c fragment]
#pragma unroll
  for (int t = 0; t < 9; ++t)
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
      for (int c = 0; c < 4; ++c) acc[t][j][c] = 0.f;
  float bsum = 0.f;
  // per-lane ldmatrix row offsets (bytes, relative to the k-step / tap origin)
  const uint32_t a_off = uint32_t((((lane & 7) + ((lane >> 4) << 3)) * kWgPitch + 32 * mb + ((lane >> 3) & 1) * 8) * 2);
  const uint32_t b_off = uint32_t((((lane & 7) + ((lane >> 3) & 1) * 8) * kWgPitch + 16 * nb + (lane >> 4) * 8) * 2);
  const int strips = W / kWgPx;
  const int bands = B * H * strips;
  // band-independent half of the halo-load addressing (the same 16-byte chunks every band): shared-memory
  // destination, source offset relative to the band origin and which image borders the chunk depends on.  The loads
  // are issued by the warps that issue the MMAs, so every instruction spent here is taken from them.
  constexpr int kXChunks = 3 * (kWgPx + 2) * 8, kXIter = (kXChunks + 255) / 256;
  int x_rel[kXIter];
  uint32_t x_dst[kXIter], x_border = 0;
#pragma unroll
  for (int k = 0; k < kXIter; ++k) {
    const int i = tid + 256 * k;
    const int chunk = i & 7, col = (i >> 3) % (kWgPx + 2), r = i / (8 * (kWgPx + 2));
    x_rel[k] = ((r - 1) * W + (col - 1)) * kC + chunk * 8;
    x_dst[k] = uint32_t(((r * (kWgPx + 2) + col) * kWgPitch + chunk * 8) * 2);
    x_border |= (uint32_t(r == 0) | uint32_t(r == 2) << 1 | uint32_t(col == 0) << 2 | uint32_t(col == kWgPx + 1) << 3)
                << (4 * k);
  }
  const uint32_t y_dst0 = uint32_t(((tid >> 3) * kWgPitch + (tid & 7) * 8) * 2);
  auto prefetch = [&](int band, int buf) {
    const int sidx = band % strips;
    const int y = (band / strips) % H;
    const int b = band / (strips * H);
    const int x0 = sidx * kWgPx;
    const uint32_t sYb = smem_u32 + uint32_t(buf * kWgStageElems * 2);
    const uint32_t sXb = sYb + uint32_t(kWgPx * kWgPitch * 2);
    const size_t origin = ((size_t(b) * H + y) * W + x0) * kC;
    const bf16* src = dY + origin + tid * 8;
    cp_async_16(sYb + y_dst0, src, 16);
    cp_async_16(sYb + y_dst0 + uint32_t(32 * kWgPitch * 2), src + 256 * 8, 16);
    const uint32_t missing = (uint32_t(y == 0) | uint32_t(y == H - 1) << 1 | uint32_t(x0 == 0) << 2 |
                              uint32_t(x0 + kWgPx == W) << 3) * 0x11111111u;
    const uint32_t bad = x_border & missing;
    const bf16* xo = X + origin;
#pragma unroll
    for (int k = 0; k < kXIter; ++k) {
      if (k < kXIter - 1 || tid + 256 * k < kXChunks) {
        const bool ok = ((bad >> (4 * k)) & 0xFu) == 0u;
        cp_async_16(sXb + x_dst[k], ok ? xo + x_rel[k] : X, ok ? 16 : 0);
      }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
#ifdef FEN_EXP_WG_NOLOAD    // developer experiment: compute on whatever shared memory holds
#define prefetch(a, b) ((void)0)
#endif
  int buf = 0;
  if (int(blockIdx.x) < bands) prefetch(blockIdx.x, 0);
  for (int band = blockIdx.x; band < bands; band += gridDim.x, buf ^= 1) {
    const int next = band + gridDim.x;
    if (next < bands) {
      prefetch(next, buf ^ 1);
      asm volatile("cp.async.wait_group 1;" ::: "memory");
    } else {
      asm volatile("cp.async.wait_group 0;" ::: "memory");
    }
    __syncthreads();
    const uint32_t sY_u32 = smem_u32 + uint32_t(buf * kWgStageElems * 2);
    const uint32_t sX_u32 = sY_u32 + uint32_t(kWgPx * kWgPitch * 2);
    if (tid < kC) {
      const bf16* sY = reinterpret_cast<const bf16*>(wg_smem) + buf * kWgStageElems;
      float s = 0.f;
#pragma unroll 8
      for (int px = 0; px < kWgPx; ++px) s += __bfloat162float(sY[px * kWgPitch + tid]);
      bsum += s;
    }
#pragma unroll kWgUnroll
    for (int ks = 0; ks < kWgPx / 16; ++ks) {
      uint32_t a0[4], a1[4];
      ldmatrix_x4_trans(sY_u32 + uint32_t(ks * 16 * kWgPitch * 2) + a_off, a0);        // co rows 0..15 of the warp's 32
      ldmatrix_x4_trans(sY_u32 + uint32_t(ks * 16 * kWgPitch * 2) + a_off + 32, a1);   // co rows 16..31
#pragma unroll
      for (int r = 0; r < 3; ++r)
#pragma unroll
        for (int kx = 0; kx < 3; ++kx) {
          uint32_t bq[4];
          ldmatrix_x4_trans(sX_u32 + uint32_t(((r * (kWgPx + 2) + ks * 16 + kx) * kWgPitch) * 2) + b_off, bq);
          mma_bf16_16816(acc[r * 3 + kx][0], a0, bq[0], bq[1]);
          mma_bf16_16816(acc[r * 3 + kx][1], a0, bq[2], bq[3]);
          mma_bf16_16816(acc[r * 3 + kx][2], a1, bq[0], bq[1]);
          mma_bf16_16816(acc[r * 3 + kx][3], a1, bq[2], bq[3]);
        }
    }
    __syncthreads();   // this buffer is refilled by the prefetch of the next iteration
  }
#ifdef FEN_EXP_WG_NOFLUSH   // developer experiment: time the kernel without its output phase
  if (bands >= 0) { if (acc[0][0][0] == 123.456f) dW[0] = bsum; return; }
#endif
  // ---- flush: 4 passes of 16 output rows (co) through shared memory, then 128-bit reductions
  float* stage = reinterpret_cast<float*>(wg_smem);
  const int g = lane >> 2, t4 = lane & 3;
#pragma unroll
  for (int pass = 0; pass < 4; ++pass) {   // output rows co = 16 pass .. 16 pass + 15: m16 tile (pass & 1) of warps mb = pass >> 1
    if (mb == (pass >> 1)) {
#pragma unroll
      for (int j = 0; j < 2; ++j)
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          float* dst = stage + (g + (c >> 1) * 8) * kWgOutPitch + (16 * nb + 8 * j + 2 * t4 + (c & 1)) * 9;
#pragma unroll
          for (int t = 0; t < 9; ++t) dst[t] = acc[t][2 * (pass & 1) + j][c];
        }
    }
    __syncthreads();
    for (int i = tid; i < 16 * (kC * 9 / 4); i += 256) {
      const int row = i / (kC * 9 / 4), q = i % (kC * 9 / 4);
      const float4 v = *reinterpret_cast<const float4*>(stage + row * kWgOutPitch + 4 * q);
      const int co = (16 * pass + row) * co_mul + co_off;
      red_add_v4(dW + size_t(co) * (kC * 9) + 4 * q, v);
    }
    __syncthreads();
  }
  if (tid < kC) atomicAdd(dB + tid * co_mul + co_off, bsum);
}

#endif  // FEN_DEV
// --------------------------------------------------------------------------------------------------------------
// PReLU backward on NHWC bf16 (RCAB conv1, blocks.py:139-141; upsample stage, blocks.py:227):
//   out = g * (pre > 0 ? 1 : slope[c])      dslope[c] += g * min(pre, 0),  min(pre, 0) = act / slope[c] on that side
// act = prelu(pre) is the saved activation, `mask` the sign bits of pre (ConvParams::mask_out of the forward).  unshuffle != 0: g / act are the PixelShuffle'd [B,H,W,64] tensors and
// `out` is written as the four sub-pixel planes [4][B][H/2][W/2][64] of the producing convolution.
__global__ void __launch_bounds__(256)
prelu_bwd_kernel(const bf16* g, const bf16* __restrict__ act, const uint32_t* __restrict__ mask,
                 const float* __restrict__ slope, bf16* out, long long* __restrict__ dslope, int B, int H, int W,
                 int unshuffle) {   // out may alias g (in place); dslope: fixed point (gs_add)
  __shared__ unsigned long long s_ds[kC];     // fixed point as well: the order in which the threads of a block arrive must not matter
  if (threadIdx.x < kC) s_ds[threadIdx.x] = 0ull;
  __syncthreads();
  const int cg = threadIdx.x & 7;
  float sl[8], ds[8];
#pragma unroll
  for (int c = 0; c < 8; ++c) { sl[c] = __ldg(slope + cg * 8 + c); ds[c] = 0.f; }
  const size_t total = size_t(B) * H * W * 8;
  const uint4* gv = reinterpret_cast<const uint4*>(g);
  const uint4* av = reinterpret_cast<const uint4*>(act);
  uint4* ov = reinterpret_cast<uint4*>(out);
  for (size_t i = size_t(blockIdx.x) * blockDim.x + threadIdx.x; i < total; i += size_t(gridDim.x) * blockDim.x) {
    float gf[8], af[8], of[8];
    unpack8(gv[i], gf);
    unpack8(__ldg(av + i), af);
    const uint32_t pos_bits = __ldg(mask + (i >> 3) * 2 + (cg >> 2)) >> ((cg & 3) * 8);
#pragma unroll
    for (int c = 0; c < 8; ++c) {
      const bool pos = (pos_bits >> c) & 1u;
      of[c] = pos ? gf[c] : gf[c] * sl[c];
      ds[c] += (pos || sl[c] == 0.f) ? 0.f : gf[c] * (af[c] / sl[c]);
    }
    size_t o = i;
    if (unshuffle) {
      const size_t pix = i >> 3;
      const int x = int(pix % W), y = int((pix / W) % H), n = int(pix / (size_t(W) * H));
      const int sub = 2 * (y & 1) + (x & 1);
      o = ((((size_t(sub) * B + n) * (H >> 1) + (y >> 1)) * (W >> 1) + (x >> 1)) << 3) + cg;
    }
    ov[o] = pack8(of);
  }
#pragma unroll
  for (int c = 0; c < 8; ++c) atomicAdd(&s_ds[cg * 8 + c], static_cast<unsigned long long>(__float2ll_rn(ds[c] * kGsScale)));
  __syncthreads();
  if (threadIdx.x < kC) atomicAdd(reinterpret_cast<unsigned long long*>(dslope) + threadIdx.x, s_ds[threadIdx.x]);
}

// --------------------------------------------------------------------------------------------------------------
// Squeeze-and-excitation backward (blocks.py:86-92,153), x' = x + rs * s * o, s = sigmoid(W2 relu(W0 mean(o))).
// Pass 1, dsum[b][c] = sum_px dx'[b,px,c] * o[b,px,c], is done by the epilogue of the data-gradient convolution that
// produces dx' (kEpiDot in conv3x3_umma.cuh).
// Pass 2: the tiny FC chain backward per image (recomputed by every CTA of the image), parameter gradients of the
// two Linear layers (CTA 0 of the image), and  dO = rs * s[c] * dx' + dy[c] / HW.
//   y = sums / HW; z = relu(W0 y); t = W2 z; s = sigmoid(t)
//   ds = rs * dsum; dt = ds s (1 - s); dW2 += dt z^T; dz = W2^T dt; dzr = dz [z > 0]; dW0 += dzr y^T; dy = W0^T dzr
__global__ void __launch_bounds__(256)
se_bwd_apply_kernel(const bf16* __restrict__ dxo, const long long* __restrict__ sums, const long long* __restrict__ dsum,
                    const float* __restrict__ fc0, const float* __restrict__ fc2, int R, float inv_hw, float res_scale,
                    bf16* __restrict__ dO, long long* __restrict__ dfc0, long long* __restrict__ dfc2, int hw) {
  // The FC chain (4 dependent mat-vecs of 64 x R) runs on all 256 threads - 4 threads per output, both matrices staged
  // in shared memory by one coalesced pass: as a serial loop per output it was most of this kernel's 15 us.
  extern __shared__ float s_fc[];                      // fc0 [R][65] | fc2 [64][R + 1] (padded rows)
  float* s_fc0 = s_fc;
  float* s_fc2 = s_fc + R * 65;
  __shared__ float s_y[kC], s_z[kC], s_dt[kC], s_dzr[kC], s_mul[kC], s_add[kC];
  const int n = blockIdx.y, tid = threadIdx.x;
  for (int i = tid; i < R * kC; i += blockDim.x) {
    s_fc0[(i / kC) * 65 + (i % kC)] = __ldg(fc0 + i);
    s_fc2[(i / R) * (R + 1) + (i % R)] = __ldg(fc2 + i);
  }
  if (tid < kC) s_y[tid] = hs_to_float(sums[size_t(n) * kC + tid]) * inv_hw;
  // (everything above is forward state; dsum and dxo come from the launch just before this one)
  if (tid == 0) pdl_launch_dependents();
  pdl_wait();
  const float ds_n = (tid < 4 * kC) ? gs_to_float(__ldcg(dsum + size_t(n) * kC + (tid >> 2))) : 0.f;   // (fixed point: kEpiDot)
  __syncthreads();
  const int o = tid >> 2, part = tid & 3;              // output index, quarter of the dot product
  auto quad_sum = [](float a) {
    a += __shfl_xor_sync(0xffffffffu, a, 1);
    return a + __shfl_xor_sync(0xffffffffu, a, 2);
  };
  {                                                    // z = relu(fc0 y)
    float a = 0.f;
    if (o < R)
#pragma unroll
      for (int k = 0; k < 16; ++k) a = fmaf(s_fc0[o * 65 + part * 16 + k], s_y[part * 16 + k], a);
    a = quad_sum(a);
    if (o < R && part == 0) s_z[o] = fmaxf(a, 0.f);
  }
  __syncthreads();
  {                                                    // s = sigmoid(fc2 z); dt = d loss / d (fc2 z)
    float a = 0.f;
    for (int j = part; j < R; j += 4) a = fmaf(s_fc2[o * (R + 1) + j], s_z[j], a);
    a = quad_sum(a);
    if (part == 0) {
      const float sg = 1.f / (1.f + expf(-a));
      s_mul[o] = sg * res_scale;
      s_dt[o] = res_scale * ds_n * sg * (1.f - sg);
    }
  }
  __syncthreads();
  {                                                    // dz = fc2^T dt, gated by the ReLU
    float a = 0.f;
    if (o < R)
#pragma unroll
      for (int k = 0; k < 16; ++k) a = fmaf(s_fc2[(part * 16 + k) * (R + 1) + o], s_dt[part * 16 + k], a);
    a = quad_sum(a);
    if (o < R && part == 0) s_dzr[o] = s_z[o] > 0.f ? a : 0.f;
  }
  __syncthreads();
  {                                                    // d mean = fc0^T dz; every pixel gets 1 / HW of it
    float a = 0.f;
    for (int j = part; j < R; j += 4) a = fmaf(s_fc0[j * 65 + o], s_dzr[j], a);
    a = quad_sum(a);
    if (part == 0) s_add[o] = a * inv_hw;
  }
  if (blockIdx.x == 0) {
    for (int i = tid; i < kC * R; i += blockDim.x) {
      gs_add(dfc2 + i, s_dt[i / R] * s_z[i % R]);       // fc2 [64][R]
      gs_add(dfc0 + i, s_dzr[i / kC] * s_y[i % kC]);    // fc0 [R][64]
    }
  }
  __syncthreads();
  const int cg = tid & 7;
  float mul[8], add[8];
#pragma unroll
  for (int c = 0; c < 8; ++c) { mul[c] = s_mul[cg * 8 + c]; add[c] = s_add[cg * 8 + c]; }
  const size_t base = size_t(n) * hw * 8;
  const int total = hw * 8;
  const uint4* gv = reinterpret_cast<const uint4*>(dxo) + base;
  uint4* ov = reinterpret_cast<uint4*>(dO) + base;
  for (int i = blockIdx.x * blockDim.x + tid; i < total; i += gridDim.x * blockDim.x) {
    float gf[8], of[8];
    unpack8(__ldcg(gv + i), gf);
#pragma unroll
    for (int c = 0; c < 8; ++c) of[c] = fmaf(gf[c], mul[c], add[c]);
    ov[i] = pack8(of);
  }
}

// out = a + b (skip connections meeting in the backward pass), NHWC bf16, n8 = number of 8-element groups.
__global__ void __launch_bounds__(256)
add_bf16_kernel(const bf16* a, const bf16* b, bf16* out, size_t n8) {   // out may alias a or b
  const uint4* av = reinterpret_cast<const uint4*>(a);
  const uint4* bv = reinterpret_cast<const uint4*>(b);
  uint4* ov = reinterpret_cast<uint4*>(out);
  if (threadIdx.x == 0) pdl_launch_dependents();
  pdl_wait();
  for (size_t i = size_t(blockIdx.x) * blockDim.x + threadIdx.x; i < n8; i += size_t(gridDim.x) * blockDim.x) {
    float x[8], y[8];
    unpack8(__ldcg(av + i), x);
    unpack8(__ldcg(bv + i), y);
#pragma unroll
    for (int c = 0; c < 8; ++c) x[c] += y[c];
    ov[i] = pack8(x);
  }
}

// Transposed + tap-flipped weights for the data-gradient convolutions:
//   dst[grp][t][r = ci][col = c] = W[co = (groups == 4 ? 4 c + grp : c)][ci][8 - t]        (fp32 OIHW -> bf16)
// so that conv(dY_grp, dst_grp) = dX.  For the PixelShuffle convolutions the four sub-pixel planes are four
// 64 -> 64 convolutions whose results add up.
__global__ void pack_conv_T_kernel(const float* __restrict__ w, bf16* __restrict__ dst, int groups) {
  const int total = groups * 9 * kC * kC;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int c = i % kC;
    const int ci = (i / kC) % kC;
    const int t = (i / (kC * kC)) % 9;
    const int grp = i / (kC * kC * 9);
    const int co = groups == 4 ? 4 * c + grp : c;
    dst[i] = __float2bfloat16(w[(size_t(co) * kC + ci) * 9 + (8 - t)]);
  }
}
// conv_last weights OIHW [3][64][3][3] as the transposed, tap-flipped 64 -> 64 convolution of its data gradient:
//   dst[t][c][co] = W[co][c][8 - t] for co < 3, 0 for the padding rows (the input there is the 3-channel d out,
//   zero-extended to 64 channels by the tensor map)
__global__ void pack_last_T_kernel(const float* __restrict__ w, bf16* __restrict__ dst) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= 9 * kC * kC) return;
  const int co = i % kC, c = (i / kC) % kC, t = i / (kC * kC);
  dst[i] = __float2bfloat16(co < 3 ? w[(size_t(co) * kC + c) * 9 + (8 - t)] : 0.f);
}

// fp32 NCHW [B][3][H][W] -> bf16 NHWC [B][H][W][8] (channels 3 .. 7 zero): the 3-channel tensors of the backward pass
// (d out at the network's end, the LR input at its head) in the form the 64-channel tcgen05 kernels read through a
// narrow tensor map.  One pixel per thread: three coalesced plane reads, one 16-byte store.
__global__ void __launch_bounds__(256)
nchw3_to_nhwc8_kernel(const float* __restrict__ src, bf16* __restrict__ dst, int B, size_t hw) {
  const size_t total = size_t(B) * hw;
  for (size_t i = size_t(blockIdx.x) * blockDim.x + threadIdx.x; i < total; i += size_t(gridDim.x) * blockDim.x) {
    const size_t b = i / hw, px = i - b * hw;
    const float* s0 = src + b * 3 * hw + px;
    const float t8[8] = {__ldg(s0), __ldg(s0 + hw), __ldg(s0 + 2 * hw), 0.f, 0.f, 0.f, 0.f, 0.f};
    reinterpret_cast<uint4*>(dst)[i] = pack8(t8);
  }
}

// Fixed-point shadow -> fp32 gradients of the small tensors.  g64 has the layout of the flat gradient vector; this
// kernel converts, per RCAB of one group (blockIdx.y = block), the PReLU slope [64] and the two SE matrices [2 R 64], or
// (n_rcab == 0) one plain range.
__global__ void __launch_bounds__(256)
grads_from_fixed_kernel(const long long* __restrict__ g64, float* __restrict__ grads, int64_t base, int64_t stride,
                        int64_t off_a, int cnt_a, int64_t off_b, int cnt_b) {
  const int64_t r0 = base + int64_t(blockIdx.y) * stride;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < cnt_a + cnt_b; i += gridDim.x * blockDim.x) {
    const int64_t k = r0 + (i < cnt_a ? off_a + i : off_b + (i - cnt_a));
    grads[k] = gs_to_float(g64[k]);
  }
}

// Sum of the per-CTA partial weight gradients of one convolution, in a FIXED order (wgrad_*_umma_kernel write their
// [64][576] + [64] partials with plain stores instead of racing fp32 atomics): bit-identical gradients run to run.
constexpr int kWgPartFloats = kC * kC * 9 + kC;
__device__ __forceinline__ void wgrad_reduce_one(const float* __restrict__ parts, int n_parts, float* __restrict__ dW,
                                                 float* __restrict__ dB, int co_mul, int co_off, int first, int step) {
  constexpr int kRow4 = kC * 9 / 4;                  // float4 per output row
  for (int e = first; e < kC * kRow4 + kC / 4; e += step) {
    float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
    const float4* src = reinterpret_cast<const float4*>(parts) + e;
    for (int k0 = 0; k0 < n_parts; k0 += 8) {        // eight loads in flight, added in index order
      float4 v[8];
#pragma unroll
      for (int j = 0; j < 8; ++j)
        v[j] = (k0 + j < n_parts) ? __ldcg(src + size_t(k0 + j) * (kWgPartFloats / 4)) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
      for (int j = 0; j < 8; ++j) { a.x += v[j].x; a.y += v[j].y; a.z += v[j].z; a.w += v[j].w; }
    }
    if (e < kC * kRow4) {
      const int co = e / kRow4, q = e - co * kRow4;
      *reinterpret_cast<float4*>(dW + size_t(co * co_mul + co_off) * (kC * 9) + 4 * q) = a;
    } else {
      const int c4 = 4 * (e - kC * kRow4);
      dB[(c4 + 0) * co_mul + co_off] = a.x; dB[(c4 + 1) * co_mul + co_off] = a.y;
      dB[(c4 + 2) * co_mul + co_off] = a.z; dB[(c4 + 3) * co_mul + co_off] = a.w;
    }
  }
}
__global__ void __launch_bounds__(256)
wgrad_reduce_kernel(const float* __restrict__ parts, int n_parts, float* __restrict__ dW, float* __restrict__ dB,
                    int co_mul, int co_off) {
  wgrad_reduce_one(parts, n_parts, dW, dB, co_mul, co_off, blockIdx.x * blockDim.x + threadIdx.x, gridDim.x * blockDim.x);
}

}  // namespace fen
