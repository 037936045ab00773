// Weight gradient of a 64 -> 64 3x3 / pad-1 convolution on tcgen05 (third generation; fen_backward.cuh holds the
// fp32-FMA and the mma.sync ones):
//   dW[co][ci][ky][kx] += sum_{b,y,x} dY[b,y,x,co] * X[b,y+ky-1,x+kx-1,ci]        db[co] += sum dY[b,y,x,co]
// The contraction index is the PIXEL, and NHWC keeps a pixel's 64 channels contiguous (128 B): staged by TMA with
// SWIZZLE_128B, 8 pixels x 128 B is exactly the canonical MN-MAJOR swizzle atom of a tcgen05 shared-memory descriptor
// (8 rows of K, each 64 MN elements; SBO = 1024 B between groups of 8 pixels), so both operands go to the tensor core
// as they lie, with the transpose bits of the instruction descriptor set:
//   A (M = 128): TWO taps of X side by side - rows 0..63 = the 64 input channels at tap t0, rows 64..127 = at tap t1.
//                The second 64-row block is just the same buffer `LBO` bytes further (128 B for a neighbouring tap).
//   B (N = 64) : the dY row (64 output channels).            K = 16 pixels per instruction, 4 instructions per row.
// 9 taps = 4 tap pairs + tap 8, whose spare half multiplies a tile of ones: lanes 64..127 of the fifth accumulator
// hold the bias gradient.  Five 128 x 64 fp32 accumulators (320 TMEM columns) live for the whole kernel; a CTA walks
// over bands (one image row of one 64-column strip) and flushes once, through shared memory, with 128-bit reductions.
// Roles: warp 0 TMA producer, warp 1 MMA issuer (+ TMEM owner), warps 2-5 read TMEM; all six warps flush.
#pragma once
#include "conv3x3_umma.cuh"
#include "fen_backward.cuh"

namespace fen {

constexpr int kWuStages = 5;
constexpr int kWuYBytes = kStripW * kC * 2;                       // 8 192: dY row
constexpr int kWuXBytes = 3 * kPitch * kC * 2;                    // 25 344: 3 rows of X with halo columns
constexpr int kWuXPad = (kWuXBytes + 1023) / 1024 * 1024 + 1024;  // 26 624: the next stage starts on a swizzle-atom boundary
constexpr int kWuStageBytes = kWuYBytes + kWuXPad;                // multiple of 1 024
constexpr int kWuOnesBytes = 2048;                                // [16 px][64 ch] of 1.0
constexpr int kWuDynBytes = kWuStages * kWuStageBytes + kWuOnesBytes + 1024;
constexpr int kWuThreads = 192;
constexpr uint32_t kWuTmemCols = 512;                             // 5 x 64 used
static_assert(16 * kWgOutPitch * 4 <= kWuStages * kWuStageBytes, "output staging must fit the operand buffers");

__global__ void __launch_bounds__(kWuThreads, 1)
wgrad_c64_umma_kernel(const __grid_constant__ CUtensorMap tm_y, const __grid_constant__ CUtensorMap tm_x,
                      float* __restrict__ dW, float* __restrict__ dB, int B, int H, int W, int co_mul, int co_off) {
  extern __shared__ uint8_t wu_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(wu_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* ones = smem + kWuStages * kWuStageBytes;
  __shared__ uint64_t bar_full[kWuStages], bar_empty[kWuStages], bar_done;
  __shared__ uint32_t tmem_slot;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int strips = (W + kStripW - 1) / kStripW;   // a ragged last strip is zero-filled by TMA
  const int bands = B * H * strips;

  if (warp == 1) tmem_alloc(&tmem_slot, kWuTmemCols);
  if (tid == 0) {
    for (int i = 0; i < kWuStages; ++i) { mbar_init(&bar_full[i], 1); mbar_init(&bar_empty[i], 1); }
    mbar_init(&bar_done, 1);
    fence_mbar_init();
    tma_prefetch_desc(&tm_y);
    tma_prefetch_desc(&tm_x);
  }
  for (int i = tid; i < kWuOnesBytes / 4; i += kWuThreads) reinterpret_cast<uint32_t*>(ones)[i] = 0x3F803F80u;
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_slot;
  const uint32_t smem_base = smem_u32(smem);

  if (warp == 0) {
    // ============================================================ TMA producer
    int it = 0;
    for (int band = blockIdx.x; band < bands; band += gridDim.x, ++it) {
      const int slot = it % kWuStages, ph = (it / kWuStages) & 1;
      mbar_wait(&bar_empty[slot], ph ^ 1);
      __syncwarp();
      if (lane == 0) {
        const int sidx = band % strips, y = (band / strips) % H, b = band / (strips * H);
        const int x0 = sidx * kStripW;
        const uint32_t dst = smem_base + slot * kWuStageBytes;
        mbar_expect_tx(&bar_full[slot], kWuYBytes + kWuXBytes);
        tma_load_4d(&tm_y, &bar_full[slot], dst, 0, x0, y, b);
        tma_load_4d(&tm_x, &bar_full[slot], dst + kWuYBytes, 0, x0 - 1, y - 1, b);
      }
      __syncwarp();
    }
  } else if (warp == 1) {
    // ============================================================ MMA issuer
    constexpr uint32_t idesc = umma_idesc_bf16(128, kC) | (1u << 15) | (1u << 16);   // A and B MN-major
    const uint32_t ones_u32 = smem_u32(ones);
    int it = 0;
    for (int band = blockIdx.x; band < bands; band += gridDim.x, ++it) {
      const int slot = it % kWuStages, ph = (it / kWuStages) & 1;
      mbar_wait(&bar_full[slot], ph);
      __syncwarp();
      tc_fence_after();
      if (lane == 0) {
        const uint32_t sY = smem_base + slot * kWuStageBytes, sX = sY + kWuYBytes;
#pragma unroll
        for (int g = 0; g < 5; ++g) {
          const int t0 = 2 * g, t1 = 2 * g + 1;
          const uint32_t a0 = sX + uint32_t(((t0 / 3) * kPitch + (t0 % 3)) * 128);
          const uint32_t a1 = sX + uint32_t(((t1 / 3) * kPitch + (t1 % 3)) * 128);   // g == 4: replaced by the ones tile
#pragma unroll
          for (int ks = 0; ks < kStripW / 16; ++ks) {
            const uint32_t a_start = a0 + ks * 2048;
            const uint32_t lbo = (g < 4) ? (a1 - a0) : (ones_u32 - a_start);
            const uint64_t adesc = umma_smem_desc(a_start, lbo, 1024, UMMA_LAYOUT_SW128);
            const uint64_t bdesc = umma_smem_desc(sY + ks * 2048, 128, 1024, UMMA_LAYOUT_SW128);
            umma_bf16_ss(tmem_base + g * kC, adesc, bdesc, idesc, (it | ks) != 0 ? 1u : 0u);
          }
        }
        umma_commit(&bar_empty[slot]);   // the stage is free once these MMAs have read it
      }
      __syncwarp();
    }
    if (lane == 0) umma_commit(&bar_done);
    __syncwarp();
  } else {
    mbar_wait(&bar_done, 0);
    __syncwarp();
    tc_fence_after();
  }
  // ---- flush: 4 passes of 16 output rows (co) through shared memory, then 128-bit reductions by all warps.
  // TMEM lane = 64 * (second tap of the pair) + ci, column = 64 * pair + co.
  __syncthreads();   // every MMA has completed (the epilogue warps waited for bar_done): the stages are free
  float* stage = reinterpret_cast<float*>(smem);
  const int q = warp & 3;                       // TMEM lane quarter this warp may read
  const int tsel = q >> 1, ci = (32 * q + lane) & 63;
  for (int pass = 0; pass < 4; ++pass) {
    if (warp >= 2) {
#pragma unroll
      for (int g = 0; g < 5; ++g) {
        uint32_t v[16];
        tmem_ld_32x16(tmem_base + (uint32_t(32 * q) << 16) + uint32_t(g * kC + 16 * pass), v);
        tmem_ld_wait();
        const int tap = 2 * g + tsel;
        if (tap < 9) {
#pragma unroll
          for (int j = 0; j < 16; ++j) stage[j * kWgOutPitch + ci * 9 + tap] = __uint_as_float(v[j]);
        } else if (ci == 0) {   // lanes 64..127 of the fifth accumulator: the bias gradient (64 identical copies)
#pragma unroll
          for (int j = 0; j < 16; ++j) atomicAdd(dB + (16 * pass + j) * co_mul + co_off, __uint_as_float(v[j]));
        }
      }
    }
    __syncthreads();
    for (int i = tid; i < 16 * (kC * 9 / 4); i += kWuThreads) {
      const int row = i / (kC * 9 / 4), qq = i % (kC * 9 / 4);
      const float4 v4 = *reinterpret_cast<const float4*>(stage + row * kWgOutPitch + 4 * qq);
      const int co = (16 * pass + row) * co_mul + co_off;
      red_add_v4(dW + size_t(co) * (kC * 9) + 4 * qq, v4);
    }
    __syncthreads();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, kWuTmemCols);
}

}  // namespace fen
