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
                      float* __restrict__ parts, int B, int H, int W) {
  // parts: [gridDim.x][64 x 576 + 64] fp32 - every CTA leaves its partial weight / bias gradient there with plain
  // stores; wgrad_reduce_kernel adds them up in CTA order (deterministic; the racing fp32 atomics this replaces gave
  // run-to-run differences in the last bits).
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
  float* part = parts + size_t(blockIdx.x) * kWgPartFloats;
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
          for (int j = 0; j < 16; ++j) part[kC * kC * 9 + 16 * pass + j] = __uint_as_float(v[j]);
        }
      }
    }
    __syncthreads();
    for (int i = tid; i < 16 * (kC * 9 / 4); i += kWuThreads) {
      const int row = i / (kC * 9 / 4), qq = i % (kC * 9 / 4);
      const float4 v4 = *reinterpret_cast<const float4*>(stage + row * kWgOutPitch + 4 * qq);
      *reinterpret_cast<float4*>(part + size_t(16 * pass + row) * (kC * 9) + 4 * qq) = v4;
    }
    __syncthreads();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, kWuTmemCols);
}


// ====================================================================================================================
// Batched variant: the weight gradients of ALL 64 -> 64 body convolutions of (a range of) the backward pass in ONE
// persistent launch.  One launch per convolution (the kernel above) spends most of its 25 us at batch 32 on fixed costs
// - TMEM allocation, the first TMA round trip, a 147 KB reduction flush by each of 148 CTAs - for 9.7 us of MMA work.
// Here the backward keeps every layer's output gradient (fen_step_host.cuh: one buffer per convolution), and after a
// run of groups this kernel sweeps over work items (job, chunk of bands): TMEM, barriers and pipeline are set up once, a
// job is split over `chunks` CTAs only (16: 2.4 MB of reductions per job instead of 21.8 MB), and TWO issuer warps feed
// the tensor pipe (one thread sustains one tcgen05.mma per ~81 cycles, the pipe takes one per ~48: issuer 0 owns the
// accumulators of tap pairs 0-2, issuer 1 those of pair 3 and of tap 8 / bias: 12 + 8 instructions per band).
// All activation-sized buffers of the step workspace lie at one stride: two 5-D tensor maps (dY box: one 64-pixel row;
// X box: three 66-pixel rows) address any of them by index.  Jobs are decoded from the layer structure (no table):
//   job 0: conv_after_body;  then per group g = G-1 .. 0: its group conv, then per RCAB b = Bk-1 .. 0: conv2, conv1.
struct WgBatchParams {
  int B, H, W;
  int G, Bk;
  int job_begin, job_end;           // jobs of this launch (backward order, see above)
  int chunks;                       // CTAs-worth of band ranges per job
  // buffer indices (units of the workspace's activation stride)
  int i_xs0, i_h0, i_gout0, i_dO0, i_dH0, i_dG0, i_dBody;     // f0 is buffer 0
  // flat-gradient offsets (floats)
  int64_t p_rcab0, p_rcab_stride, p_group_stride, p_gconv_w_in_group, p_after_w;
  float* grads;
  float* parts;                     // [n_items][64 x 576 + 64] fp32 partial gradients, item = (job - job_begin) * chunks + chunk
};
struct WgJob { int y_buf, x_buf; float* dW; float* dB; };
__device__ __forceinline__ WgJob wg_job(const WgBatchParams& p, int j) {
  constexpr int64_t kW = 64 * 64 * 9;
  WgJob o;
  if (j == 0) { o.y_buf = p.i_dBody; o.x_buf = p.i_gout0 + p.G - 1; o.dW = p.grads + p.p_after_w; o.dB = o.dW + kW; return o; }
  const int per = 1 + 2 * p.Bk;
  const int gi = (j - 1) / per, k = (j - 1) - gi * per, g = p.G - 1 - gi;
  float* pg = p.grads + p.p_rcab0 + int64_t(g) * p.p_group_stride;
  if (k == 0) {
    o.y_buf = p.i_dG0 + g; o.x_buf = p.i_xs0 + g * p.Bk + p.Bk - 1; o.dW = pg + p.p_gconv_w_in_group; o.dB = o.dW + kW;
    return o;
  }
  const int b = p.Bk - 1 - (k - 1) / 2, r = g * p.Bk + b;
  float* pr = pg + int64_t(b) * p.p_rcab_stride;            // conv1.w conv1.b prelu conv2.w conv2.b fc0 fc2
  if (((k - 1) & 1) == 0) { o.y_buf = p.i_dO0 + r; o.x_buf = p.i_h0 + r; o.dW = pr + kW + 64 + 64; o.dB = o.dW + kW; }
  else {
    o.y_buf = p.i_dH0 + r;
    o.x_buf = (b == 0) ? (g == 0 ? 0 : p.i_gout0 + g - 1) : p.i_xs0 + r - 1;
    o.dW = pr; o.dB = pr + kW;
  }
  return o;
}

__device__ __forceinline__ void tma_load_5d(const CUtensorMap* m, uint64_t* bar, uint32_t dst_smem, int c0, int c1, int c2,
                                            int c3, int c4) {
  asm volatile(
      "cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
      :
      : "r"(dst_smem), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
      : "memory");
}

constexpr int kWbStages = 4;
constexpr int kWbFlushBytes = 16 * kWgOutPitch * 4;               // 37 120: dedicated flush staging (16 output rows)
constexpr int kWbDynBytes = kWbStages * kWuStageBytes + kWuOnesBytes + ((kWbFlushBytes + 1023) / 1024) * 1024 + 1024;
constexpr int kWbThreads = 224;                                    // warp 0 TMA, warps 1-2 MMA issuers, warps 3-6 TMEM readers
static_assert(kWbDynBytes <= 227 * 1024, "batched weight gradient: shared memory");

__global__ void __launch_bounds__(kWbThreads, 1)
wgrad_batch_umma_kernel(const __grid_constant__ CUtensorMap tm_y, const __grid_constant__ CUtensorMap tm_x,
                        const WgBatchParams p) {
  extern __shared__ uint8_t wb_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(wb_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* ones = smem + kWbStages * kWuStageBytes;
  float* stage = reinterpret_cast<float*>(ones + kWuOnesBytes);
  __shared__ uint64_t bar_full[kWbStages], bar_empty[kWbStages], bar_done;
  __shared__ uint32_t tmem_slot;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int strips = (p.W + kStripW - 1) / kStripW;
  const int bands = p.B * p.H * strips;
  const int per_chunk = (bands + p.chunks - 1) / p.chunks;
  const int n_items = (p.job_end - p.job_begin) * p.chunks;

  if (warp == 1) tmem_alloc(&tmem_slot, kWuTmemCols);
  if (tid == 0) {
    for (int i = 0; i < kWbStages; ++i) { mbar_init(&bar_full[i], 1); mbar_init(&bar_empty[i], 2); }
    mbar_init(&bar_done, 2);
    fence_mbar_init();
    tma_prefetch_desc(&tm_y);
    tma_prefetch_desc(&tm_x);
  }
  for (int i = tid; i < kWuOnesBytes / 4; i += kWbThreads) reinterpret_cast<uint32_t*>(ones)[i] = 0x3F803F80u;
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_slot;
  const uint32_t smem_base = smem_u32(smem);

  uint32_t it = 0;          // running band counter of this CTA (ring position), identical in every role
  uint32_t n_done = 0;      // items finished (phase of bar_done)
  for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++n_done) {
    const WgJob job = wg_job(p, p.job_begin + item / p.chunks);
    const int chunk = item % p.chunks;
    const int band0 = chunk * per_chunk, band1 = min(bands, band0 + per_chunk);
    const int nb = max(0, band1 - band0);
    float* part = p.parts + size_t(item) * kWgPartFloats;      // this item's partial (plain stores: see wgrad_reduce_kernel)
    if (nb == 0)                                               // (an empty band range still owes its - zero - partial)
      for (int i = tid; i < kWgPartFloats; i += kWbThreads) part[i] = 0.f;
    if (warp == 0) {
      // ============================================================ TMA producer
      for (int k = 0; k < nb; ++k) {
        const uint32_t slot = (it + k) % kWbStages, ph = ((it + k) / kWbStages) & 1;
        mbar_wait(&bar_empty[slot], ph ^ 1);
        __syncwarp();
        if (lane == 0) {
          const int band = band0 + k;
          const int sidx = band % strips, y = (band / strips) % p.H, b = band / (strips * p.H);
          const int x0 = sidx * kStripW;
          const uint32_t dst = smem_base + slot * kWuStageBytes;
          mbar_expect_tx(&bar_full[slot], kWuYBytes + kWuXBytes);
          tma_load_5d(&tm_y, &bar_full[slot], dst, 0, x0, y, b, job.y_buf);
          tma_load_5d(&tm_x, &bar_full[slot], dst + kWuYBytes, 0, x0 - 1, y - 1, b, job.x_buf);
        }
        __syncwarp();
      }
    } else if (warp <= 2) {
      // ============================================================ MMA issuers (warp 1: tap pairs 0-2, warp 2: pair 3 + tap 8 / bias)
      constexpr uint32_t idesc = umma_idesc_bf16(128, kC) | (1u << 15) | (1u << 16);   // A and B MN-major
      const uint32_t ones_u32 = smem_u32(ones);
      const int g_lo = (warp == 1) ? 0 : 3, g_hi = (warp == 1) ? 3 : 5;
      for (int k = 0; k < nb; ++k) {
        const uint32_t slot = (it + k) % kWbStages, ph = ((it + k) / kWbStages) & 1;
        mbar_wait(&bar_full[slot], ph);
        __syncwarp();
        tc_fence_after();
        if (lane == 0) {
          const uint32_t sY = smem_base + slot * kWuStageBytes, sX = sY + kWuYBytes;
#pragma unroll
          for (int g = 0; g < 5; ++g) {
            if (g < g_lo || g >= g_hi) continue;
            const int t0 = 2 * g, t1 = 2 * g + 1;
            const uint32_t a0 = sX + uint32_t(((t0 / 3) * kPitch + (t0 % 3)) * 128);
            const uint32_t a1 = sX + uint32_t(((t1 / 3) * kPitch + (t1 % 3)) * 128);   // g == 4: replaced by the ones tile
#pragma unroll
            for (int ks = 0; ks < kStripW / 16; ++ks) {
              const uint32_t a_start = a0 + ks * 2048;
              const uint32_t lbo = (g < 4) ? (a1 - a0) : (ones_u32 - a_start);
              const uint64_t adesc = umma_smem_desc(a_start, lbo, 1024, UMMA_LAYOUT_SW128);
              const uint64_t bdesc = umma_smem_desc(sY + ks * 2048, 128, 1024, UMMA_LAYOUT_SW128);
              umma_bf16_ss(tmem_base + g * kC, adesc, bdesc, idesc, (k | ks) != 0 ? 1u : 0u);
            }
          }
          umma_commit(&bar_empty[slot]);   // (count 2: the stage is free once both issuers' MMAs have read it)
        }
        __syncwarp();
      }
      if (lane == 0) umma_commit(&bar_done);
      __syncwarp();
    }
    it += uint32_t(nb);
    // ---- every MMA of the item has completed: flush the five accumulators (4 passes of 16 output rows)
    mbar_wait(&bar_done, n_done & 1);
    __syncwarp();
    tc_fence_after();
    const int q = warp & 3;                       // TMEM lane quarter this warp may read
    const int tsel = q >> 1, ci = (32 * q + lane) & 63;
    for (int pass = 0; pass < 4; ++pass) {
      if (warp >= 3 && nb > 0) {
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
            for (int j = 0; j < 16; ++j) part[kC * kC * 9 + 16 * pass + j] = __uint_as_float(v[j]);
          }
        }
      }
      tc_fence_before();
      __syncthreads();
      if (nb > 0) {
        for (int i = tid; i < 16 * (kC * 9 / 4); i += kWbThreads) {
          const int row = i / (kC * 9 / 4), qq = i % (kC * 9 / 4);
          const float4 v4 = *reinterpret_cast<const float4*>(stage + row * kWgOutPitch + 4 * qq);
          *reinterpret_cast<float4*>(part + size_t(16 * pass + row) * (kC * 9) + 4 * qq) = v4;
        }
      }
      __syncthreads();
    }
    tc_fence_after();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, kWuTmemCols);
}

// The partials of a batched launch -> the flat gradient: blockIdx.y = job, the chunks summed in order.
__global__ void __launch_bounds__(256)
wgrad_reduce_batch_kernel(const WgBatchParams p) {
  const WgJob job = wg_job(p, p.job_begin + blockIdx.y);
  wgrad_reduce_one(p.parts + size_t(blockIdx.y) * p.chunks * kWgPartFloats, p.chunks, job.dW, job.dB, 1, 0,
                   blockIdx.x * blockDim.x + threadIdx.x, gridDim.x * blockDim.x);
}

}  // namespace fen
