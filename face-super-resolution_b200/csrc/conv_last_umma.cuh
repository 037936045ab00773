// conv_last (64 -> 3, 3x3, pad 1) + bias + bicubic x4 skip + clamp, with the nine taps in the N dimension of the MMA.
// Reference: src/models/custom.py:119-124,181-188 (conv_last, F.interpolate(..., 'bicubic'), torch.clamp).
//
// Why not the generic kernel (conv3x3_umma2_kernel<16>).  A tcgen05.mma of M = 128 costs ~48 cycles whatever N <= 64 is
// (its operands come from shared memory at 128 B / clock), so the generic formulation - 9 taps x 4 k-steps = 36 MMAs per
// 128 pixels for THREE output channels - is bound by MMA issue: 243 us at batch 64, 2.7 TB/s, 41 % of the HBM time of its
// 537 MB input.  Here one MMA per k-step computes all 27 (tap, output) products of a pixel at once,
//     D[s][3 t + co] = sum_ci u1[s][ci] * W[co][ci][t]        (N = 32, K = 64: FOUR MMAs per 128 staged pixels)
// for every STAGED pixel s (image + 1-pixel halo, zero-filled by TMA), and the convolution is the shifted sum
//     out(y, x)[co] = sum_t D[s(y + dy_t, x + dx_t)][3 t + co]
// taken by the epilogue from a shared-memory copy of D.  The kernel is then bound by reading its input once.
//
// Work item = 6 output rows of one 64-column strip of one image: an 8-row x 66-pixel TMA box (66 KB), 5 MMA tiles.
// warp 0: TMA producer (two box buffers: the input streams in continuously - with one buffer the loads of item i + 1
// waited for the MMAs of item i and the kernel ran at 40 % of its HBM time);
// warp 1: MMA issuer (two TMEM accumulator sets: the MMAs of item i + 1 run under the epilogue of item i);
// warps 2-17: epilogue - phase 1 TMEM -> shared D[528][27] fp32, phase 2 the shifted sums + bicubic skip + stores.
#pragma once
#include "conv3x3_umma.cuh"

namespace fen {

constexpr int kCLRows = 6;                                     // output rows per work item
constexpr int kCLBoxRows = kCLRows + 2;                        // staged rows (one halo row above and below)
constexpr int kCLPx = kCLBoxRows * kPitch;                     // 528 staged pixels
constexpr int kCLTiles = (kCLPx + kTileM - 1) / kTileM;        // 5 (the last tile reads 112 pixels past the box: rows nobody gathers)
constexpr int kCLN = 32;                                       // 27 = 9 taps x 3 outputs, padded to the MMA granule
constexpr int kCLBoxBytes = kCLPx * kC * 2;                    // 67 584 = 66 x 1024 (swizzle-atom aligned)
constexpr int kCLABytes = 2 * kCLBoxBytes;                     // two boxes: item i + 1 streams in while item i is multiplied
static_assert(kCLBoxBytes % 1024 == 0, "box buffers must keep the 1024-byte swizzle alignment");
constexpr int kCLWBytes = kCLN * kC * 2;                       // 4 096
constexpr int kCLDBytes = 9 * kCLPx * 16;                      // D as [tap][staged pixel][co padded to 4] fp32: one 128-bit access per tap, 76 032
constexpr int kCLDynBytes = kCLABytes + kCLWBytes + kCLDBytes + 1024;
constexpr int kCLEpiWarps = 16;                                // 4 per TMEM lane quarter: warp group g takes tiles g, g + 4
constexpr int kCLEpiThreads = 32 * kCLEpiWarps;
constexpr int kCLThreads = 32 * (2 + kCLEpiWarps);
constexpr int kCLSkipRows = 8;                                 // LR rows the 4-tap vertical filter of kCLRows output rows can touch (<= 7)
static_assert(kCLSkipRows * 3 * kStripW == 3 * kCLEpiThreads, "the horizontal skip pass is written for three sweeps");
constexpr uint32_t kCLSetCols = 256;                           // TMEM columns per accumulator set (5 x 32 used)

struct ConvLastParams {
  int B, H, W;            // size of the input u1 = size of the output
  int strips, nblk;       // ceil(W / 64), ceil(H / kCLRows)
  int items, items_per_cta;
  int training;           // no clamp when non-zero (custom.py:187)
  int bgr;                // out_u8 channel order B, G, R
  const float* bias;      // [3]
  const float* lr;        // [B][3][H/4][W/4] fp32 network input
  float* out_f32;         // [B][3][H][W] fp32, optional when out_u8 is given
  uint8_t* out_u8;        // [B][H][W][3] uint8 = trunc(clip(out * 255, 0, 255)), optional
};

__device__ __forceinline__ void cl_bar_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

__global__ void __launch_bounds__(kCLThreads, 1)
conv_last_umma_kernel(const __grid_constant__ CUtensorMap tm_in, const __grid_constant__ CUtensorMap tm_w,
                      const ConvLastParams p) {
  extern __shared__ uint8_t smem_raw[];
  // (aligned by OFFSET, not by rounding the pointer through an integer: the compiler must keep seeing a shared-memory
  // address, or every access to D below becomes a generic load / store)
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* a_smem = smem;                                      // 2 x [528 px][64 ch] bf16, SWIZZLE_128B
  uint8_t* w_smem = smem + kCLABytes;                          // [32][64] bf16, SWIZZLE_128B: row 3 t + co
  float4* d_smem = reinterpret_cast<float4*>(smem + kCLABytes + kCLWBytes);   // [9][528] x (3 outputs + pad) fp32
  __shared__ uint64_t bar_w, bar_a_full[2], bar_a_free[2], bar_acc_full[2], bar_acc_free[2];
  __shared__ uint32_t tmem_slot;
  __shared__ float s_bias[4];
  __shared__ float s_rf[kCLSkipRows][3][kStripW];              // bicubic skip, horizontally filtered: [LR row][channel][output column]

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int it0 = blockIdx.x * p.items_per_cta;
  const int it1 = min(p.items, it0 + p.items_per_cta);

  if (warp == 1) tmem_alloc(&tmem_slot, 2 * kCLSetCols);
  if (tid == 0) {
    mbar_init(&bar_w, 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&bar_a_full[i], 1); mbar_init(&bar_a_free[i], 1);
      mbar_init(&bar_acc_full[i], 1); mbar_init(&bar_acc_free[i], kCLEpiWarps);
    }
    fence_mbar_init();
    tma_prefetch_desc(&tm_in);
    tma_prefetch_desc(&tm_w);
  }
  if (tid < 3) s_bias[tid] = p.bias[tid];
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_slot;
  if (tid == 0) pdl_launch_dependents();

  auto decode = [&](int it, int& n, int& strip, int& blk) {
    blk = it % p.nblk;
    const int seg = it / p.nblk;
    strip = seg % p.strips;
    n = seg / p.strips;
  };

  auto next_item = [&](int& n, int& strip, int& blk) {        // (no divisions in the item loops)
    if (++blk == p.nblk) { blk = 0; if (++strip == p.strips) { strip = 0; ++n; } }
  };

  if (warp == 0) {
    // ============================================================ TMA producer
    // (whole-warp waits, lane 0 issues in straight-line blocks: see conv3x3_umma.cuh)
    if (it0 < it1) {
      if (lane == 0) {
        mbar_expect_tx(&bar_w, kCLWBytes);
        tma_load_2d(&tm_w, &bar_w, w_smem, 0, 0);                // (packed long before the predecessor started)
      }
      __syncwarp();
      pdl_wait();
      int n, strip, blk;
      decode(it0, n, strip, blk);
      for (int it = it0, k = 0; it < it1; ++it, ++k, next_item(n, strip, blk)) {
        const uint32_t slot = uint32_t(k) & 1u;
        if (k >= 2) mbar_wait(&bar_a_free[slot], ((uint32_t(k) >> 1) - 1u) & 1u);   // the MMAs of item k - 2 have read the buffer
        __syncwarp();
        if (lane == 0) {
          mbar_expect_tx(&bar_a_full[slot], kCLBoxBytes);
          tma_load_4d(&tm_in, &bar_a_full[slot], smem_u32(a_smem + slot * kCLBoxBytes), 0, strip * kStripW - 1, blk * kCLRows - 1, n);
        }
        __syncwarp();
      }
    }
  } else if (warp == 1) {
    // ============================================================ MMA issuer
    if (it0 < it1) {
      constexpr uint32_t idesc = umma_idesc_bf16(kTileM, kCLN);
      constexpr uint32_t kDescHi = (1024u >> 4) | (1u << 14) | (2u << 29);
      constexpr uint32_t kLbo = 1u << 16;
      const uint32_t a_lo0 = (smem_u32(a_smem) >> 4) | kLbo;
      const uint32_t b_lo0 = (smem_u32(w_smem) >> 4) | kLbo;
      const bool leader = elect_one();
      mbar_wait(&bar_w, 0);
      for (int it = it0, k = 0; it < it1; ++it, ++k) {
        const uint32_t sel = uint32_t(k) & 1u;
        mbar_wait(&bar_a_full[sel], (uint32_t(k) >> 1) & 1u);
        if (k >= 2) mbar_wait(&bar_acc_free[sel], ((uint32_t(k) >> 1) - 1u) & 1u);   // set last used by item k - 2
        __syncwarp();                              // converge after the spin-waits before any tcgen05 issue
        tc_fence_after();
        if (leader) {
#pragma unroll
          for (int t = 0; t < kCLTiles; ++t) {
#pragma unroll
            for (int kk = 0; kk < 4; ++kk)
              umma_bf16_ss_lohi(tmem_base + sel * kCLSetCols + t * kCLN,
                                a_lo0 + sel * (kCLBoxBytes >> 4) + t * (kTileM * 128 >> 4) + 2 * kk, b_lo0 + 2 * kk, kDescHi,
                                idesc, kk != 0);
          }
          umma_commit(&bar_acc_full[sel]);
          umma_commit(&bar_a_free[sel]);
        }
        __syncwarp();
      }
    }
  } else {
    // ============================================================ epilogue
    const int ew = warp - 2, q = warp & 3, grp = ew >> 2;      // TMEM lane quarter, tile group
    const int et = ew * 32 + lane;
    pdl_wait();
    int n, strip, blk;
    decode(it0, n, strip, blk);
    for (int it = it0, k = 0; it < it1; ++it, ++k, next_item(n, strip, blk)) {
      const uint32_t sel = uint32_t(k) & 1u;
      const int y0 = blk * kCLRows;
      const int rows = min(kCLRows, p.H - y0);
      // ---- bicubic x4 skip, horizontal pass (needs only the network input: done while the MMAs run).  Output row y
      // reads LR rows oy .. oy + 3, oy = (y >> 2) + (y & 3 < 2 ? -2 : -1), clamped to the image: the item's rows touch
      // LR rows lo .. lo + 7 at most.  48 global loads per output pixel made this epilogue the kernel's bound.
      const int h = p.H >> 2, w = p.W >> 2;
      const int lo = (y0 >> 2) - 2;
      {
        // thread -> one output column xs (512 threads = 8 (row, channel) pairs x 64 columns per sweep, 3 sweeps)
        const int xs = et & 63;
        const int x = min(strip * kStripW + xs, p.W - 1);
        const int rx = x & 3, ox = (x >> 2) + ((rx < 2) ? -2 : -1);
        int xi[4];
        float wx[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) { xi[j] = min(max(ox + j, 0), w - 1); wx[j] = c_bicubic_w[rx][j]; }
        const float* img = p.lr + size_t(n) * 3 * h * w;
        float v[3][4];
#pragma unroll
        for (int sw = 0; sw < 3; ++sw) {                       // all 12 loads in flight before the first use
          const int rc = (et >> 6) + 8 * sw, row = rc / 3, c = rc - 3 * row;
          const float* rowp = img + (c * h + min(max(lo + row, 0), h - 1)) * w;
#pragma unroll
          for (int j = 0; j < 4; ++j) v[sw][j] = __ldg(rowp + xi[j]);
        }
#pragma unroll
        for (int sw = 0; sw < 3; ++sw) {
          const int rc = (et >> 6) + 8 * sw, row = rc / 3, c = rc - 3 * row;
          float r = 0.f;
#pragma unroll
          for (int j = 0; j < 4; ++j) r = fmaf(wx[j], v[sw][j], r);
          s_rf[row][c][xs] = r;
        }
      }
      mbar_wait(&bar_acc_full[sel], (uint32_t(k) >> 1) & 1u);
      tc_fence_after();
      // ---- phase 1: D (TMEM lane = staged pixel of the tile, column = 3 tap + co) -> shared memory
#pragma unroll
      for (int tt = 0; tt < 2; ++tt) {
        const int t = grp + 4 * tt;
        if (t < kCLTiles && t * kTileM + q * 32 < kCLPx) {     // (the last tile holds 16 staged pixels: one warp's worth)
          uint32_t v[32];
          tmem_ld_32x32(tmem_base + sel * kCLSetCols + t * kCLN + (uint32_t(q * 32) << 16), v);
          tmem_ld_wait();
          const int s_px = t * kTileM + q * 32 + lane;
          if (s_px < kCLPx) {
#pragma unroll
            for (int tap = 0; tap < 9; ++tap)
              d_smem[tap * kCLPx + s_px] = make_float4(__uint_as_float(v[3 * tap]), __uint_as_float(v[3 * tap + 1]),
                                                       __uint_as_float(v[3 * tap + 2]), 0.f);
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&bar_acc_free[sel]);
      cl_bar_sync(1, kCLEpiThreads);
      // ---- phase 2: shifted sums, bias, bicubic x4 skip, clamp, stores.  Consecutive threads take consecutive columns.
      for (int idx = et; idx < rows * kStripW; idx += kCLEpiThreads) {
        const int yl = idx >> 6, xs = idx & 63;
        const int x = strip * kStripW + xs, y = y0 + yl;
        if (x >= p.W) continue;                                // (the last strip of a ragged width is partial)
        const float4* d0 = d_smem + (yl + 1) * kPitch + xs + 1;
        float o[3] = {s_bias[0], s_bias[1], s_bias[2]};
#pragma unroll
        for (int t = 0; t < 9; ++t) {
          const float4 d = d0[t * kCLPx + (t / 3 - 1) * kPitch + (t % 3 - 1)];
          o[0] += d.x; o[1] += d.y; o[2] += d.z;
        }
        const int ry = y & 3;
        const int r0 = (y >> 2) + ((ry < 2) ? -2 : -1) - lo;   // first of the four filtered LR rows, 0 .. 4
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          float accv = 0.f;
#pragma unroll
          for (int i = 0; i < 4; ++i) accv = fmaf(c_bicubic_w[ry][i], s_rf[r0 + i][c][xs], accv);
          float ov = o[c] + accv;
          if (!p.training) ov = fminf(fmaxf(ov, 0.f), 1.f);
          if (p.out_f32) p.out_f32[((size_t(n) * 3 + c) * p.H + y) * p.W + x] = ov;
          if (p.out_u8)    // the scripts' to_numpy (test_model.py:176-190): trunc(clip(v * 255, 0, 255)), HWC, optional BGR
            p.out_u8[((size_t(n) * p.H + y) * p.W + x) * 3 + (p.bgr ? 2 - c : c)] =
                uint8_t(int(fminf(fmaxf(__fmul_rn(ov, 255.0f), 0.f), 255.f)));
        }
      }
      cl_bar_sync(1, kCLEpiThreads);            // D may be overwritten by the next item
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, 2 * kCLSetCols);
}

// conv_last weights OIHW [3][64][3][3] -> bf16 [32][64]: row 3 t + co = W[co][:][t], rows 27 .. 31 zero
__global__ void pack_last_n_kernel(const float* __restrict__ w, bf16* __restrict__ dst) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= kCLN * kC) return;
  const int ci = i % kC, j = i / kC, t = j / 3, co = j % 3;
  dst[i] = __float2bfloat16(j < 27 ? w[(size_t(co) * kC + ci) * 9 + t] : 0.f);
}

}  // namespace fen
