// 3x3 / pad 1 convolution, 64 input channels, NHWC bf16, as an implicit GEMM on tcgen05 + TMEM,
// fed by TMA.  One kernel serves every 64-input conv of FaceEnhanceNet
// (reference: src/models/blocks.py:122-130,145-147 RCAB convs; :181-189 group conv; :210-226
// upsample conv + PixelShuffle + PReLU; src/models/custom.py:109-124,172-188 conv_after_body,
// conv_last + bicubic skip + clamp).
//
// Geometry.  An image is cut into vertical strips of 64 columns.  A strip is staged in shared memory
// as rows of 66 pixels (1 halo pixel left and right, zero-filled by TMA out-of-bounds handling at
// image borders), each pixel = 64 channels = 128 B, SWIZZLE_128B.  In this "strip-linear" pixel
// space every 3x3 tap is a constant offset (dy+1)*66 + (dx+1), so the A operand of tap (dy,dx) for
// the 128 output pixels [128 t, 128 t + 128) is simply the 128 consecutive smem pixels starting at
// 128 t + offset: nine shifted views of ONE staged copy, no im2col and no re-load (the UMMA smem
// descriptor start address only needs 16 B alignment; the swizzle is a function of the absolute
// smem address, verified on B200 by tools/umma_probe.cu).  Two of every 66 output pixels are halo
// columns and are discarded by the epilogue (3 % of MMA work).
//
// Pipeline.  warp 0: TMA producer (4-row boxes into a 3-slot ring + a mirror slot that keeps views
// contiguous across the ring wrap).  warps 1-2: tcgen05.mma issuers (one elected lane each), 9 taps
// x 4 k-steps per tile, alternating tiles: the tensor pipe accepts only ~1-2 queued MMAs, so the
// ~800 cycles of barrier checks between two tiles of one issuer are covered by the other issuer's
// MMAs.  fp32 accumulators live in TMEM, 4 buffers (2 per issuer).  warps 3..: epilogue, TMEM ->
// registers -> bias / PReLU / residual / SE partial sums / PixelShuffle / bicubic skip -> global,
// written straight from registers (256-bit stores) so shared-memory bandwidth - the binding
// resource: every MMA re-reads its A and B operands, tools/umma_probe2.cu - is left to the tensor
// core.
#pragma once
#include "fen_common.cuh"
#include "ptx_sm100.cuh"

#ifndef FEN_C1_TURN
#define FEN_C1_TURN 0   // 1: the two MMA issuers of conv3x3_umma_kernel take turns, one whole tile each (measured: 3.63 vs 3.60 ms per batch-64 forward - its CTAs run 228-tile passes without the body kernel's per-pass waits; off)
#endif

namespace fen {

__device__ __forceinline__ void tma_load_4d(const CUtensorMap* m, uint64_t* bar, uint32_t dst_smem,
                                            int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4, %5, %6}], [%2];"
      :
      : "r"(dst_smem), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1),
        "r"(c2), "r"(c3)
      : "memory");
}

// A unit = a run of tiles [t0, t1) inside one segment (one strip of one image).
struct Unit {
  int n, strip, t0, t1;  // image, strip, tile range
  int ra;                // first staged row, in the shifted row space (row 0 = image row -1)
  int nboxes;            // 4-row boxes staged for this unit
};

__device__ __forceinline__ Unit make_unit(const ConvParams& p, int g, int g_end) {
  Unit u;
  const int seg = g / p.tiles_per_seg;
  u.t0 = g - seg * p.tiles_per_seg;
  u.t1 = min(p.tiles_per_seg, u.t0 + (g_end - g));
  u.n = seg / p.strips;
  u.strip = seg - u.n * p.strips;
  u.ra = (kTileM * u.t0) / kPitch;
  int rb = (kTileM * u.t1 + kMaxShift - 1) / kPitch;  // last shifted row any view touches
  rb = min(rb, p.H + 1);
  u.nboxes = (rb - u.ra) / kBoxRows + 1;
  return u;
}

template <int N>
struct ConvCfg {
  static constexpr int kWBytes = 9 * N * kC * 2;
  static constexpr int kDynBytes = kWBytes + kRingBytes + 1024;  // + alignment slack
  // epilogue warps: N = 64 -> 8 (lane quarter x column half), N = 16 -> 4
  static constexpr int kEpiWarps = (N == 64) ? 8 : 4;
  static constexpr int kMmaWarps = 2;
  static constexpr int kFirstEpiWarp = 1 + kMmaWarps;
  static constexpr int kThreads = 32 * (1 + kMmaWarps + kEpiWarps);
  static constexpr int kAccBufs = 4;
  static constexpr int kColsPerWarp = (N == 64) ? 32 : N;
};

__device__ __forceinline__ float bf16lo(uint32_t v) { return __uint_as_float(v << 16); }
__device__ __forceinline__ float bf16hi(uint32_t v) { return __uint_as_float(v & 0xffff0000u); }
__device__ __forceinline__ uint32_t pack_bf16(float a, float b) {
  __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}
__device__ __forceinline__ void st_global_256(void* ptr, const uint32_t (&v)[8]) {
  // L1::no_allocate: the line is never read back by this SM; allocating it in L1 cost 5 % of a forward
  asm volatile("st.global.L1::no_allocate.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(ptr), "r"(v[0]),
               "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7])
               : "memory");
}
__device__ __forceinline__ void ld_global_nc_256(const void* ptr, uint32_t (&v)[8]) {
  asm volatile("ld.global.nc.v8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]),
                 "=r"(v[7])
               : "l"(ptr));
}

// Cubic-convolution phase filters of F.interpolate(scale_factor=4, mode='bicubic',
// align_corners=False) (A = -0.75), times 2048; output d = 4q + r reads q + off[r] - 1 .. + 2.
__device__ __constant__ float c_bicubic_w[4][4] = {{-135.f / 2048.f, 873.f / 2048.f, 1535.f / 2048.f, -225.f / 2048.f},
                                                   {-21.f / 2048.f, 235.f / 2048.f, 1981.f / 2048.f, -147.f / 2048.f},
                                                   {-147.f / 2048.f, 1981.f / 2048.f, 235.f / 2048.f, -21.f / 2048.f},
                                                   {-225.f / 2048.f, 1535.f / 2048.f, 873.f / 2048.f, -135.f / 2048.f}};

template <int N>
__global__ void __launch_bounds__(ConvCfg<N>::kThreads, 1)
conv3x3_umma_kernel(const __grid_constant__ CUtensorMap tm_in, const __grid_constant__ CUtensorMap tm_w,
                    const ConvParams p) {
  using Cfg = ConvCfg<N>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* w_smem = smem;                                   // [9][N][64] bf16, SWIZZLE_128B
  uint8_t* ring = smem + Cfg::kWBytes;                      // (kRingSlots + 1) x kSlotBytes
  __shared__ uint64_t bar_w[9], bar_full[kRingSlots], bar_empty[kRingSlots], bar_acc_full[Cfg::kAccBufs], bar_acc_empty[Cfg::kAccBufs];
  __shared__ uint64_t bar_turn[4];                          // FEN_C1_TURN: tile T's issuer arrives on [T % 4] once its MMAs are issued
  __shared__ uint32_t tmem_slot;
  __shared__ __align__(16) float s_bias[N];
  __shared__ __align__(16) float s_slope[kC];
  __shared__ __align__(16) float s_islope[kC];              // 1 / slope (0 where the slope is 0): kEpiGate

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  constexpr uint32_t kTmemCols = Cfg::kAccBufs * N;  // accumulator buffers (power of two >= 32)
  static_assert(kTmemCols >= 32 && (kTmemCols & (kTmemCols - 1)) == 0 && kTmemCols <= 512, "TMEM columns");

  const int g_begin = blockIdx.x * p.tiles_per_cta;
  const int g_end = min(p.total_tiles, g_begin + p.tiles_per_cta);

  if (warp == 1) tmem_alloc(&tmem_slot, kTmemCols);
  if (tid == 0) {
    for (int i = 0; i < 9; ++i) mbar_init(&bar_w[i], 1);
    for (int i = 0; i < kRingSlots; ++i) { mbar_init(&bar_full[i], 1); mbar_init(&bar_empty[i], Cfg::kMmaWarps); }
    for (int i = 0; i < Cfg::kAccBufs; ++i) { mbar_init(&bar_acc_full[i], 1); mbar_init(&bar_acc_empty[i], Cfg::kEpiWarps); }
    for (int i = 0; i < 4; ++i) mbar_init(&bar_turn[i], 1);
    fence_mbar_init();
    tma_prefetch_desc(&tm_in);
    tma_prefetch_desc(&tm_w);
  }
  for (int i = tid; i < N; i += Cfg::kThreads) s_bias[i] = p.bias ? p.bias[blockIdx.y * N + i] : 0.f;
  for (int i = tid; i < kC; i += Cfg::kThreads) {
    const float sl = p.slope ? p.slope[i] : 1.f;
    s_slope[i] = sl;
    s_islope[i] = sl != 0.f ? 1.f / sl : 0.f;
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_slot;
  if (tid == 0) pdl_launch_dependents();     // the next launch may begin its prologue (it waits before touching data)

  if (warp == 0) {
    // ============================================================ TMA producer
    // (every wait is executed by the whole, converged warp; lane 0 only issues, in straight-line blocks: TMA issue is a
    // uniform-datapath instruction and must not share an elected-lane block with a spin loop - DESIGN.md 4.2, rule 1)
    if (g_begin < g_end) {
      // Weights arrive tap by tap (one barrier each) so the first MMAs need not wait for all 9.  They (like the
      // bias / slope vectors above) were packed long before this launch's predecessor started: all nine taps go
      // out before the dependency wait, the activation boxes after it.
      if (lane == 0) {
#pragma unroll
        for (int tap = 0; tap < 9; ++tap) {
          mbar_expect_tx(&bar_w[tap], N * kC * 2);
          tma_load_2d(&tm_w, &bar_w[tap], w_smem + tap * N * 128, 0, (blockIdx.y * 9 + tap) * N);
        }
      }
      __syncwarp();
      pdl_wait();
      uint32_t gb = 0;  // running box counter of this CTA
      for (int g = g_begin; g < g_end;) {
        const Unit u = make_unit(p, g, g_end);
        const int x0 = u.strip * kStripW - 1;
        for (int j = 0; j < u.nboxes; ++j, ++gb) {
          const uint32_t slot = gb % kRingSlots, ph = (gb / kRingSlots) & 1;
          mbar_wait(&bar_empty[slot], ph ^ 1);
          __syncwarp();
          if (lane == 0) {
            const bool mirror = (slot == kRingSlots - 1) && (j + 1 < u.nboxes);
            mbar_expect_tx(&bar_full[slot], mirror ? 2 * kSlotBytes : kSlotBytes);
            const int y0 = u.ra - 1 + j * kBoxRows;
            tma_load_4d(&tm_in, &bar_full[slot], smem_u32(ring + slot * kSlotBytes), 0, x0, y0, u.n);
            if (mirror)
              tma_load_4d(&tm_in, &bar_full[slot], smem_u32(ring + kRingSlots * kSlotBytes), 0, x0,
                          y0 + kBoxRows, u.n);
          }
          __syncwarp();
        }
        g += u.t1 - u.t0;
      }
    }
  } else if (warp < Cfg::kFirstEpiWarp) {
    // ============================================================ MMA issuers
    // Both issuer warps walk every tile (same bookkeeping, both arrive on the ring's empty barriers)
    // but issue MMAs only for the tiles of their parity.  The whole warp runs the (warp-uniform)
    // control flow so address math stays on the uniform datapath; one elected lane issues
    // tcgen05.mma / tcgen05.commit.  Per MMA only the low
    // descriptor word changes (start address); the high word (SBO = 1024 B, version 1, SWIZZLE_128B)
    // is a constant.
    if (g_begin < g_end) {
      constexpr uint32_t idesc = umma_idesc_bf16(kTileM, N);
      constexpr uint32_t kDescHi = (1024u >> 4) | (1u << 14) | (2u << 29);
      constexpr uint32_t kLbo = 1u << 16;
      const uint32_t ring_lo = (smem_u32(ring) >> 4) | kLbo;
      const uint32_t w_lo = (smem_u32(w_smem) >> 4) | kLbo;
      const bool leader = elect_one();
      const uint32_t my_parity = warp - 1;
      bool first = true;
      long long dbg_t0 = p.dbg ? clock64() : 0, dbg_acc = 0, dbg_full = 0, dbg_issue = 0;
      uint32_t gb_base = 0, tile_ctr = 0;
      for (int g = g_begin; g < g_end;) {
        const Unit u = make_unit(p, g, g_end);
        int waited = 0, released = 0;
        for (int t = u.t0; t < u.t1; ++t, ++tile_ctr) {
          const uint32_t acc = tile_ctr & (Cfg::kAccBufs - 1);
          const bool mine = (tile_ctr & 1) == my_parity;
          long long tw = p.dbg ? clock64() : 0;
          if (mine) mbar_wait(&bar_acc_empty[acc], ((tile_ctr / Cfg::kAccBufs) & 1) ^ 1);
          if (p.dbg) { const long long n = clock64(); dbg_acc += n - tw; tw = n; }
          const int base = kTileM * t - kPitch * u.ra;
          const int need_last = min((base + kTileM + kMaxShift - 1) / kBoxPx, u.nboxes - 1);
          while (mine && waited <= need_last) {   // each issuer confirms every box it reads itself
            const uint32_t gb = gb_base + waited;
            mbar_wait(&bar_full[gb % kRingSlots], (gb / kRingSlots) & 1);
            ++waited;
          }
          if (p.dbg) dbg_full += clock64() - tw;
          tc_fence_after();
          const uint32_t d_tmem = tmem_base + acc * N;
          // views of this tile start in box lb0 or lb0 + 1 (tap offsets are < one box)
          const int lb0 = base / kBoxPx, r0 = base - lb0 * kBoxPx;
          const uint32_t slot0 = (gb_base + lb0) % kRingSlots;
          const uint32_t slot1 = (slot0 + 1 == kRingSlots) ? 0 : slot0 + 1;
          const uint32_t a0 = ring_lo + slot0 * (kSlotBytes >> 4) + r0 * 8;
          const uint32_t a1 = ring_lo + slot1 * (kSlotBytes >> 4) + (r0 - kBoxPx) * 8;
          const long long dbg_i0 = p.dbg ? clock64() : 0;
#if FEN_C1_TURN
          // the two issuers take turns, a whole tile each, in tile order (body2_umma.cuh: issuing concurrently they fall
          // into lock-step and do their per-tile bookkeeping at the same time, with the pipe idle)
          if (mine && tile_ctr > 0) mbar_wait(&bar_turn[(tile_ctr - 1) & 3u], ((tile_ctr - 1) >> 2) & 1u);
          __syncwarp();
#endif
          if (leader && mine) {
            if (first) {
#pragma unroll
              for (int tap = 0; tap < 9; ++tap) {
                mbar_wait(&bar_w[tap], 0);
                const int off = (tap / 3) * kPitch + (tap % 3);
                const uint32_t a_lo = ((r0 + off < kBoxPx) ? a0 : a1) + off * 8;
                const uint32_t b_lo = w_lo + tap * (N * 128 >> 4);
#pragma unroll
                for (int k = 0; k < 4; ++k)
                  umma_bf16_ss_lohi(d_tmem, a_lo + 2 * k, b_lo + 2 * k, kDescHi, idesc, (tap | k) != 0);
              }
            } else {
              // a real loop over the taps: 36 unrolled MMAs with 72 distinct descriptors spill the uniform registers
              uint32_t b_lo = w_lo;
              int off_row = 0, dx = 0;
#pragma unroll 1
              for (int tap = 0; tap < 9; ++tap) {
                const int off = off_row + dx;
                const uint32_t a_lo = ((r0 + off < kBoxPx) ? a0 : a1) + off * 8;
#pragma unroll
                for (int k = 0; k < 4; ++k)
                  umma_bf16_ss_lohi_p(d_tmem, a_lo + 2 * k, b_lo + 2 * k, kDescHi, idesc, k ? 1u : uint32_t(tap));
                b_lo += N * 128 >> 4;
                if (++dx == 3) { dx = 0; off_row += kPitch; }
              }
            }
#if FEN_C1_TURN
            mbar_arrive(&bar_turn[tile_ctr & 3u]);
#endif
          }
          __syncwarp();
          if (p.dbg) dbg_issue += clock64() - dbg_i0;
          // boxes that no later tile of this unit will read can go back to the producer
          const int next_first = (t + 1 < u.t1) ? (base + kTileM) / kBoxPx : u.nboxes;
          while (released < next_first) {
            if (leader) umma_commit(&bar_empty[(gb_base + released) % kRingSlots]);
            ++released;
          }
          if (leader && mine) umma_commit(&bar_acc_full[acc]);
          if (mine) first = false;
          __syncwarp();
        }
        gb_base += u.nboxes;
        g += u.t1 - u.t0;
      }
      if (p.dbg && leader && warp == 1) {
        p.dbg[blockIdx.x * 8 + 2] = dbg_acc;               // waiting for a free accumulator
        p.dbg[blockIdx.x * 8 + 3] = dbg_full;              // waiting for TMA data
        p.dbg[blockIdx.x * 8 + 4] = clock64() - dbg_t0;    // MMA warp total
        p.dbg[blockIdx.x * 8 + 5] = tile_ctr;
        p.dbg[blockIdx.x * 8 + 1] = dbg_issue;             // inside the 36-MMA issue loops
      }
    }
  } else {
    // ============================================================ epilogue
    // thread <-> one TMEM lane (= one output pixel of the tile); a warp may only touch the lane
    // quarter (warp % 4); for N = 64 two warps share a quarter and take 32 columns each.
    constexpr int CW = Cfg::kColsPerWarp;
    const int q = warp & 3;
    const int half = (N == 64) ? ((warp - Cfg::kFirstEpiWarp) >> 2) : 0;
    const int col0 = half * CW;
    const int row_in_tile = q * 32 + lane;
    uint32_t tile_ctr = 0;
    long long dbg_e0 = p.dbg ? clock64() : 0, dbg_ewait = 0, dbg_etail = 0, dbg_eld = 0;
    pdl_wait();                                // everything below reads / writes what predecessors produced / still read
    for (int g = g_begin; g < g_end;) {
      const Unit u = make_unit(p, g, g_end);
      float csum[CW];
#pragma unroll
      for (int c = 0; c < CW; ++c) csum[c] = 0.f;
      for (int t = u.t0; t < u.t1; ++t, ++tile_ctr) {
        const uint32_t acc = tile_ctr & (Cfg::kAccBufs - 1);
        const int lin = kTileM * t + row_in_tile;       // strip-linear output pixel
        const int y = lin / kPitch, xs = lin - y * kPitch;
        const int x = u.strip * kStripW + xs;
        const bool valid = (xs < kStripW) && (y < p.H) && (x < p.W);   // (the last strip of a ragged width is partial)
        // Second operands of the epilogue (residual / saved activation / SE operand, PReLU sign bits): requested BEFORE
        // waiting for the accumulator, so their L2 / DRAM latency hides behind the tile's MMAs (the backward's kEpiGate
        // and kEpiDot launches were epilogue-bound: 32 us against 19 us for a plain convolution at batch 32).
        uint32_t ra[(N == kC) ? 16 : 1], rb[(N == kC) ? 16 : 1], pos_bits = 0;
        if constexpr (N == kC) {
          const size_t ipix = (size_t(u.n) * p.H + y) * p.W + x;
          const bool want_a = p.epi == kEpiGate || p.epi == kEpiResidual || (p.epi == kEpiDot && p.residual != nullptr);
          if (valid && want_a) {
            const bf16* ap = p.residual + ipix * kC + col0;
            ld_global_cg_256(ap, *reinterpret_cast<uint32_t(*)[8]>(&ra[0]));
            ld_global_cg_256(ap + 16, *reinterpret_cast<uint32_t(*)[8]>(&ra[8]));
          } else {
#pragma unroll
            for (int e = 0; e < 16; ++e) ra[e] = 0u;
          }
          if (valid && p.epi == kEpiDot) {
            const bf16* xp = p.aux + ipix * kC + col0;
            ld_global_cg_256(xp, *reinterpret_cast<uint32_t(*)[8]>(&rb[0]));
            ld_global_cg_256(xp + 16, *reinterpret_cast<uint32_t(*)[8]>(&rb[8]));
          } else {
#pragma unroll
            for (int e = 0; e < 16; ++e) rb[e] = 0u;
          }
          if (valid && p.epi == kEpiGate) pos_bits = __ldcg(p.mask_in + ipix * 2 + half);
        }
        long long dbg_w0 = p.dbg ? clock64() : 0;
        mbar_wait(&bar_acc_full[acc], (tile_ctr / Cfg::kAccBufs) & 1);
        const long long dbg_w1 = p.dbg ? clock64() : 0;
        if (p.dbg) dbg_ewait += dbg_w1 - dbg_w0;
        tc_fence_after();
        uint32_t v[CW];
        const uint32_t taddr = tmem_base + acc * N + col0 + (uint32_t(q * 32) << 16);
        if constexpr (CW == 16) {
          tmem_ld_32x16(taddr, v);
        } else {
          tmem_ld_32x32(taddr, v);
        }
        tmem_ld_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&bar_acc_empty[acc]);
        if (p.dbg) dbg_eld += clock64() - dbg_w1;

        if constexpr (N == 16) {
          // ---- conv_last: + bias + bicubic x4 skip (+ clamp), fp32 NCHW
          if (valid) {
            const int h = p.H >> 2, w = p.W >> 2;
            const int qy = y >> 2, ry = y & 3, qx = x >> 2, rx = x & 3;
            const int oy = qy + ((ry < 2) ? -2 : -1), ox = qx + ((rx < 2) ? -2 : -1);
            int xi[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) xi[j] = min(max(ox + j, 0), w - 1);
#pragma unroll
            for (int c = 0; c < 3; ++c) {
              const float* src = p.lr + (size_t(u.n) * 3 + c) * h * w;
              float accv = 0.f;
#pragma unroll
              for (int i = 0; i < 4; ++i) {
                const float* rowp = src + min(max(oy + i, 0), h - 1) * w;
                float r = 0.f;
#pragma unroll
                for (int j = 0; j < 4; ++j) r = fmaf(c_bicubic_w[rx][j], __ldg(rowp + xi[j]), r);
                accv = fmaf(c_bicubic_w[ry][i], r, accv);
              }
              float o = __uint_as_float(v[c]) + s_bias[c] + accv;
              if (!p.training) o = fminf(fmaxf(o, 0.f), 1.f);
              if (p.out_f32) p.out_f32[((size_t(u.n) * 3 + c) * p.H + y) * p.W + x] = o;
              if (p.out_u8)    // the scripts' to_numpy (test_model.py:176-190): trunc(clip(v * 255, 0, 255)), HWC, optional BGR
                p.out_u8[((size_t(u.n) * p.H + y) * p.W + x) * 3 + (p.bgr ? 2 - c : c)] =
                    uint8_t(int(fminf(fmaxf(__fmul_rn(o, 255.0f), 0.f), 255.f)));
            }
          }
        } else {
          float f[CW];
          uint32_t mbits = 0;
#pragma unroll
          for (int j = 0; j < CW / 4; ++j) {
            const float4 b4 = *reinterpret_cast<const float4*>(&s_bias[col0 + 4 * j]);
            f[4 * j + 0] = __uint_as_float(v[4 * j + 0]) + b4.x;
            f[4 * j + 1] = __uint_as_float(v[4 * j + 1]) + b4.y;
            f[4 * j + 2] = __uint_as_float(v[4 * j + 2]) + b4.z;
            f[4 * j + 3] = __uint_as_float(v[4 * j + 3]) + b4.w;
          }
          if (p.epi == kEpiSum) {
            if (valid) {
#pragma unroll
              for (int c = 0; c < CW; ++c) csum[c] += f[c];
            }
          } else if (p.epi == kEpiPrelu || p.epi == kEpiShuffle) {
            if (p.mask_out) {                     // training forward: the sign bits of the pre-activation
#pragma unroll
              for (int c = 0; c < CW; ++c) mbits |= (f[c] > 0.f ? 1u : 0u) << c;
            }
#pragma unroll
            for (int j = 0; j < CW / 4; ++j) {
              const float4 s4 = *reinterpret_cast<const float4*>(&s_slope[col0 + 4 * j]);
              f[4 * j + 0] = f[4 * j + 0] > 0.f ? f[4 * j + 0] : f[4 * j + 0] * s4.x;
              f[4 * j + 1] = f[4 * j + 1] > 0.f ? f[4 * j + 1] : f[4 * j + 1] * s4.y;
              f[4 * j + 2] = f[4 * j + 2] > 0.f ? f[4 * j + 2] : f[4 * j + 2] * s4.z;
              f[4 * j + 3] = f[4 * j + 3] > 0.f ? f[4 * j + 3] : f[4 * j + 3] * s4.w;
            }
          }
          if (valid) {
            size_t opix;
            if (p.epi == kEpiShuffle) {
              const int sub = blockIdx.y;
              opix = (size_t(u.n) * (2 * p.H) + (2 * y + (sub >> 1))) * (2 * p.W) + (2 * x + (sub & 1));
            } else {
              opix = (size_t(u.n) * p.H + y) * p.W + x;
            }
            size_t spix = opix;
            if (p.unshuffle)
              spix = ((size_t(2 * (y & 1) + (x & 1)) * p.B + u.n) * (p.H >> 1) + (y >> 1)) * (p.W >> 1) + (x >> 1);
            bf16* dst = p.out + spix * kC + col0;
            if ((p.epi == kEpiPrelu || p.epi == kEpiShuffle) && p.mask_out) p.mask_out[opix * 2 + half] = mbits;
            if (p.epi == kEpiGate) {
#pragma unroll
              for (int e = 0; e < 16; ++e) {
#pragma unroll
                for (int hlf = 0; hlf < 2; ++hlf) {
                  const int c = 2 * e + hlf;
                  const float a = hlf ? bf16hi(ra[e]) : bf16lo(ra[e]);
                  const bool pos = (pos_bits >> c) & 1u;
                  // negative side: pre-activation z = a / slope (slope == 0 loses z: that term is dropped)
                  csum[c] += pos ? 0.f : f[c] * (a * s_islope[col0 + c]);
                  f[c] = pos ? f[c] : f[c] * s_slope[col0 + c];
                }
              }
            }
            if (p.epi == kEpiResidual || (p.epi == kEpiDot && p.residual != nullptr)) {
#pragma unroll
              for (int e = 0; e < 16; ++e) {
                f[2 * e] += bf16lo(ra[e]);
                f[2 * e + 1] += bf16hi(ra[e]);
              }
            }
            if (p.epi == kEpiDot) {
#pragma unroll
              for (int e = 0; e < 16; ++e) {
                csum[2 * e] = fmaf(f[2 * e], bf16lo(rb[e]), csum[2 * e]);
                csum[2 * e + 1] = fmaf(f[2 * e + 1], bf16hi(rb[e]), csum[2 * e + 1]);
              }
            }
#pragma unroll
            for (int j = 0; j < CW / 16; ++j) {
              uint32_t o[8];
#pragma unroll
              for (int e = 0; e < 8; ++e) o[e] = pack_bf16(f[16 * j + 2 * e], f[16 * j + 2 * e + 1]);
              st_global_256(dst + 16 * j, o);
            }
          }
        }
      }
      long long dbg_t2 = p.dbg ? clock64() : 0;
      if constexpr (N == kC) {
        if (p.epi == kEpiSum || p.epi == kEpiGate || p.epi == kEpiDot) {
          // Per-image channel sums for the squeeze-and-excitation pool (kEpiGate: one vector for the whole batch).  Reduce-scatter butterfly
          // over the warp: 16+8+4+2+1 shuffles leave lane l with the total of channel col0 + l
          // (the lane bits 16,8,4,2,1 select the upper/lower half kept at each level).
#pragma unroll
          for (int d = 16, len = CW; d >= 1; d >>= 1, len >>= 1) {
            const bool hi = (lane & d) != 0;
#pragma unroll
            for (int i = 0; i < len / 2; ++i) {
              const float send = hi ? csum[i] : csum[i + len / 2];
              const float keep = hi ? csum[i + len / 2] : csum[i];
              csum[i] = keep + __shfl_xor_sync(0xffffffffu, send, d);
            }
          }
          const size_t si = (p.epi == kEpiGate ? size_t(0) : size_t(u.n) * kC) + col0 + lane;
          if (p.sums64) { if (p.epi == kEpiSum) hs_add(p.sums64 + si, csum[0]); else gs_add(p.sums64 + si, csum[0]); }
          else atomicAdd(p.sums + si, csum[0]);
        }
      }
      if (p.dbg) dbg_etail += clock64() - dbg_t2;
      g += u.t1 - u.t0;
    }
    if (p.dbg && tid == 32 * Cfg::kFirstEpiWarp) {
      p.dbg[blockIdx.x * 8 + 0] = dbg_eld;                 // TMEM load + release of the accumulator
      p.dbg[blockIdx.x * 8 + 6] = dbg_ewait;               // epilogue warp waiting for accumulators
      p.dbg[blockIdx.x * 8 + 7] = clock64() - dbg_e0;      // epilogue warp total
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, kTmemCols);
}

}  // namespace fen
