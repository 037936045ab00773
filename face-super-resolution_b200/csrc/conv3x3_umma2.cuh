// Second generation of the per-layer 3x3 convolution kernel (see conv3x3_umma.cuh for the geometry, the
// epilogues and the references).  Same tiles, same epilogue code; what changed is what body2_umma.cuh taught:
//  * table-driven loops - the per-tile bookkeeping (ring position of the view, boxes to wait for / release,
//    image / strip of the tile) is computed once per CTA into shared memory; each issuer warp visits only its
//    own (even / odd) tiles.  The first-generation issuers walked every tile with a dozen integer divisions
//    and ran in lock-step: ~2 500 cycles per tile against 1 728 (N = 64) / 576 (conv_last) of MMA time.
//  * every mbarrier wait is executed by the whole, converged warp; elected-lane blocks hold straight-line
//    tcgen05 code only; a ring slot is only released by a warp that has seen it arrive.
// Used when a CTA's run fits the tables (<= 256 tiles, <= 192 boxes: batch 64 of the benchmark); larger runs
// take the first-generation kernel.
#pragma once
#include "conv3x3_umma.cuh"

#ifndef FEN_C2_ISSUE
#define FEN_C2_ISSUE 1   // shape of a tile's 36-MMA issue block: 0 unrolled, 1 loop over tap rows, 2 loop over taps (body2_umma.cuh)
#endif

namespace fen {

constexpr int kC2MaxTiles = 256;
constexpr int kC2MaxBoxes = 192;
constexpr int kC2RingPx = kRingSlots * kBoxPx;     // 792 pixels + the 264-pixel mirror slot

struct C2Tile { uint32_t m; uint8_t wait_upto, rel_upto, unit_lo, unit_hi; };   // m: view position relative to the first box
struct C2Box { int16_t n, x0, y0; uint8_t cont, pad; };
struct C2Unit { int16_t n, strip; };

template <int N>
__global__ void __launch_bounds__(ConvCfg<N>::kThreads, 1)
conv3x3_umma2_kernel(const __grid_constant__ CUtensorMap tm_in, const __grid_constant__ CUtensorMap tm_w,
                     const ConvParams p) {
  using Cfg = ConvCfg<N>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* w_smem = smem;                                   // [9][N][64] bf16, SWIZZLE_128B
  uint8_t* ring = smem + Cfg::kWBytes;                      // (kRingSlots + 1) x kSlotBytes
  __shared__ uint64_t bar_w[9], bar_full[kRingSlots], bar_empty[kRingSlots], bar_acc_full[Cfg::kAccBufs],
      bar_acc_empty[Cfg::kAccBufs];
  __shared__ uint32_t tmem_slot;
  __shared__ __align__(16) float s_bias[N];
  __shared__ __align__(16) float s_slope[kC];
  __shared__ C2Tile tile_tab[kC2MaxTiles];
  __shared__ C2Box box_tab[kC2MaxBoxes];
  __shared__ C2Unit unit_tab[kC2MaxTiles];                  // (a run may touch many short units)
  __shared__ int s_meta[2];

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  constexpr uint32_t kTmemCols = Cfg::kAccBufs * N;
  const int g_begin = blockIdx.x * p.tiles_per_cta;
  const int g_end = min(p.total_tiles, g_begin + p.tiles_per_cta);

  if (warp == 1) tmem_alloc(&tmem_slot, kTmemCols);
  if (tid == 0) {
    int b_cum = 0, i = 0, u = 0;
    int first_box[kC2MaxTiles + 2];
    for (int g = g_begin; g < g_end; ++u) {
      const Unit un = make_unit(p, g, g_end);
      unit_tab[u].n = int16_t(un.n); unit_tab[u].strip = int16_t(un.strip);
      for (int j = 0; j < un.nboxes; ++j) {
        C2Box e;
        e.n = int16_t(un.n); e.x0 = int16_t(un.strip * kStripW - 1); e.y0 = int16_t(un.ra - 1 + j * kBoxRows);
        e.cont = uint8_t(j > 0); e.pad = 0;
        box_tab[b_cum + j] = e;
      }
      for (int t = un.t0; t < un.t1; ++t, ++i) {
        const int base = kTileM * t - kPitch * un.ra;
        first_box[i] = b_cum + base / kBoxPx;
        C2Tile e;
        e.m = uint32_t(b_cum * kBoxPx + base) | (uint32_t(t) << 20);      // low 20 bits: position; high: tile in segment
        e.wait_upto = uint8_t(b_cum + min((base + kTileM + kMaxShift - 1) / kBoxPx, un.nboxes - 1) + 1);
        e.rel_upto = 0; e.unit_lo = uint8_t(u & 255); e.unit_hi = uint8_t(u >> 8);
        tile_tab[i] = e;
      }
      b_cum += un.nboxes;
      g += un.t1 - un.t0;
    }
    for (int k = 0; k < i; ++k) tile_tab[k].rel_upto = uint8_t((k + 2 < i) ? first_box[k + 2] : b_cum);
    s_meta[0] = i; s_meta[1] = b_cum;
    const uint32_t n_issuers = (i >= 2) ? 2u : 1u;
    for (int k = 0; k < 9; ++k) mbar_init(&bar_w[k], 1);
    for (int k = 0; k < kRingSlots; ++k) { mbar_init(&bar_full[k], 1); mbar_init(&bar_empty[k], n_issuers); }
    for (int k = 0; k < Cfg::kAccBufs; ++k) { mbar_init(&bar_acc_full[k], 1); mbar_init(&bar_acc_empty[k], Cfg::kEpiWarps); }
    fence_mbar_init();
    tma_prefetch_desc(&tm_in);
    tma_prefetch_desc(&tm_w);
  }
  for (int i = tid; i < N; i += Cfg::kThreads) s_bias[i] = p.bias ? p.bias[blockIdx.y * N + i] : 0.f;
  for (int i = tid; i < kC; i += Cfg::kThreads) s_slope[i] = p.slope ? p.slope[i] : 1.f;
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_slot;
  const int n_tiles = s_meta[0], n_boxes = s_meta[1];
  if (tid == 0) pdl_launch_dependents();     // (see conv3x3_umma.cuh: weights before the dependency wait, data after it)

  if (warp == 0) {
    // ============================================================ TMA producer
    // (whole-warp waits, lane 0 issues in straight-line blocks: see conv3x3_umma.cuh)
    if (n_tiles > 0) {
      if (lane == 0) {
#pragma unroll
        for (int tap = 0; tap < 9; ++tap) {
          mbar_expect_tx(&bar_w[tap], N * kC * 2);
          tma_load_2d(&tm_w, &bar_w[tap], w_smem + tap * N * 128, 0, (blockIdx.y * 9 + tap) * N);
        }
      }
      __syncwarp();
      pdl_wait();
      uint32_t slot = 0, use = 0;
      for (int b = 0; b < n_boxes; ++b) {
        const C2Box e = box_tab[b];
        mbar_wait(&bar_empty[slot], (use & 1u) ^ 1u);
        __syncwarp();
        if (lane == 0) {
          const bool mirror = e.cont && slot == 0;
          mbar_expect_tx(&bar_full[slot], mirror ? 2 * kSlotBytes : kSlotBytes);
          tma_load_4d(&tm_in, &bar_full[slot], smem_u32(ring + slot * kSlotBytes), 0, e.x0, e.y0, e.n);
          if (mirror) tma_load_4d(&tm_in, &bar_full[slot], smem_u32(ring + kRingSlots * kSlotBytes), 0, e.x0, e.y0, e.n);
        }
        __syncwarp();
        if (++slot == kRingSlots) { slot = 0; ++use; }
      }
    }
  } else if (warp < Cfg::kFirstEpiWarp) {
    // ============================================================ MMA issuers: warp 1 even tiles, warp 2 odd tiles
    constexpr uint32_t idesc = umma_idesc_bf16(kTileM, N);
    constexpr uint32_t kDescHi = (1024u >> 4) | (1u << 14) | (2u << 29);
    constexpr uint32_t kLbo = 1u << 16;
    const uint32_t ring_lo = (smem_u32(ring) >> 4) | kLbo;
    const uint32_t w_lo = (smem_u32(w_smem) >> 4) | kLbo;
    const bool leader = elect_one();
    const int wi = warp - 1;
    bool w_seen = false;
    uint32_t waited = 0, released = 0;
    for (int i = wi; i < n_tiles; i += 2) {
      const C2Tile e = tile_tab[i];
      const uint32_t acc = uint32_t(i) & (Cfg::kAccBufs - 1), aph = (uint32_t(i) / Cfg::kAccBufs) & 1;
      mbar_wait(&bar_acc_empty[acc], aph ^ 1);
      while (waited < e.wait_upto) {
        mbar_wait(&bar_full[waited % kRingSlots], (waited / kRingSlots) & 1u);
        ++waited;
      }
      __syncwarp();                               // converge after the spin-waits before any tcgen05 issue
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + acc * N;
      const uint32_t m = (e.m & 0xfffffu) % kC2RingPx;
#define FEN_C2_ISSUE_TAP(tap)                                                                        \
  {                                                                                                  \
    const uint32_t off = ((tap) / 3) * kPitch + ((tap) % 3);                                         \
    uint32_t pos = m + off;                                                                          \
    if (pos >= uint32_t(kC2RingPx)) pos -= kC2RingPx;                                                \
    const uint32_t a_lo = ring_lo + pos * 8;                                                         \
    const uint32_t b_lo = w_lo + (tap) * (N * 128 >> 4);                                             \
    _Pragma("unroll") for (int k = 0; k < 4; ++k)                                                    \
        umma_bf16_ss_lohi(d_tmem, a_lo + 2 * k, b_lo + 2 * k, kDescHi, idesc, ((tap) | k) != 0);     \
  }
      if (!w_seen) {
#pragma unroll
        for (int tap = 0; tap < 9; ++tap) {
          mbar_wait(&bar_w[tap], 0);
          __syncwarp();
          if (leader) FEN_C2_ISSUE_TAP(tap)
          __syncwarp();
        }
      } else if (leader) {
#if FEN_C2_ISSUE == 0
#pragma unroll
        for (int tap = 0; tap < 9; ++tap) FEN_C2_ISSUE_TAP(tap)
#else
        // a real loop instead of 36 unrolled MMAs: see body2_umma.cuh (the unrolled block spills uniform registers)
        uint32_t b_lo = w_lo, row = m;
#if FEN_C2_ISSUE == 1
#pragma unroll 1
        for (int dy = 0; dy < 3; ++dy) {
#pragma unroll
          for (int dx = 0; dx < 3; ++dx) {
            uint32_t pos = row + dx;
            if (pos >= uint32_t(kC2RingPx)) pos -= kC2RingPx;
            const uint32_t a_lo = ring_lo + pos * 8;
#pragma unroll
            for (int k = 0; k < 4; ++k)
              umma_bf16_ss_lohi_p(d_tmem, a_lo + 2 * k, b_lo + 2 * k, kDescHi, idesc, (dx | k) ? 1u : uint32_t(dy));
            b_lo += N * 128 >> 4;
          }
          row += kPitch;
          if (row >= uint32_t(kC2RingPx)) row -= kC2RingPx;
        }
#else
        uint32_t dx = 0;
#pragma unroll 1
        for (int tap = 0; tap < 9; ++tap) {
          uint32_t pos = row + dx;
          if (pos >= uint32_t(kC2RingPx)) pos -= kC2RingPx;
          const uint32_t a_lo = ring_lo + pos * 8;
#pragma unroll
          for (int k = 0; k < 4; ++k)
            umma_bf16_ss_lohi_p(d_tmem, a_lo + 2 * k, b_lo + 2 * k, kDescHi, idesc, k ? 1u : uint32_t(tap));
          b_lo += N * 128 >> 4;
          if (++dx == 3) {
            dx = 0;
            row += kPitch;
            if (row >= uint32_t(kC2RingPx)) row -= kC2RingPx;
          }
        }
#endif
#endif
      }
      w_seen = true;
      __syncwarp();
      while (waited < e.rel_upto) {               // never hand back a box this warp has not seen arrive
        mbar_wait(&bar_full[waited % kRingSlots], (waited / kRingSlots) & 1u);
        ++waited;
      }
      __syncwarp();
      while (released < e.rel_upto) {
        if (leader) umma_commit(&bar_empty[released % kRingSlots]);
        ++released;
      }
      if (leader) umma_commit(&bar_acc_full[acc]);
      __syncwarp();
    }
  } else {
    // ============================================================ epilogue (as in conv3x3_umma.cuh)
    constexpr int CW = Cfg::kColsPerWarp;
    const int q = warp & 3;
    const int half = (N == 64) ? ((warp - Cfg::kFirstEpiWarp) >> 2) : 0;
    const int col0 = half * CW;
    const int row_in_tile = q * 32 + lane;
    int cur_unit = -1, un_n = 0, un_strip = 0;
    float csum[CW];
#pragma unroll
    for (int c = 0; c < CW; ++c) csum[c] = 0.f;
    auto flush_sums = [&]() {
      if constexpr (N == kC) {
        if (p.epi == kEpiSum && cur_unit >= 0) {
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
          atomicAdd(p.sums + size_t(un_n) * kC + col0 + lane, csum[0]);
        }
      }
    };
    pdl_wait();
    for (int i = 0; i < n_tiles; ++i) {
      const C2Tile e = tile_tab[i];
      const int unit = int(e.unit_lo) | (int(e.unit_hi) << 8);
      if (unit != cur_unit) {
        flush_sums();
        cur_unit = unit; un_n = unit_tab[unit].n; un_strip = unit_tab[unit].strip;
#pragma unroll
        for (int c = 0; c < CW; ++c) csum[c] = 0.f;
      }
      const int t = int(e.m >> 20);
      const uint32_t acc = uint32_t(i) & (Cfg::kAccBufs - 1);
      mbar_wait(&bar_acc_full[acc], (uint32_t(i) / Cfg::kAccBufs) & 1);
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

      const int lin = kTileM * t + row_in_tile;       // strip-linear output pixel
      const int y = lin / kPitch, xs = lin - y * kPitch;
      const int x = un_strip * kStripW + xs;
      const bool valid = (xs < kStripW) && (y < p.H) && (x < p.W);   // (the last strip of a ragged width is partial)

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
            const float* src = p.lr + (size_t(un_n) * 3 + c) * h * w;
            float accv = 0.f;
#pragma unroll
            for (int ii = 0; ii < 4; ++ii) {
              const float* rowp = src + min(max(oy + ii, 0), h - 1) * w;
              float r = 0.f;
#pragma unroll
              for (int j = 0; j < 4; ++j) r = fmaf(c_bicubic_w[rx][j], __ldg(rowp + xi[j]), r);
              accv = fmaf(c_bicubic_w[ry][ii], r, accv);
            }
            float o = __uint_as_float(v[c]) + s_bias[c] + accv;
            if (!p.training) o = fminf(fmaxf(o, 0.f), 1.f);
            if (p.out_f32) p.out_f32[((size_t(un_n) * 3 + c) * p.H + y) * p.W + x] = o;
            if (p.out_u8)    // the scripts' to_numpy (test_model.py:176-190): trunc(clip(v * 255, 0, 255)), HWC, optional BGR
              p.out_u8[((size_t(un_n) * p.H + y) * p.W + x) * 3 + (p.bgr ? 2 - c : c)] =
                  uint8_t(int(fminf(fmaxf(__fmul_rn(o, 255.0f), 0.f), 255.f)));
          }
        }
      } else {
        float f[CW];
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
            opix = (size_t(un_n) * (2 * p.H) + (2 * y + (sub >> 1))) * (2 * p.W) + (2 * x + (sub & 1));
          } else {
            opix = (size_t(un_n) * p.H + y) * p.W + x;
          }
          bf16* dst = p.out + opix * kC + col0;
          if (p.epi == kEpiResidual) {
            const bf16* rsd = p.residual + opix * kC + col0;
#pragma unroll
            for (int j = 0; j < CW / 16; ++j) {
              uint32_t r[8];
              ld_global_cg_256(rsd + 16 * j, r);
#pragma unroll
              for (int ee = 0; ee < 8; ++ee) {
                f[16 * j + 2 * ee] += bf16lo(r[ee]);
                f[16 * j + 2 * ee + 1] += bf16hi(r[ee]);
              }
            }
          }
#pragma unroll
          for (int j = 0; j < CW / 16; ++j) {
            uint32_t o[8];
#pragma unroll
            for (int ee = 0; ee < 8; ++ee) o[ee] = pack_bf16(f[16 * j + 2 * ee], f[16 * j + 2 * ee + 1]);
            st_global_256(dst + 16 * j, o);
          }
        }
      }
    }
    flush_sums();
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, kTmemCols);
}

}  // namespace fen
