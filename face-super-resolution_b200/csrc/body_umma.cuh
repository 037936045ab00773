// Persistent "body" kernel: ALL 64->64 3x3 convolutions of the residual body
// (num_groups x (2 x blocks_per_group + 1) + conv_after_body = 127 layers for the 6 x 10 model)
// in ONE launch, with the squeeze-and-excitation scale and the RCAB residual folded into the
// producer of the following conv.  Reference: src/models/custom.py:167-175 (body loop, long skip),
// src/models/blocks.py:135-153 (RCAB), :75-92 (ChannelAttention), :185-189 (ResidualGroup).
//
// Every CTA owns the same run of output tiles in every layer (all body layers share one geometry),
// so a layer boundary is not a grid-wide barrier: a CTA only waits for the CTAs that share an image
// with it ("peers": halo rows + the per-image SE pool) through release/acquire flags in global
// memory.  Launched cooperatively with one CTA per SM, so all CTAs are co-resident.
//
// Per layer the tile pipeline is the one of conv3x3_umma.cuh (ring of boxes with a mirror slot -
// here 2-row boxes, 7 slots -, two tcgen05.mma issuer warps, 8 epilogue warps, TMEM accumulators).  Differences:
//   * warp 0 issues all TMA loads.  Fused layers (conv1 of RCAB b > 0, group conv): TMA brings x
//     into the ring slot, the 256 transform threads (warps 1-8) read o (conv2 output of the
//     previous block) from global memory one box ahead, update the slot in place with
//     x' = x + (res_scale * s[c]) * o (fp32 math, bf16 store, SWIZZLE_128B pattern by hand) and
//     write the rows they own back to global memory as the next residual stream - no standalone
//     elementwise pass exists.
//   * the SE vector s = sigmoid(W2 relu(W0 mean(o))) is recomputed per CTA from the per-image
//     channel sums the conv2 epilogue accumulated (2 KMAC, fp32).
//   * biases and PReLU slopes come from __constant__ memory (the epilogue must stay off shared
//     memory, whose bandwidth the tensor core needs), weights of the next layer are loaded tap by
//     tap as soon as the current layer has issued its last MMA on that tap.
#pragma once
#include "conv3x3_umma.cuh"

namespace fen {

#ifndef FEN_BODY_DEBUG
#define FEN_BODY_DEBUG 0   // 1: per-CTA cycle counters into BodyParams::dbg (developer builds)
#endif
#define BDBG (FEN_BODY_DEBUG && p.dbg)

constexpr int kBodyXformWarps = 8;                    // warps 1..8: x' = x + s*o transform (fused layers)
constexpr int kBodyMmaWarps = 2;
constexpr int kBodyEpiWarps = 8;
constexpr int kBodyFirstXformWarp = 1;                // warp 0: TMA issuer
constexpr int kBodyFirstMmaWarp = 1 + kBodyXformWarps;
constexpr int kBodyFirstEpiWarp = kBodyFirstMmaWarp + kBodyMmaWarps;
constexpr int kBodyThreads = 32 * (kBodyFirstEpiWarp + kBodyEpiWarps);   // 608
constexpr int kBodyAccBufs = 4;
constexpr int kBodyWBytes = 9 * kC * kC * 2;
// activation ring: 2-row boxes (132 px, 16 896 B; TMA SWIZZLE_128B only needs 128 B alignment, the
// swizzle follows absolute address bits - tools/umma_probe4.cu), 7 slots + 1 mirror slot
constexpr int kBBoxRows = 2;
constexpr int kBBoxPx = kBBoxRows * kPitch;           // 132
constexpr int kBSlotBytes = kBBoxPx * kC * 2;         // 16896
constexpr int kBSlots = 7;
constexpr int kBRingBytes = (kBSlots + 1) * kBSlotBytes;
constexpr int kBodyDynBytes = kBodyWBytes + kBRingBytes + 1024;
constexpr int kConstVecFloats = 15872;   // 62 KB of __constant__ for biases + slopes

__device__ __constant__ float c_vec[kConstVecFloats];

enum BodyBuf : int { kBufF0 = 0, kBufX0 = 1, kBufX1 = 2, kBufH = 3, kBufO = 4, kBufG0 = 5 };  // G0.. = group outputs
constexpr int kBodyMaxBufs = 5 + 16;

struct BodyMaps {
  CUtensorMap act[kBodyMaxBufs];   // one per activation buffer, all [B][H][W][64] bf16
  CUtensorMap w;                   // the whole packed blob as rows of 128 B
};

struct BodyParams {
  int B, H, W;
  int G, Bk, R;                    // groups, blocks per group, SE hidden width
  int n_layers;                    // G * (2 Bk + 1) + 1
  int tiles_per_seg, total_tiles, tiles_per_cta;
  float res_scale, inv_hw;
  bf16* buf[kBodyMaxBufs];         // activation buffers (same order as BodyMaps::act)
  const uint8_t* packed;           // packed weight blob (fc matrices are read from here)
  int64_t k_rcab0, k_rcab_stride, k_rcab_w2, k_rcab_fc0, k_rcab_fc2;   // byte offsets in the blob
  int64_t k_gconv0, k_gconv_stride, k_after;
  int cv_rcab0, cv_gconv0, cv_after;   // float offsets in c_vec: per RCAB [b1 64][slope 64][b2 64]; per plain conv [b 64]
  float* sums;                     // [n_rcab][B][64]
  float* se_out;                   // [B][n_rcab][64] or nullptr
  int* flags;                      // [gridDim.x], zeroed before launch
  long long* dbg;
};

struct BodyLayer {
  int fused;        // producer transforms x' = x + s*o instead of a TMA load
  int epi;          // kEpiPrelu / kEpiSum / kEpiResidual
  int in;           // input buffer (plain) or x buffer (fused)
  int xout;         // fused: buffer receiving x' (-1: do not store)
  int res;          // residual buffer (kEpiResidual)
  int out;          // output buffer
  int rcab_in;      // fused: RCAB index whose SE vector scales o
  int rcab_out;     // kEpiSum: RCAB index receiving the channel sums
  int w_row;        // first row (128 B units) of this layer's weights in the blob
  int cv_bias, cv_slope;   // offsets in c_vec
};

__device__ __forceinline__ BodyLayer body_layer(const BodyParams& p, int L) {
  BodyLayer l;
  const int per_group = 2 * p.Bk + 1;
  const int g = L / per_group, r = L - g * per_group;
  l.fused = 0; l.xout = -1; l.res = -1; l.rcab_in = -1; l.rcab_out = -1; l.cv_slope = 0;
  if (g == p.G) {                                  // conv_after_body + long skip -> X0
    l.epi = kEpiResidual; l.in = kBufG0 + p.G - 1; l.res = kBufF0; l.out = kBufX0;
    l.w_row = int(p.k_after >> 7); l.cv_bias = p.cv_after;
    return l;
  }
  const int gin = (g == 0) ? kBufF0 : kBufG0 + g - 1;
  if (r == 2 * p.Bk) {                             // group conv: input x' of the last block, + group input
    l.fused = 1; l.epi = kEpiResidual;
    l.in = (p.Bk == 1) ? gin : kBufX0 + ((p.Bk - 2) & 1);
    l.rcab_in = g * p.Bk + p.Bk - 1;
    l.res = gin; l.out = kBufG0 + g;
    l.w_row = int((p.k_gconv0 + g * p.k_gconv_stride) >> 7); l.cv_bias = p.cv_gconv0 + g * 64;
    return l;
  }
  const int b = r >> 1, rc = g * p.Bk + b;
  const int64_t rec = p.k_rcab0 + int64_t(rc) * p.k_rcab_stride;
  if ((r & 1) == 0) {                              // conv1 (+ PReLU) -> H
    l.epi = kEpiPrelu; l.out = kBufH;
    if (b == 0) {
      l.in = gin;
    } else {                                       // x' = X_{b-1} + s_{b-1} * o_{b-1}, stored to X[(b-1)&1]
      l.fused = 1;
      l.in = (b == 1) ? gin : kBufX0 + ((b - 2) & 1);
      l.xout = kBufX0 + ((b - 1) & 1);
      l.rcab_in = rc - 1;
    }
    l.w_row = int(rec >> 7); l.cv_bias = p.cv_rcab0 + rc * 192; l.cv_slope = l.cv_bias + 64;
  } else {                                         // conv2 -> O, channel sums
    l.epi = kEpiSum; l.in = kBufH; l.out = kBufO; l.rcab_out = rc;
    l.w_row = int((rec + p.k_rcab_w2) >> 7); l.cv_bias = p.cv_rcab0 + rc * 192 + 128;
  }
  return l;
}

struct BUnit { int n, t0, t1, ra, nboxes; };
__device__ __forceinline__ BUnit body_unit(const BodyParams& p, int g, int g_end) {
  BUnit u;
  u.n = g / p.tiles_per_seg;
  u.t0 = g - u.n * p.tiles_per_seg;
  u.t1 = min(p.tiles_per_seg, u.t0 + (g_end - g));
  u.ra = (kTileM * u.t0) / kPitch;                              // first staged row (row 0 = image row -1)
  const int rb = min((kTileM * u.t1 + kMaxShift - 1) / kPitch, p.H + 1);
  u.nboxes = (rb - u.ra) / kBBoxRows + 1;
  return u;
}

__device__ __forceinline__ int ld_acquire_gpu(const int* p) {
  int v;
  asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_gpu(int* p, int v) {
  asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint4 ld_cg_128(const void* p) {   // L2-coherent load (data written by other SMs)
  uint4 v;
  asm volatile("ld.global.cg.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
  return v;
}
// L2 eviction hints (same encodings CUTLASS uses for TMA cache hints): data read for the last time
// (x, o, h, residuals) is marked evict-first so the LIVE tensors of an RCAB (67 MB at batch 64)
// stay resident in the 126 MB L2 instead of being pushed out by dead ones.
constexpr uint64_t kPolicyEvictFirst = 0x12F0000000000000ull;
constexpr uint64_t kPolicyEvictLast = 0x14F0000000000000ull;
__device__ __forceinline__ uint4 ld_cg_128_hint(const void* p, uint64_t policy) {
  uint4 v;
  asm volatile("ld.global.cg.L2::cache_hint.v4.u32 {%0, %1, %2, %3}, [%4], %5;"
               : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p), "l"(policy));
  return v;
}
__device__ __forceinline__ void tma_load_4d_hint(const CUtensorMap* m, uint64_t* bar, uint32_t dst_smem, int c0, int c1,
                                                 int c2, int c3, uint64_t policy) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint"
      " [%0], [%1, {%3, %4, %5, %6}], [%2], %7;"
      :
      : "r"(dst_smem), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3),
        "l"(policy)
      : "memory");
}
__device__ __forceinline__ float ld_cg_f32(const float* p) {
  float v;
  asm volatile("ld.global.cg.f32 %0, [%1];" : "=f"(v) : "l"(p));
  return v;
}
__device__ __forceinline__ void fence_proxy_async_all() { asm volatile("fence.proxy.async;" ::: "memory"); }
__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

__global__ void __launch_bounds__(kBodyThreads, 1)
body_umma_kernel(const __grid_constant__ BodyMaps maps, const BodyParams p) {
  constexpr int N = kC;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* w_smem = smem;
  uint8_t* ring = smem + kBodyWBytes;
  __shared__ uint64_t bar_w[9], bar_wfree[9], bar_x[kBSlots], bar_full[kBSlots], bar_empty[kBSlots];
  __shared__ uint64_t bar_acc_full[kBodyAccBufs], bar_acc_empty[kBodyAccBufs], bar_done;
  __shared__ uint32_t tmem_slot;
  __shared__ float s_mean[2][kC], s_hid[2][kC], s_scale[2][kC];

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  constexpr uint32_t kTmemCols = kBodyAccBufs * N;
  constexpr int kFrontThreads = 32 * kBodyFirstMmaWarp;   // TMA warp + transform warps (named barrier 1)

  const int g_begin = blockIdx.x * p.tiles_per_cta;
  const int g_end = min(p.total_tiles, g_begin + p.tiles_per_cta);
  const int n_tiles = g_end - g_begin;
  // peers: CTAs owning tiles of the images this CTA touches (including itself)
  const int img0 = g_begin / p.tiles_per_seg, img1 = (g_end - 1) / p.tiles_per_seg;
  const int peer0 = (img0 * p.tiles_per_seg) / p.tiles_per_cta;
  const int peer1 = min(int(gridDim.x) - 1, ((img1 + 1) * p.tiles_per_seg - 1) / p.tiles_per_cta);

  if (warp == kBodyFirstMmaWarp) tmem_alloc(&tmem_slot, kTmemCols);
  if (tid == 0) {
    for (int i = 0; i < 9; ++i) { mbar_init(&bar_w[i], 1); mbar_init(&bar_wfree[i], kBodyMmaWarps); }
    for (int i = 0; i < kBSlots; ++i) {
      mbar_init(&bar_x[i], 1);
      mbar_init(&bar_full[i], kBodyXformWarps + 1);     // 8 transform warps + the TMA issuer, every box
      mbar_init(&bar_empty[i], kBodyMmaWarps);
    }
    for (int i = 0; i < kBodyAccBufs; ++i) { mbar_init(&bar_acc_full[i], 1); mbar_init(&bar_acc_empty[i], kBodyEpiWarps); }
    mbar_init(&bar_done, kBodyXformWarps + kBodyEpiWarps);
    fence_mbar_init();
    tma_prefetch_desc(&maps.w);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_slot;
  if (n_tiles <= 0) return;   // never happens with the host's grid sizing (all CTAs have tiles)

  if (warp == 0) {
    // ============================================================ TMA issuer (one lane)
    uint32_t gb = 0;   // running box counter
    for (int L = 0; L < p.n_layers; ++L) {
      const BodyLayer ly = body_layer(p, L);
      if (lane == 0) {
        // weights of this layer, tap by tap, as soon as the previous layer released the tap
        for (int tap = 0; tap < 9; ++tap) {
          if (L > 0) mbar_wait(&bar_wfree[tap], (L - 1) & 1);
          mbar_expect_tx(&bar_w[tap], N * kC * 2);
          tma_load_2d(&maps.w, &bar_w[tap], w_smem + tap * N * 128, 0, ly.w_row + tap * N);
        }
      }
      __syncwarp();
      if (L > 0) named_bar_sync(1, kFrontThreads);       // peers have finished layer L-1 (polled by warp 1)
      if (ly.fused) named_bar_sync(1, kFrontThreads);    // keeps the barrier sequence of the transform warps (SE)
      if (lane == 0) {
        const uint64_t pol = (ly.fused || ly.in == kBufH || ly.epi == kEpiResidual) ? kPolicyEvictFirst
                                                                                   : 0x1000000000000000ull;
        for (int g = g_begin; g < g_end;) {
          const BUnit u = body_unit(p, g, g_end);
          for (int j = 0; j < u.nboxes; ++j, ++gb) {
            const uint32_t slot = gb % kBSlots, ph = (gb / kBSlots) & 1;
            const int y0 = u.ra - 1 + j * kBBoxRows;
            mbar_wait(&bar_empty[slot], ph ^ 1);
            const uint32_t dst = smem_u32(ring + slot * kBSlotBytes);
            if (ly.fused) {
              // x goes to the slot; the transform warps finish the box (and the mirror copy)
              mbar_expect_tx(&bar_x[slot], kBSlotBytes);
              tma_load_4d_hint(&maps.act[ly.in], &bar_x[slot], dst, 0, -1, y0, u.n, pol);
              mbar_arrive(&bar_full[slot]);
            } else {
              const bool mirror = (slot == 0) && (j > 0);
              mbar_expect_tx(&bar_full[slot], mirror ? 2 * kBSlotBytes : kBSlotBytes);
              tma_load_4d_hint(&maps.act[ly.in], &bar_full[slot], dst, 0, -1, y0, u.n, pol);
              if (mirror)
                tma_load_4d_hint(&maps.act[ly.in], &bar_full[slot], smem_u32(ring + kBSlots * kBSlotBytes), 0, -1, y0,
                                 u.n, pol);
            }
          }
          g += u.t1 - u.t0;
        }
      }
      __syncwarp();
    }
  } else if (warp < kBodyFirstMmaWarp) {
    // ============================================================ transform warps (256 threads)
    constexpr int kXT = 32 * kBodyXformWarps;
    const int xt = tid - 32 * kBodyFirstXformWarp;      // 0..255
    const int chunk = xt & 7;                           // 16-byte channel chunk this thread always handles
    uint32_t gb = 0;                                    // running box counter (same sequence as the issuer)
    uint32_t x_phase = 0;                               // bar_x completes only in fused layers: one parity bit per slot
    long long d_flag = 0, d_se = 0, d_fused = 0, d_plain = 0, d_t = BDBG ? clock64() : 0;
    const long long d_start = d_t;
#define DBG_LAP(acc) if (BDBG) { const long long n_ = clock64(); acc += n_ - d_t; d_t = n_; }
    for (int L = 0; L < p.n_layers; ++L) {
      const BodyLayer ly = body_layer(p, L);
      DBG_LAP(d_plain)
      // ---- wait until every peer finished layer L-1 (their outputs are my inputs / halos, and my
      //      outputs of this layer overwrite buffers they were still reading in L-1)
      if (L > 0) {
        if (warp == kBodyFirstXformWarp) {
          for (int k = peer0 + lane; k <= peer1; k += 32)
            while (ld_acquire_gpu(p.flags + k) < L) { __nanosleep(32); }
          __syncwarp();
          fence_proxy_async_all();
        }
        named_bar_sync(1, kFrontThreads);
      }
      DBG_LAP(d_flag)
      // ---- SE vectors of the (at most two) images of this CTA
      if (ly.fused) {
        const float* sums = p.sums + size_t(ly.rcab_in) * p.B * kC;
        const uint8_t* rec = p.packed + p.k_rcab0 + int64_t(ly.rcab_in) * p.k_rcab_stride;
        const float* fc0 = reinterpret_cast<const float*>(rec + p.k_rcab_fc0);
        const float* fc2 = reinterpret_cast<const float*>(rec + p.k_rcab_fc2);
        const int u = (xt >> 6) & 1, c = xt & 63;         // unit (image) slot, channel; threads >= 128 idle here
        const int n = min(img0 + u, img1);
        if (xt < 128) s_mean[u][c] = ld_cg_f32(sums + size_t(n) * kC + c) * p.inv_hw;
        named_bar_sync(2, kXT);
        if (xt < 128 && c < p.R) {
          float a = 0.f;
          for (int k = 0; k < kC; ++k) a = fmaf(__ldg(fc0 + c * kC + k), s_mean[u][k], a);
          s_hid[u][c] = fmaxf(a, 0.f);
        }
        named_bar_sync(2, kXT);
        if (xt < 128) {
          float a = 0.f;
          for (int j = 0; j < p.R; ++j) a = fmaf(__ldg(fc2 + c * p.R + j), s_hid[u][j], a);
          const float s = 1.f / (1.f + expf(-a));
          s_scale[u][c] = s * p.res_scale;
          // the CTA owning tile 0 of the image publishes the attention vector
          if (p.se_out && (img0 + u <= img1) && (n * p.tiles_per_seg >= g_begin) && (n * p.tiles_per_seg < g_end))
            p.se_out[(size_t(n) * (p.G * p.Bk) + ly.rcab_in) * kC + c] = s;
        }
        named_bar_sync(1, kFrontThreads);               // also releases the TMA issuer into the box loop
      }
      DBG_LAP(d_se)
      // ---- boxes
      const bf16* oin = p.buf[kBufO];
      bf16* xout = ly.xout >= 0 ? p.buf[ly.xout] : nullptr;
      constexpr int kSteps = (kBBoxPx * 8 + kXT - 1) / kXT;            // 5 (last one: 32 threads)
      for (int g = g_begin; g < g_end;) {
        const BUnit u = body_unit(p, g, g_end);
        if (!ly.fused) {
          // every transform warp arrives once per box so the full barrier always counts 9
          for (int j = 0; j < u.nboxes; ++j, ++gb) {
            const uint32_t slot = gb % kBSlots, ph = (gb / kBSlots) & 1;
            mbar_wait(&bar_empty[slot], ph ^ 1);
            __syncwarp();
            if (lane == 0) mbar_arrive(&bar_full[slot]);
          }
        } else {
          const int us = u.n - img0;                                     // which s_scale row
          float sc[8];
#pragma unroll
          for (int e = 0; e < 8; ++e) sc[e] = s_scale[us][chunk * 8 + e];
          const int lo = kTileM * u.t0, hi = kTileM * u.t1;              // owned strip-linear range
          // o of a box is loaded one box ahead of its use
          uint4 ov[kSteps];
          auto load_o = [&](int j, uint4 (&dst)[kSteps]) {
            const int y0 = u.ra - 1 + j * kBBoxRows;
#pragma unroll
            for (int q = 0; q < kSteps; ++q) {
              const int px = (q * kXT + xt) >> 3;
              const int row = px / kPitch, col = px - row * kPitch;
              const int iy = y0 + row, ix = col - 1;
              const bool ok = (px < kBBoxPx) && iy >= 0 && iy < p.H && ix >= 0 && ix < p.W;
              dst[q] = ok ? ld_cg_128_hint(oin + ((size_t(u.n) * p.H + iy) * p.W + ix) * kC + chunk * 8, kPolicyEvictFirst)
                          : make_uint4(0, 0, 0, 0);
            }
          };
          load_o(0, ov);
          for (int j = 0; j < u.nboxes; ++j, ++gb) {
            const uint32_t slot = gb % kBSlots;
            const int y0 = u.ra - 1 + j * kBBoxRows;
            uint4 on[kSteps];
            if (j + 1 < u.nboxes) load_o(j + 1, on);
            mbar_wait(&bar_x[slot], (x_phase >> slot) & 1);              // x has landed in the slot
            x_phase ^= 1u << slot;
            uint8_t* dst = ring + slot * kBSlotBytes;
            uint8_t* dst_mirror = (slot == 0 && j > 0) ? ring + kBSlots * kBSlotBytes : nullptr;
#pragma unroll
            for (int q = 0; q < kSteps; ++q) {
              const int px = (q * kXT + xt) >> 3;
              if (px < kBBoxPx) {
                // SWIZZLE_128B: 16-byte chunk index XOR (128-byte row index of the absolute address mod 8)
                const uint32_t lin_b = uint32_t(slot) * kBSlotBytes + uint32_t(px) * 128u;   // ring is 1024-aligned
                const uint32_t so = uint32_t(px) * 128u + (uint32_t(chunk ^ ((lin_b >> 7) & 7)) << 4);
                const uint4 xv = *reinterpret_cast<const uint4*>(dst + so);
                uint4 r;
                r.x = pack_bf16(fmaf(bf16lo(ov[q].x), sc[0], bf16lo(xv.x)), fmaf(bf16hi(ov[q].x), sc[1], bf16hi(xv.x)));
                r.y = pack_bf16(fmaf(bf16lo(ov[q].y), sc[2], bf16lo(xv.y)), fmaf(bf16hi(ov[q].y), sc[3], bf16hi(xv.y)));
                r.z = pack_bf16(fmaf(bf16lo(ov[q].z), sc[4], bf16lo(xv.z)), fmaf(bf16hi(ov[q].z), sc[5], bf16hi(xv.z)));
                r.w = pack_bf16(fmaf(bf16lo(ov[q].w), sc[6], bf16lo(xv.w)), fmaf(bf16hi(ov[q].w), sc[7], bf16hi(xv.w)));
                *reinterpret_cast<uint4*>(dst + so) = r;
                if (dst_mirror) {
                  const uint32_t lin_m = uint32_t(kBSlots) * kBSlotBytes + uint32_t(px) * 128u;
                  *reinterpret_cast<uint4*>(dst_mirror + uint32_t(px) * 128u + (uint32_t(chunk ^ ((lin_m >> 7) & 7)) << 4)) = r;
                }
                if (xout) {
                  const int row = px / kPitch, col = px - row * kPitch;
                  const int iy = y0 + row, ix = col - 1;
                  const int lin = iy * kPitch + ix;                      // strip-linear output index of the pixel
                  if (ix >= 0 && ix < p.W && lin >= lo && lin < hi)
                    *reinterpret_cast<uint4*>(xout + ((size_t(u.n) * p.H + iy) * p.W + ix) * kC + chunk * 8) = r;
                }
              }
            }
            fence_proxy_async_smem();          // generic-proxy smem writes -> visible to tcgen05.mma
            __syncwarp();
            if (lane == 0) mbar_arrive(&bar_full[slot]);
#pragma unroll
            for (int q = 0; q < kSteps; ++q) ov[q] = on[q];
          }
        }
        g += u.t1 - u.t0;
      }
      if (ly.fused) { DBG_LAP(d_fused) } else { DBG_LAP(d_plain) }
      // ---- this warp's global writes (x') are complete: fence, then count the warp as done
      __threadfence();
      __syncwarp();
      if (lane == 0) mbar_arrive(&bar_done);
    }
    if (BDBG && xt == 0) {
      long long* d = p.dbg + blockIdx.x * 16;
      d[0] = 0; d[1] = d_flag; d[2] = d_se; d[3] = d_fused; d[4] = d_plain; d[5] = clock64() - d_start;
    }
  } else if (warp < kBodyFirstEpiWarp) {
    // ============================================================ MMA issuers (2 warps)
    constexpr uint32_t idesc = umma_idesc_bf16(kTileM, N);
    constexpr uint32_t kDescHi = (1024u >> 4) | (1u << 14) | (2u << 29);
    constexpr uint32_t kLbo = 1u << 16;
    const uint32_t ring_lo = (smem_u32(ring) >> 4) | kLbo;
    const uint32_t w_lo = (smem_u32(w_smem) >> 4) | kLbo;
    const bool leader = elect_one();
    const uint32_t my_parity = warp - kBodyFirstMmaWarp;
    uint32_t gb_base = 0, tile_ctr = 0;
    long long m_acc = 0, m_full = 0, m_issue = 0, m_t = 0, m_fl = 0, m_pl = 0, m_ffull = 0;
    const long long m_start = BDBG ? clock64() : 0;
    for (int L = 0; L < p.n_layers; ++L) {
      const long long m_l0 = BDBG ? clock64() : 0;
      const long long m_full0 = m_full;
      const bool m_is_fused = body_layer(p, L).fused != 0;
      // index (within the layer) of this warp's first / last tile
      const int first_mine = ((tile_ctr & 1) == my_parity) ? 0 : 1;
      const int last_mine = (((tile_ctr + n_tiles - 1) & 1) == my_parity) ? n_tiles - 1 : n_tiles - 2;
      if (last_mine < first_mine) {            // no tile in this layer: release the weight taps right away
        if (leader)
          for (int tap = 0; tap < 9; ++tap) mbar_arrive(&bar_wfree[tap]);
      }
      int i_layer = 0;
      for (int g = g_begin; g < g_end;) {
        const BUnit u = body_unit(p, g, g_end);
        int waited = 0, released = 0;
        for (int t = u.t0; t < u.t1; ++t, ++tile_ctr, ++i_layer) {
          const uint32_t acc = tile_ctr & (kBodyAccBufs - 1);
          const bool mine = (tile_ctr & 1) == my_parity;
          if (BDBG) m_t = clock64();
          if (mine) mbar_wait(&bar_acc_empty[acc], ((tile_ctr / kBodyAccBufs) & 1) ^ 1);
          if (BDBG) { const long long n_ = clock64(); m_acc += n_ - m_t; m_t = n_; }
          const int base = kTileM * t - kPitch * u.ra;
          const int need_last = min((base + kTileM + kMaxShift - 1) / kBBoxPx, u.nboxes - 1);
          while (mine && waited <= need_last) {
            const uint32_t gb = gb_base + waited;
            mbar_wait(&bar_full[gb % kBSlots], (gb / kBSlots) & 1);
            ++waited;
          }
          if (BDBG) { const long long n_ = clock64(); m_full += n_ - m_t; m_t = n_; }
          tc_fence_after();
          const uint32_t d_tmem = tmem_base + acc * N;
          // a view starts in box lb0, lb0 + 1 or lb0 + 2 (tap offsets reach 134 px, a box is 132)
          const int lb0 = base / kBBoxPx, r0 = base - lb0 * kBBoxPx;
          const uint32_t slot0 = (gb_base + lb0) % kBSlots;
          const uint32_t slot1 = (slot0 + 1 >= kBSlots) ? slot0 + 1 - kBSlots : slot0 + 1;
          const uint32_t slot2 = (slot0 + 2 >= kBSlots) ? slot0 + 2 - kBSlots : slot0 + 2;
          const uint32_t a0 = ring_lo + slot0 * (kBSlotBytes >> 4) + r0 * 8;
          const uint32_t a1 = ring_lo + slot1 * (kBSlotBytes >> 4) + (r0 - kBBoxPx) * 8;
          const uint32_t a2 = ring_lo + slot2 * (kBSlotBytes >> 4) + (r0 - 2 * kBBoxPx) * 8;
          const bool w_first = (i_layer == first_mine), w_last = (i_layer == last_mine);
          if (leader && mine) {
#pragma unroll
            for (int tap = 0; tap < 9; ++tap) {
              if (w_first) mbar_wait(&bar_w[tap], L & 1);
              const int off = (tap / 3) * kPitch + (tap % 3);
              const uint32_t a_lo = ((r0 + off < kBBoxPx) ? a0 : (r0 + off < 2 * kBBoxPx) ? a1 : a2) + off * 8;
              const uint32_t b_lo = w_lo + tap * (N * 128 >> 4);
#pragma unroll
              for (int k = 0; k < 4; ++k)
                umma_bf16_ss_lohi(d_tmem, a_lo + 2 * k, b_lo + 2 * k, kDescHi, idesc, (tap | k) != 0);
              if (w_last) umma_commit(&bar_wfree[tap]);   // next layer's tap may overwrite once these MMAs finish
            }
          }
          __syncwarp();
          if (BDBG) { const long long n_ = clock64(); m_issue += n_ - m_t; m_t = n_; }
          const int next_first = (t + 1 < u.t1) ? (base + kTileM) / kBBoxPx : u.nboxes;
          while (released < next_first) {
            if (leader) umma_commit(&bar_empty[(gb_base + released) % kBSlots]);
            ++released;
          }
          if (leader && mine) umma_commit(&bar_acc_full[acc]);
          __syncwarp();
        }
        gb_base += u.nboxes;
        g += u.t1 - u.t0;
      }
      if (BDBG) { if (m_is_fused) { m_fl += clock64() - m_l0; m_ffull += m_full - m_full0; } else m_pl += clock64() - m_l0; }
    }
    if (BDBG && leader && warp == kBodyFirstMmaWarp) {
      long long* d = p.dbg + blockIdx.x * 16;
      d[13] = m_fl; d[14] = m_pl; d[15] = m_ffull;
      d[6] = m_acc; d[7] = m_full; d[8] = m_issue; d[9] = clock64() - m_start;
    }
  } else {
    // ============================================================ epilogue (8 warps)
    constexpr int CW = 32;
    const int q = warp & 3;
    const int half = (warp - kBodyFirstEpiWarp) >> 2;
    const int col0 = half * CW;
    const int row_in_tile = q * 32 + lane;
    const bool flag_writer = (warp == kBodyFirstEpiWarp);
    uint32_t tile_ctr = 0;
    long long e_wait = 0, e_done = 0, e_t = 0;
    const long long e_start = BDBG ? clock64() : 0;
    for (int L = 0; L < p.n_layers; ++L) {
      const BodyLayer ly = body_layer(p, L);
      bf16* outp = p.buf[ly.out];
      const bf16* resp = ly.res >= 0 ? p.buf[ly.res] : nullptr;
      const float* cbias = c_vec + ly.cv_bias + col0;
      const float* cslope = c_vec + ly.cv_slope + col0;
      for (int g = g_begin; g < g_end;) {
        const BUnit u = body_unit(p, g, g_end);
        float csum[CW];
#pragma unroll
        for (int c = 0; c < CW; ++c) csum[c] = 0.f;
        for (int t = u.t0; t < u.t1; ++t, ++tile_ctr) {
          const uint32_t acc = tile_ctr & (kBodyAccBufs - 1);
          if (BDBG) e_t = clock64();
          mbar_wait(&bar_acc_full[acc], (tile_ctr / kBodyAccBufs) & 1);
          if (BDBG) e_wait += clock64() - e_t;
          tc_fence_after();
          const int lin = kTileM * t + row_in_tile;
          const int y = lin / kPitch, x = lin - y * kPitch;
          const bool valid = (x < kStripW) && (y < p.H);
          const size_t opix = (size_t(u.n) * p.H + y) * p.W + x;
#pragma unroll
          for (int hp = 0; hp < 2; ++hp) {      // two passes of 16 columns keep the live register set small
            uint32_t v[16];
            tmem_ld_32x16(tmem_base + acc * N + col0 + 16 * hp + (uint32_t(q * 32) << 16), v);
            tmem_ld_wait();
            if (hp == 1) {                      // accumulator fully read: hand it back to the MMA issuers
              tc_fence_before();
              __syncwarp();
              if (lane == 0) mbar_arrive(&bar_acc_empty[acc]);
            }
            float f[16];
#pragma unroll
            for (int c = 0; c < 16; ++c) f[c] = __uint_as_float(v[c]) + cbias[16 * hp + c];
            if (ly.epi == kEpiSum) {
              if (valid) {
#pragma unroll
                for (int c = 0; c < 16; ++c) csum[16 * hp + c] += f[c];
              }
            } else if (ly.epi == kEpiPrelu) {
#pragma unroll
              for (int c = 0; c < 16; ++c) f[c] = f[c] > 0.f ? f[c] : f[c] * cslope[16 * hp + c];
            }
            if (valid) {
              if (ly.epi == kEpiResidual) {
                const bf16* rsd = resp + opix * kC + col0 + 16 * hp;
#pragma unroll
                for (int j = 0; j < 2; ++j) {
                  const uint4 r = ld_cg_128(rsd + 8 * j);
                  f[8 * j + 0] += bf16lo(r.x); f[8 * j + 1] += bf16hi(r.x);
                  f[8 * j + 2] += bf16lo(r.y); f[8 * j + 3] += bf16hi(r.y);
                  f[8 * j + 4] += bf16lo(r.z); f[8 * j + 5] += bf16hi(r.z);
                  f[8 * j + 6] += bf16lo(r.w); f[8 * j + 7] += bf16hi(r.w);
                }
              }
              uint32_t o[8];
#pragma unroll
              for (int e = 0; e < 8; ++e) o[e] = pack_bf16(f[2 * e], f[2 * e + 1]);
              st_global_256(outp + opix * kC + col0 + 16 * hp, o);
            }
          }
        }
        if (ly.epi == kEpiSum) {
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
          atomicAdd(p.sums + (size_t(ly.rcab_out) * p.B + u.n) * kC + col0 + lane, csum[0]);
        }
        g += u.t1 - u.t0;
      }
      // ---- layer done for this warp: make its global writes visible, then publish the CTA's flag
      __threadfence();
      __syncwarp();
      if (lane == 0) mbar_arrive(&bar_done);
      if (flag_writer) {
        if (BDBG) e_t = clock64();
        mbar_wait(&bar_done, L & 1);
        if (BDBG) e_done += clock64() - e_t;
        if (lane == 0) {
          fence_proxy_async_all();
          __threadfence();
          st_release_gpu(p.flags + blockIdx.x, L + 1);
        }
        __syncwarp();
      }
    }
    if (BDBG && flag_writer && lane == 0) {
      long long* d = p.dbg + blockIdx.x * 16;
      d[10] = e_wait; d[11] = e_done; d[12] = clock64() - e_start;
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == kBodyFirstMmaWarp) tmem_dealloc(tmem_base, kTmemCols);
}

}  // namespace fen
