// Persistent "body" kernel: ALL 64->64 3x3 convolutions of the residual body
// (num_groups x (2 x blocks_per_group + 1) + conv_after_body = 127 layers for the 6 x 10 model)
// in ONE launch, with the squeeze-and-excitation scale and the RCAB residual fused into the conv2
// epilogue.  Reference: src/models/custom.py:167-175 (body loop, long skip),
// src/models/blocks.py:135-153 (RCAB), :75-92 (ChannelAttention), :185-189 (ResidualGroup).
//
// Every CTA owns the same run of output tiles in every layer (all body layers share one geometry),
// so a layer boundary is not a grid-wide barrier: a CTA only waits for the CTAs that share an image
// with it ("peers": halo rows + the per-image SE pool) through release/acquire flags in global
// memory.  Launched cooperatively with one CTA per SM, so all CTAs are co-resident.
//
// SE without a second pass.  The reference computes s = sigmoid(W2' relu(W0' mean_hw(o))) from the
// conv2 OUTPUT o and then x' = o*s*0.2 + x, which would force o through memory.  But mean_hw(o) is
// linear in conv2's INPUT h:  mean(o)[c] = b2[c] + 1/HW * sum_{tap,ci} W2[c,ci,tap] * S_tap[ci], where
// S_tap[ci] is the sum of h[ci] over the window the tap sees = (total) - (excluded border row)
// - (excluded border column) + (corner).  The conv1 epilogue accumulates those 9 sums of the
// bf16-rounded h per image (the total; border rows / columns / corners are re-read from h, a few
// KB per image); at the start of the conv2 layer the epilogue warps turn them into s (a 64x576
// mat-vec + the two tiny FC layers, fp32) while the first tiles' MMAs already run into the 8
// TMEM accumulator buffers; the conv2 epilogue then writes
// x' = x + (0.2 s[c]) * (acc + b2[c]) directly.  o is never materialised and no elementwise pass or
// transform producer exists: every layer is a plain TMA-fed convolution.
//
// Per layer the tile pipeline is the one of conv3x3_umma.cuh (ring of boxes with a mirror slot -
// here 2-row boxes, 7 slots -, two tcgen05.mma issuer warps, 8 epilogue warps, TMEM accumulators).
// The epilogue keeps bias, PReLU slope and SE scale of its 32 columns in registers (loaded once per
// layer from __constant__ memory): it must stay off shared memory, whose bandwidth the tensor core
// needs, and indexed constant loads per tile starve its two warps per scheduler.  The weights of the
// next layer are loaded tap by tap as soon as the current layer has issued its last MMA on that tap.
#pragma once
#include "conv3x3_umma.cuh"

namespace fen {

#ifndef FEN_SE_GATE
#define FEN_SE_GATE 0   // 1: MMA issuers wait for the SE vector in conv2 layers (measured slower: 4.57 vs 4.35 ms)
#endif
#ifndef FEN_BODY_DEBUG
#define FEN_BODY_DEBUG 0   // 1: per-CTA cycle counters into BodyParams::dbg (developer builds)
#endif
#define BDBG (FEN_BODY_DEBUG && p.dbg)
// timeline trace (FEN_BODY_DEBUG=2): event e of layer L of CTA 70 -> dbg[4096 + L*16 + e]
// per-tile trace (FEN_BODY_DEBUG=2): layers 20..23 of CTA 70 -> dbg[8192 + (L-20)*256 + e*32 + i]
#define BT2(L, e, i) do { if (FEN_BODY_DEBUG == 2 && p.dbg && blockIdx.x == 70 && (L) >= 20 && (L) < 24 && (i) < 32) p.dbg[8192 + ((L) - 20) * 512 + (e) * 32 + (i)] = clock64(); } while (0)
#define BTRACE(L, e) do { if (FEN_BODY_DEBUG == 2 && p.dbg && blockIdx.x == 70) p.dbg[4096 + (L) * 16 + (e)] = clock64(); } while (0)

constexpr int kBodyMmaWarps = 2;
constexpr int kBodyEpiWarps = 8;
constexpr int kBodyFirstMmaWarp = 1;                  // warp 0: TMA issuer + peer-flag poller
constexpr int kBodyFirstEpiWarp = kBodyFirstMmaWarp + kBodyMmaWarps;
constexpr int kBodyThreads = 32 * (kBodyFirstEpiWarp + kBodyEpiWarps);   // 352 -> up to 186 registers per thread
constexpr int kBodyMaxUnits = 4;                      // images a CTA may touch in one layer (host caps the batch per launch)
constexpr int kBodyAccBufs = 8;                        // 8 x 64 = all 512 TMEM columns: the MMAs can run 8 tiles ahead of the SE vector
constexpr int kBodyWBytes = 9 * kC * kC * 2;
// activation ring: 2-row boxes (132 px, 16 896 B; TMA SWIZZLE_128B only needs 128 B alignment, the
// swizzle follows absolute address bits - tools/umma_probe4.cu), 7 slots + 1 mirror slot
constexpr int kBBoxRows = 2;
constexpr int kBBoxPx = kBBoxRows * kPitch;           // 132
constexpr int kBSlotBytes = kBBoxPx * kC * 2;         // 16896
constexpr int kBSlots = 7;
constexpr int kBRingBytes = (kBSlots + 1) * kBSlotBytes;
constexpr int kBodyDynBytes = kBodyWBytes + kBRingBytes + 1024;
#ifdef FEN_DEV
constexpr int kConstVecFloats = 15872;   // 62 KB of __constant__ for biases + slopes (first-generation kernel only)
__device__ __constant__ float c_vec[kConstVecFloats];
#endif

enum BodyBuf : int { kBufF0 = 0, kBufX0 = 1, kBufX1 = 2, kBufH = 3, kBufO = 4 /* unused */, kBufG0 = 5 };  // G0.. = group outputs
enum BodyEpi : int { kBEpiPreluHsum = 0, kBEpiSeResidual = 1, kBEpiResidual = 2 };
enum HSum : int { kHsTotal = 0, kHsRow0, kHsRowL, kHsCol0, kHsColL, kHsC00, kHsC0L, kHsCL0, kHsCLL, kHsCount };
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
  float* hsum;                     // [n_rcab][B][64]: per-image channel sums of the bf16-rounded h
  float* se_out;                   // [B][n_rcab][64] or nullptr
  int* flags;                      // [gridDim.x], zeroed before launch
  long long* dbg;
};

struct BodyLayer {
  int epi;          // BodyEpi
  int in;           // input buffer
  int res;          // residual buffer (kBEpiSeResidual: x of the block, kBEpiResidual: skip) or -1
  int out;          // output buffer
  int rcab;         // RCAB index (h sums written by conv1, consumed by conv2) or -1
  int last_use;     // the input is dead after this layer (L2 evict-first hint)
  int w_row;        // first row (128 B units) of this layer's weights in the blob
  int64_t w_off;    // byte offset of those weights in the blob (conv2: read again for the SE mat-vec)
  int cv_bias, cv_slope;   // offsets in c_vec
};

__device__ __forceinline__ BodyLayer body_layer(const BodyParams& p, int L) {
  BodyLayer l;
  const int per_group = 2 * p.Bk + 1;
  const int g = L / per_group, r = L - g * per_group;
  l.res = -1; l.rcab = -1; l.cv_slope = 0; l.last_use = 1;
  if (g == p.G) {                                  // conv_after_body + long skip -> X0
    l.epi = kBEpiResidual; l.in = kBufG0 + p.G - 1; l.res = kBufF0; l.out = kBufX0;
    l.w_off = p.k_after; l.cv_bias = p.cv_after;
  } else {
    const int gin = (g == 0) ? kBufF0 : kBufG0 + g - 1;
    if (r == 2 * p.Bk) {                           // group conv on the last block's output, + group input
      l.epi = kBEpiResidual; l.in = kBufX0 + ((p.Bk - 1) & 1); l.res = gin; l.out = kBufG0 + g;
      l.w_off = p.k_gconv0 + g * p.k_gconv_stride; l.cv_bias = p.cv_gconv0 + g * 64;
    } else {
      const int b = r >> 1, rc = g * p.Bk + b;
      const int xb = (b == 0) ? gin : kBufX0 + ((b - 1) & 1);          // input of block b
      const int64_t rec = p.k_rcab0 + int64_t(rc) * p.k_rcab_stride;
      l.rcab = rc;
      if ((r & 1) == 0) {                          // conv1 + PReLU -> H, plus the 9 channel sums of h
        l.epi = kBEpiPreluHsum; l.in = xb; l.out = kBufH; l.last_use = 0;   // xb is read again as conv2's residual
        l.w_off = rec; l.cv_bias = p.cv_rcab0 + rc * 192; l.cv_slope = l.cv_bias + 64;
      } else {                                     // conv2, SE scale, residual -> X[b & 1]
        l.epi = kBEpiSeResidual; l.in = kBufH; l.res = xb; l.out = kBufX0 + (b & 1);
        l.w_off = rec + p.k_rcab_w2; l.cv_bias = p.cv_rcab0 + rc * 192 + 128;
      }
    }
  }
  l.w_row = int(l.w_off >> 7);
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
__device__ __forceinline__ void ld_cg_256_hint(const void* p, uint64_t policy, uint32_t (&v)[8]) {
  asm volatile("ld.global.cg.L2::cache_hint.v8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8], %9;"
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
               : "l"(p), "l"(policy));
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
__device__ __forceinline__ float4 ld_cg_f32x4(const float* p) {
  float4 v;
  asm volatile("ld.global.cg.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
  return v;
}
__device__ __forceinline__ void fence_proxy_async_all() { asm volatile("fence.proxy.async;" ::: "memory"); }
__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

#ifdef FEN_DEV   // first-generation persistent kernel: developer builds only (A/B runs), not in the shipped library
__global__ void __launch_bounds__(kBodyThreads, 1)
body_umma_kernel(const __grid_constant__ BodyMaps maps, const BodyParams p) {
  constexpr int N = kC;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* w_smem = smem;
  uint8_t* ring = smem + kBodyWBytes;
  __shared__ uint64_t bar_w[9], bar_wfree[9], bar_full[kBSlots], bar_empty[kBSlots];
  __shared__ uint64_t bar_acc_full[kBodyAccBufs], bar_acc_empty[kBodyAccBufs], bar_done, bar_flags, bar_se;
  __shared__ uint32_t tmem_slot;
  __shared__ __align__(16) float s_scale[kBodyMaxUnits][kC];   // res_scale * s per image of this CTA
  __shared__ __align__(16) float s_S[2][9][kC], s_q[2][kHsCount][kC], s_mean[2][kC], s_hid[2][kC];   // SE work arrays
  __shared__ float s_part[2][4][kC];

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  constexpr uint32_t kTmemCols = kBodyAccBufs * N;
  constexpr int kEpiThreads = 32 * kBodyEpiWarps;         // named barrier 1

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
    for (int i = 0; i < kBSlots; ++i) { mbar_init(&bar_full[i], 1); mbar_init(&bar_empty[i], kBodyMmaWarps); }
    for (int i = 0; i < kBodyAccBufs; ++i) { mbar_init(&bar_acc_full[i], 1); mbar_init(&bar_acc_empty[i], kBodyEpiWarps); }
    mbar_init(&bar_done, kBodyEpiWarps);
    mbar_init(&bar_flags, 1);
    mbar_init(&bar_se, 1);
    fence_mbar_init();
    tma_prefetch_desc(&maps.w);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_slot;
  if (n_tiles <= 0) return;   // never happens with the host's grid sizing (all CTAs have tiles)

  if (warp == 0) {
    // ============================================================ TMA issuer + peer-flag poller
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
      // wait until every peer finished layer L-1 (their outputs are my inputs / halos, and my outputs of
      // this layer overwrite buffers they were still reading in L-1)
      if (L > 0) {
        for (int k = peer0 + lane; k <= peer1; k += 32)
          while (ld_acquire_gpu(p.flags + k) < L) { __nanosleep(32); }
        __syncwarp();
        fence_proxy_async_all();
        if (lane == 0) mbar_arrive(&bar_flags);          // the epilogue may read peers' h / sums (SE vector)
      }
      if (lane == 0) BTRACE(L, 0);   // flags passed, first box about to be issued
      const uint32_t gb_l0 = gb;
      if (lane == 0) {
        const uint64_t pol = ly.last_use ? kPolicyEvictFirst : 0x1000000000000000ull;
        for (int g = g_begin; g < g_end;) {
          const BUnit u = body_unit(p, g, g_end);
          for (int j = 0; j < u.nboxes; ++j, ++gb) {
            const uint32_t slot = gb % kBSlots, ph = (gb / kBSlots) & 1;
            const int y0 = u.ra - 1 + j * kBBoxRows;
            mbar_wait(&bar_empty[slot], ph ^ 1);
            BT2(L, 4, int(gb - gb_l0));
            const bool mirror = (slot == 0) && (j > 0);
            mbar_expect_tx(&bar_full[slot], mirror ? 2 * kBSlotBytes : kBSlotBytes);
            tma_load_4d_hint(&maps.act[ly.in], &bar_full[slot], smem_u32(ring + slot * kBSlotBytes), 0, -1, y0, u.n, pol);
            if (mirror)
              tma_load_4d_hint(&maps.act[ly.in], &bar_full[slot], smem_u32(ring + kBSlots * kBSlotBytes), 0, -1, y0,
                               u.n, pol);
          }
          g += u.t1 - u.t0;
        }
      }
      __syncwarp();
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
    uint32_t gb_base = 0, tile_ctr = 0, se_seen = 0;
    long long m_acc = 0, m_full = 0, m_issue = 0, m_t = 0, m_fl = 0, m_pl = 0, m_ffull = 0;
    const long long m_start = BDBG ? clock64() : 0;
    for (int L = 0; L < p.n_layers; ++L) {
      // SE layers: the shuffles / loads of the SE chain crawl while MMAs saturate the shared-memory pipe
      // (40k cycles instead of a few k), so the issuers hold back until the vector is ready; TMA keeps
      // filling the ring meanwhile.
      if (FEN_SE_GATE && body_layer(p, L).epi == kBEpiSeResidual) {
        mbar_wait(&bar_se, se_seen & 1);
        ++se_seen;
      }
      const long long m_l0 = BDBG ? clock64() : 0;
      const long long m_full0 = m_full;
      const bool m_is_fused = body_layer(p, L).epi == kBEpiSeResidual;
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
          if (leader && mine) BT2(L, 0, i_layer);
          if (leader && !mine) BT2(L, 11, i_layer);
          if (mine) mbar_wait(&bar_acc_empty[acc], ((tile_ctr / kBodyAccBufs) & 1) ^ 1);
          if (BDBG) { const long long n_ = clock64(); m_acc += n_ - m_t; m_t = n_; }
          if (leader && mine) BT2(L, 6, i_layer);
          const int base = kTileM * t - kPitch * u.ra;
          const int need_last = min((base + kTileM + kMaxShift - 1) / kBBoxPx, u.nboxes - 1);
          while (mine && waited <= need_last) {
            const uint32_t gb = gb_base + waited;
            mbar_wait(&bar_full[gb % kBSlots], (gb / kBSlots) & 1);
            ++waited;
          }
          if (BDBG) { const long long n_ = clock64(); m_full += n_ - m_t; m_t = n_; }
          if (leader && i_layer == 0) BTRACE(L, 1);   // data for the first tile of the layer present
          if (leader && mine) BT2(L, 5, i_layer);
          if (leader && !mine) BT2(L, 12, i_layer);
          tc_fence_after();
          if (leader && !mine) BT2(L, 13, i_layer);
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
          if (leader && mine) BT2(L, 1, i_layer);
          if (leader && !mine) BT2(L, 14, i_layer);
          if (BDBG) { const long long n_ = clock64(); m_issue += n_ - m_t; m_t = n_; }
          const int next_first = (t + 1 < u.t1) ? (base + kTileM) / kBBoxPx : u.nboxes;
          while (released < next_first) {
            if (leader) umma_commit(&bar_empty[(gb_base + released) % kBSlots]);
            ++released;
          }
          if (leader) BT2(L, 9, i_layer);
          if (leader && mine) umma_commit(&bar_acc_full[acc]);
          if (leader) BT2(L, 10, i_layer);
          if (leader && i_layer >= n_tiles - 2) BTRACE(L, 2 + (i_layer == n_tiles - 1));   // last two tiles issued
          __syncwarp();
        }
        gb_base += u.nboxes;
        g += u.t1 - u.t0;
      }
      if (BDBG) { if (m_is_fused) { m_fl += clock64() - m_l0; m_ffull += m_full - m_full0; } else m_pl += clock64() - m_l0; }
    }
    if (BDBG && leader && warp == kBodyFirstMmaWarp) {
      long long* d = p.dbg + blockIdx.x * 16;
      d[15] = m_fl;
      d[6] = m_acc; d[7] = m_full; d[8] = m_issue; d[9] = clock64() - m_start;
    }
  } else {
    // ============================================================ epilogue (8 warps, 256 threads)
    constexpr int CW = 32;
    const int q = warp & 3;
    const int half = (warp - kBodyFirstEpiWarp) >> 2;
    const int col0 = half * CW;
    const int row_in_tile = q * 32 + lane;
    const int et = tid - 32 * kBodyFirstEpiWarp;         // 0..255
    const int ew = et >> 5;
    const bool flag_writer = (warp == kBodyFirstEpiWarp);
    uint32_t tile_ctr = 0;
    long long e_wait = 0, e_done = 0, e_t = 0, e_swait = 0, e_c2 = 0, e_c2w = 0, e_c1 = 0, e_c1w = 0, e_se = 0;
    const long long e_start = BDBG ? clock64() : 0;
    for (int L = 0; L < p.n_layers; ++L) {
      const BodyLayer ly = body_layer(p, L);
      bf16* outp = p.buf[ly.out];
      const bf16* resp = ly.res >= 0 ? p.buf[ly.res] : nullptr;
      // per-layer constants of this thread's 32 columns, in registers
      float bias[CW], slope[CW];
#pragma unroll
      for (int j = 0; j < CW / 4; ++j) {
        const float4 b4 = *reinterpret_cast<const float4*>(c_vec + ly.cv_bias + col0 + 4 * j);
        bias[4 * j] = b4.x; bias[4 * j + 1] = b4.y; bias[4 * j + 2] = b4.z; bias[4 * j + 3] = b4.w;
      }
      if (ly.epi == kBEpiPreluHsum) {
#pragma unroll
        for (int j = 0; j < CW / 4; ++j) {
          const float4 s4 = *reinterpret_cast<const float4*>(c_vec + ly.cv_slope + col0 + 4 * j);
          slope[4 * j] = s4.x; slope[4 * j + 1] = s4.y; slope[4 * j + 2] = s4.z; slope[4 * j + 3] = s4.w;
        }
      }
      if (ly.epi == kBEpiSeResidual) {
        // ---- SE vector of every image of this CTA, from the sums of h that the conv1 layer left.
        // There is no L1 left beside 226 KB of shared memory, so every global load is an L2 round trip
        // (~1k cycles): all weight loads are data-independent and are requested BEFORE waiting for the
        // peers; after the wait the chain is one round trip (border pixels of h) plus six short phases.
        if (BDBG) e_t = clock64();
        const bf16* w2 = reinterpret_cast<const bf16*>(p.packed + ly.w_off);        // [tap][c][ci] bf16
        const uint8_t* rec = p.packed + p.k_rcab0 + int64_t(ly.rcab) * p.k_rcab_stride;
        const float* fc0 = reinterpret_cast<const float*>(rec + p.k_rcab_fc0);
        const float* fc2 = reinterpret_cast<const float*>(rec + p.k_rcab_fc2);
        const int mc = et & 63, mq = et >> 6;              // mat-vec: output channel, quarter of the input channels
        const int f1j = et >> 4, f1p = et & 15;            // FC1: hidden unit, 4 of its 64 inputs
        const int f2c = et >> 2, f2p = et & 3;             // FC2: channel, quarter of the hidden units
        const bool fast_fc = (p.R == 16);
        float4 f1 = make_float4(0.f, 0.f, 0.f, 0.f), f2 = f1;
        if (fast_fc) {
          f1 = __ldg(reinterpret_cast<const float4*>(fc0 + f1j * kC + f1p * 4));
          f2 = __ldg(reinterpret_cast<const float4*>(fc2 + f2c * 16 + f2p * 4));
        }
        mbar_wait(&bar_flags, (L - 1) & 1);              // peers have finished the conv1 layer
        if (BDBG) { const long long n_ = clock64(); e_swait += n_ - e_t; e_t = n_; }
        if (et == 0) BTRACE(L, 8);
        for (int u0 = 0; img0 + u0 <= img1; u0 += 2) {
          const int nimg = min(2, img1 - (img0 + u0) + 1);
          // -- P1: border rows / columns of h (one warp per image and quantity), corners, totals.
          // Every lane issues ALL its loads before touching any result (one L2 round trip): divergent
          // branches with a load each would serialise into one round trip per branch.
          {
            const int li = ew >> 2, qn = ew & 3;
            if (li < nimg) {
              const bf16* hb = p.buf[kBufH] + size_t(img0 + u0 + li) * p.H * p.W * kC;
              const int chunk = lane & 7, pg = lane >> 3;
              const int npix = (qn < 2) ? p.W : p.H;
              // extra load of this lane: warps qn<2 -> corner pixels of their row (lanes 8..23); warps qn>=2 ->
              // totals from the conv1 epilogue (lanes 24..31); other lanes re-read pixel 0 (ignored)
              const bool is_corner = (qn < 2) && (lane >= 8) && (lane < 24);
              const bool is_total = (qn >= 2) && (lane >= 24);
              const int cx = (lane >= 16) ? p.W - 1 : 0;
              const uint4 xv = ld_cg_128(hb + (size_t((qn & 1) ? p.H - 1 : 0) * p.W + (is_corner ? cx : 0)) * kC + chunk * 8);
              const int c4 = ((qn & 1) * 8 + (lane & 7)) * 4;
              const float4 t4 = ld_cg_f32x4(p.hsum + (size_t(ly.rcab) * p.B + img0 + u0 + li) * kC + c4);
              float a8[8];
#pragma unroll
              for (int e = 0; e < 8; ++e) a8[e] = 0.f;
#pragma unroll 16
              for (int px = pg; px < npix; px += 4) {
                const size_t pix = (qn == 0) ? size_t(px) : (qn == 1) ? size_t(p.H - 1) * p.W + px
                                 : (qn == 2) ? size_t(px) * p.W : size_t(px) * p.W + p.W - 1;
                const uint4 v = ld_cg_128(hb + pix * kC + chunk * 8);
                a8[0] += bf16lo(v.x); a8[1] += bf16hi(v.x); a8[2] += bf16lo(v.y); a8[3] += bf16hi(v.y);
                a8[4] += bf16lo(v.z); a8[5] += bf16hi(v.z); a8[6] += bf16lo(v.w); a8[7] += bf16hi(v.w);
              }
#pragma unroll
              for (int e = 0; e < 8; ++e) {
                a8[e] += __shfl_xor_sync(0xffffffffu, a8[e], 8);
                a8[e] += __shfl_xor_sync(0xffffffffu, a8[e], 16);
              }
              if (lane < 8) {
                float4* d = reinterpret_cast<float4*>(&s_q[li][kHsRow0 + qn][chunk * 8]);
                d[0] = make_float4(a8[0], a8[1], a8[2], a8[3]);
                d[1] = make_float4(a8[4], a8[5], a8[6], a8[7]);
              } else if (is_corner) {                       // (row qn ? H-1 : 0, column cx)
                float4* d = reinterpret_cast<float4*>(&s_q[li][kHsC00 + 2 * qn + (lane >= 16 ? 1 : 0)][chunk * 8]);
                d[0] = make_float4(bf16lo(xv.x), bf16hi(xv.x), bf16lo(xv.y), bf16hi(xv.y));
                d[1] = make_float4(bf16lo(xv.z), bf16hi(xv.z), bf16lo(xv.w), bf16hi(xv.w));
              } else if (is_total) {
                *reinterpret_cast<float4*>(&s_q[li][kHsTotal][c4]) = t4;
              }
            }
          }
          // mat-vec weights (this thread's 16 input channels x 9 taps of output channel mc): requested now,
          // consumed in P3 - their L2 round trip hides behind P2 (not earlier: P1 needs the registers)
          uint4 wreg[18];
#pragma unroll
          for (int tap = 0; tap < 9; ++tap) {
            const uint4* wp = reinterpret_cast<const uint4*>(w2 + (size_t(tap) * kC + mc) * kC + mq * 16);
            wreg[2 * tap] = __ldg(wp);
            wreg[2 * tap + 1] = __ldg(wp + 1);
          }
          if (et == 0) BTRACE(L, 9);
          named_bar_sync(1, kEpiThreads);
          if (et == 0) BTRACE(L, 10);
          // -- P2: S_tap = total - excluded border row - excluded border column + corner
          for (int i = et; i < nimg * 9 * kC; i += kEpiThreads) {
            const int li = i / (9 * kC), r = i - li * 9 * kC;
            const int tap = r >> 6, ci = r & 63, dy = tap / 3 - 1, dx = tap % 3 - 1;
            float v = s_q[li][kHsTotal][ci];
            if (dy == 1) v -= s_q[li][kHsRow0][ci];
            if (dy == -1) v -= s_q[li][kHsRowL][ci];
            if (dx == 1) v -= s_q[li][kHsCol0][ci];
            if (dx == -1) v -= s_q[li][kHsColL][ci];
            if (dy == 1 && dx == 1) v += s_q[li][kHsC00][ci];
            if (dy == 1 && dx == -1) v += s_q[li][kHsC0L][ci];
            if (dy == -1 && dx == 1) v += s_q[li][kHsCL0][ci];
            if (dy == -1 && dx == -1) v += s_q[li][kHsCLL][ci];
            s_S[li][tap][ci] = v;
          }
          named_bar_sync(1, kEpiThreads);
          if (et == 0) BTRACE(L, 11);
          // -- P3: mat-vec, thread (c, quarter): 16 input channels of all 9 taps, both images
          {
            float a0 = 0.f, a1 = 0.f;
#pragma unroll
            for (int tap = 0; tap < 9; ++tap) {
#pragma unroll
              for (int j = 0; j < 2; ++j) {
                const uint4 w = wreg[2 * tap + j];
                const float4 p0 = *reinterpret_cast<const float4*>(&s_S[0][tap][mq * 16 + 8 * j]);
                const float4 p1 = *reinterpret_cast<const float4*>(&s_S[0][tap][mq * 16 + 8 * j + 4]);
                a0 = fmaf(bf16lo(w.x), p0.x, a0); a0 = fmaf(bf16hi(w.x), p0.y, a0);
                a0 = fmaf(bf16lo(w.y), p0.z, a0); a0 = fmaf(bf16hi(w.y), p0.w, a0);
                a0 = fmaf(bf16lo(w.z), p1.x, a0); a0 = fmaf(bf16hi(w.z), p1.y, a0);
                a0 = fmaf(bf16lo(w.w), p1.z, a0); a0 = fmaf(bf16hi(w.w), p1.w, a0);
                if (nimg > 1) {
                  const float4 r0 = *reinterpret_cast<const float4*>(&s_S[1][tap][mq * 16 + 8 * j]);
                  const float4 r1 = *reinterpret_cast<const float4*>(&s_S[1][tap][mq * 16 + 8 * j + 4]);
                  a1 = fmaf(bf16lo(w.x), r0.x, a1); a1 = fmaf(bf16hi(w.x), r0.y, a1);
                  a1 = fmaf(bf16lo(w.y), r0.z, a1); a1 = fmaf(bf16hi(w.y), r0.w, a1);
                  a1 = fmaf(bf16lo(w.z), r1.x, a1); a1 = fmaf(bf16hi(w.z), r1.y, a1);
                  a1 = fmaf(bf16lo(w.w), r1.z, a1); a1 = fmaf(bf16hi(w.w), r1.w, a1);
                }
              }
            }
            s_part[0][mq][mc] = a0;
            s_part[1][mq][mc] = a1;
          }
          named_bar_sync(1, kEpiThreads);
          if (et == 0) BTRACE(L, 12);
          // -- P4: mean of o = b2 + (W2 . S) / HW
          if (et < 128) {
            const int li = et >> 6, c = et & 63;
            s_mean[li][c] = c_vec[ly.cv_bias + c] +
                            (s_part[li][0][c] + s_part[li][1][c] + s_part[li][2][c] + s_part[li][3][c]) * p.inv_hw;
          }
          named_bar_sync(1, kEpiThreads);
          if (et == 0) BTRACE(L, 13);
          // -- P5: FC1 + ReLU
          if (fast_fc) {
#pragma unroll
            for (int li = 0; li < 2; ++li) {
              const float4 m4 = *reinterpret_cast<const float4*>(&s_mean[li][f1p * 4]);
              float a = f1.x * m4.x + f1.y * m4.y + f1.z * m4.z + f1.w * m4.w;
              a += __shfl_xor_sync(0xffffffffu, a, 1); a += __shfl_xor_sync(0xffffffffu, a, 2);
              a += __shfl_xor_sync(0xffffffffu, a, 4); a += __shfl_xor_sync(0xffffffffu, a, 8);
              if (f1p == 0) s_hid[li][f1j] = fmaxf(a, 0.f);
            }
          } else if (et < 2 * p.R) {
            const int li = et / p.R, j = et - li * p.R;
            float a = 0.f;
            for (int k = 0; k < kC; ++k) a = fmaf(__ldg(fc0 + j * kC + k), s_mean[li][k], a);
            s_hid[li][j] = fmaxf(a, 0.f);
          }
          named_bar_sync(1, kEpiThreads);
          if (et == 0) BTRACE(L, 14);
          // -- P6: FC2 + sigmoid
          {
            float a0 = 0.f, a1 = 0.f;
            if (fast_fc) {
              const float4 h0 = *reinterpret_cast<const float4*>(&s_hid[0][f2p * 4]);
              const float4 h1 = *reinterpret_cast<const float4*>(&s_hid[1][f2p * 4]);
              a0 = f2.x * h0.x + f2.y * h0.y + f2.z * h0.z + f2.w * h0.w;
              a1 = f2.x * h1.x + f2.y * h1.y + f2.z * h1.z + f2.w * h1.w;
            } else {
              const int rq = p.R >> 2;
              for (int j = f2p * rq; j < (f2p + 1) * rq; ++j) {
                const float w = __ldg(fc2 + f2c * p.R + j);
                a0 = fmaf(w, s_hid[0][j], a0);
                a1 = fmaf(w, s_hid[1][j], a1);
              }
            }
            a0 += __shfl_xor_sync(0xffffffffu, a0, 1); a0 += __shfl_xor_sync(0xffffffffu, a0, 2);
            a1 += __shfl_xor_sync(0xffffffffu, a1, 1); a1 += __shfl_xor_sync(0xffffffffu, a1, 2);
            if (f2p < nimg) {                               // lane `f2p` finishes image `f2p` of the pair
              const int n = img0 + u0 + f2p;
              const float sv = 1.f / (1.f + expf(-(f2p == 0 ? a0 : a1)));
              s_scale[u0 + f2p][f2c] = sv * p.res_scale;
              // the CTA owning tile 0 of the image publishes the attention vector
              if (p.se_out && (n * p.tiles_per_seg >= g_begin) && (n * p.tiles_per_seg < g_end))
                p.se_out[(size_t(n) * (p.G * p.Bk) + ly.rcab) * kC + f2c] = sv;
            }
          }
          named_bar_sync(1, kEpiThreads);
        }
        if (et == 0) mbar_arrive(&bar_se);                // release the MMA issuers (see there)
        if (BDBG) e_se += clock64() - e_t;
        if (et == 0) BTRACE(L, 4);   // SE vector ready
      }
      const long long e_l0 = BDBG ? clock64() : 0;
      const long long e_w0 = e_wait;
      for (int g = g_begin; g < g_end;) {
        const BUnit u = body_unit(p, g, g_end);
        float csum[CW];
#pragma unroll
        for (int c = 0; c < CW; ++c) csum[c] = 0.f;
        if (ly.epi == kBEpiSeResidual) {                  // `slope` doubles as the SE scale of this image
          const int us = u.n - img0;
#pragma unroll
          for (int j = 0; j < CW / 4; ++j) {
            const float4 s4 = *reinterpret_cast<const float4*>(&s_scale[us][col0 + 4 * j]);
            slope[4 * j] = s4.x; slope[4 * j + 1] = s4.y; slope[4 * j + 2] = s4.z; slope[4 * j + 3] = s4.w;
          }
        }
        for (int t = u.t0; t < u.t1; ++t, ++tile_ctr) {
          const uint32_t acc = tile_ctr & (kBodyAccBufs - 1);
          const int lin = kTileM * t + row_in_tile;
          const int y = lin / kPitch, x = lin - y * kPitch;
          const bool valid = (x < kStripW) && (y < p.H);
          const size_t opix = (size_t(u.n) * p.H + y) * p.W + x;
          // residual / skip values of this pixel: requested before waiting for the accumulator so the
          // L2 round trip overlaps the MMAs of the tile
          uint32_t rv[16];                          // 32 bf16, two 256-bit loads (one L1 line lookup each per lane)
#ifdef FEN_EXP_NORES
          if (false) {
#else
          if (valid && ly.epi != kBEpiPreluHsum) {
#endif
            const bf16* rsd = resp + opix * kC + col0;
            ld_cg_256_hint(rsd, kPolicyEvictFirst, *reinterpret_cast<uint32_t(*)[8]>(&rv[0]));
            ld_cg_256_hint(rsd + 16, kPolicyEvictFirst, *reinterpret_cast<uint32_t(*)[8]>(&rv[8]));
          } else {
#pragma unroll
            for (int j = 0; j < 16; ++j) rv[j] = 0u;
          }
          if (BDBG) e_t = clock64();
          mbar_wait(&bar_acc_full[acc], (tile_ctr / kBodyAccBufs) & 1);
          if (BDBG) e_wait += clock64() - e_t;
          if (et == 0 && tile_ctr % n_tiles == 0) BTRACE(L, 5);   // first accumulator of the layer ready
          if (et == 0) BT2(L, 2, int(tile_ctr % n_tiles));
          tc_fence_after();
          uint32_t v[CW];
          tmem_ld_32x32(tmem_base + acc * N + col0 + (uint32_t(q * 32) << 16), v);
          tmem_ld_wait();
          tc_fence_before();                       // accumulator read: hand it back to the MMA issuers
          __syncwarp();
          if (lane == 0) mbar_arrive(&bar_acc_empty[acc]);
          float f[CW];
#pragma unroll
          for (int c = 0; c < CW; ++c) f[c] = __uint_as_float(v[c]) + bias[c];
          if (ly.epi == kBEpiPreluHsum) {
#pragma unroll
            for (int c = 0; c < CW; ++c) f[c] = fmaxf(f[c], 0.f) + slope[c] * fminf(f[c], 0.f);
          } else {
            if (ly.epi == kBEpiSeResidual) {
#pragma unroll
              for (int c = 0; c < CW; ++c) f[c] *= slope[c];
            }
#pragma unroll
            for (int j = 0; j < 16; ++j) {          // + x (RCAB residual) or + skip (group / long skip)
              f[2 * j] += bf16lo(rv[j]);
              f[2 * j + 1] += bf16hi(rv[j]);
            }
          }
          if (valid) {
            uint32_t o[16];
#pragma unroll
            for (int e = 0; e < 16; ++e) o[e] = pack_bf16(f[2 * e], f[2 * e + 1]);
#ifndef FEN_EXP_NOSTORE
            st_global_256(outp + opix * kC + col0, *reinterpret_cast<uint32_t(*)[8]>(&o[0]));
            st_global_256(outp + opix * kC + col0 + 16, *reinterpret_cast<uint32_t(*)[8]>(&o[8]));
#else
            if (o[0] == 0x12345678u && o[9] == 0x9abcdef0u) st_global_256(outp + opix * kC + col0, *reinterpret_cast<uint32_t(*)[8]>(&o[0]));
#endif
            if (ly.epi == kBEpiPreluHsum) {          // sum of the bf16-ROUNDED h (what conv2 will read)
#pragma unroll
              for (int e = 0; e < 16; ++e) {
                csum[2 * e] += bf16lo(o[e]);
                csum[2 * e + 1] += bf16hi(o[e]);
              }
            }
          }
          if (et == 0) BT2(L, 3, int(tile_ctr % n_tiles));
        }
        if (ly.epi == kBEpiPreluHsum) {
          float* hs = p.hsum + (size_t(ly.rcab) * p.B + u.n) * kC;
          // total: reduce-scatter butterfly over the warp, lane l ends with channel col0 + l
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
          atomicAdd(hs + col0 + lane, csum[0]);
        }
        g += u.t1 - u.t0;
      }
      if (BDBG) {
        if (ly.epi == kBEpiSeResidual) { e_c2 += clock64() - e_l0; e_c2w += e_wait - e_w0; }
        else { e_c1 += clock64() - e_l0; e_c1w += e_wait - e_w0; }
      }
      if (et == 0) BTRACE(L, 6);     // last tile stored
      // ---- layer done for this warp: make its global writes visible, then publish the CTA's flag
      __threadfence();
      __syncwarp();
      if (lane == 0) mbar_arrive(&bar_done);
      if (flag_writer) {
        if (BDBG) e_t = clock64();
        mbar_wait(&bar_done, L & 1);
        if (BDBG) e_done += clock64() - e_t;
        if (lane == 0) {
          st_release_gpu(p.flags + blockIdx.x, L + 1);
          BTRACE(L, 7);                // flag published
        }
        __syncwarp();
      }
    }
    if (BDBG && flag_writer && lane == 0) {
      long long* d = p.dbg + blockIdx.x * 16;
      d[10] = e_wait; d[11] = e_done; d[12] = clock64() - e_start;
      d[0] = e_swait; d[2] = e_se; d[3] = e_c2; d[4] = e_c2w; d[13] = e_c1; d[14] = e_c1w;
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == kBodyFirstMmaWarp) tmem_dealloc(tmem_base, kTmemCols);
}

#endif  // FEN_DEV
}  // namespace fen
