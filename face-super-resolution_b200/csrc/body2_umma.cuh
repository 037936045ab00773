// Persistent body kernel, second generation: all 64->64 3x3 convolutions of the residual body in
// one cooperative launch (see body_umma.cuh for the layer chain and the references:
// src/models/custom.py:167-175, src/models/blocks.py:75-92,135-153,185-189).
//
// What changed against body_umma.cuh, and why (B200 timelines: profiles/r01_body2_trace.txt):
//  * Table-driven issue loops.  Every layer has the same geometry, so the per-tile bookkeeping (ring
//    position of the tile's view, boxes to wait for / to release, image of the tile) is computed ONCE
//    into shared-memory tables.  The old issuer warps spent ~2500 cycles of scalar work per tile, in
//    lock-step, against 1728 cycles of MMA time - the tensor pipe idled a third of every layer.
//    Each issuer now only visits its own tiles (even / odd).
//  * Two interleaved image sets.  The batch is split in two halves that are processed alternately
//    (layer L set 0, layer L set 1, layer L+1 set 0, ...).  A layer boundary costs a release/acquire
//    round trip between CTAs sharing an image (halo rows, SE pool); with two independent sets that
//    round trip (and the SE vector) of one set hides behind the other set's tiles.
//  * Squeeze-and-excitation on the tensor core.  mean_hw(conv2(h)) is linear in 9 border-corrected
//    channel sums of h (see body_umma.cuh).  The conv1 epilogue now accumulates all 9 sums itself
//    (no re-read of h), a dedicated SE warp turns them into a 8-row bf16 hi/lo-split operand in
//    shared memory, and ONE extra batch of 36 tcgen05.mma against the conv2 weights that are in shared
//    memory anyway produces the 64x576 mat-vec of up to 4 images.  The SE warp finishes with the two
//    tiny FC layers + sigmoid.  The old 20k-cycle CUDA-core chain on the epilogue warps is gone.
//  * One running activation ring over all passes, the first boxes of a layer requested before its weights;
//    tile runs of q or q + 1 tiles with the extra tile on opposite ends for the two sets (all 148 SMs busy).
//  * Rules every role follows (found with cuda-gdb): every mbarrier wait is executed by the
//    whole warp and followed by __syncwarp() before any tcgen05 / TMA issue; elected-lane blocks never
//    wait; a parity wait is only trusted by a waiter that has seen every earlier phase of the barrier or
//    has been told by a counter that its phase has begun (the ring: two barriers per slot + s_issued).
//  * The two issuers take turns, a whole tile each (FEN_B2_TURN): one feeds the pipe its 36 MMAs - a real
//    loop over the taps, which one thread issues at the pipe's rate - while the other does the waits of
//    its next tile.  The SE warp issues the MMAs of its mat-vec itself (FEN_B2_SE_SELF) and is the third
//    party of the weight hand-back.  tools/body2_protocol_sim.py is a discrete-event model of all of
//    this (tests/test_body2_protocol.py).
#pragma once
#include <type_traits>
#include "body_umma.cuh"

namespace fen {

#ifndef FEN_B2_TRACE
#define FEN_B2_TRACE 0   // 1: pass-level timeline of CTA 70 into Body2Params::dbg (developer builds)
#endif
#ifndef FEN_B2_TRACE_CTA
#define FEN_B2_TRACE_CTA 70
#endif
#ifndef FEN_B2_WATCH
#define FEN_B2_WATCH 0   // 1: every role logs (stage, L, s, i) into Body2Params::dbg (host-mapped memory) - hang post-mortems
#endif
#define B2W(role, stage, L, s, i) do { if (FEN_B2_WATCH && p.dbg && lane == 0) { *(volatile long long*)(p.dbg + blockIdx.x * 8 + (role)) = (long long)(stage) | ((long long)(L) << 8) | ((long long)(s) << 20) | ((long long)(i) << 24); } } while (0)
// per-tile trace (FEN_B2_TRACE=1): passes 80..83 of CTA 70 -> dbg[4096 + (P-80)*512 + e*32 + i]
#define B2T2(P, e, i) do { if (FEN_B2_TRACE && p.dbg && blockIdx.x == FEN_B2_TRACE_CTA && (P) >= 80 && (P) < 84 && (i) < 32) p.dbg[4096 + ((P) - 80) * 512 + (e) * 32 + (i)] = clock64(); } while (0)
#define B2TS(P, e) do { if (FEN_B2_TRACE && p.dbg && blockIdx.x == FEN_B2_TRACE_CTA && (P) < 512 && lane == 0) p.dbg[6144 + (P) * 8 + (e)] = clock64(); } while (0)
#define B2TRACE(P, e) do { if (FEN_B2_TRACE && p.dbg && blockIdx.x == FEN_B2_TRACE_CTA && (P) < 512) p.dbg[(P) * 8 + (e)] = clock64(); } while (0)

#ifndef FEN_B2_WFREE1
#define FEN_B2_WFREE1 1   // 1: the weights of a layer are handed back with ONE tcgen05.commit per issuer (after its last tile's
#endif                    //    last tap) instead of one per tap.  Same time (3.14 ms both), and the last tile of a layer no
                          //    longer carries 10 commits: every tile tools/soak2.py ever caught wrong was such a tile

#ifndef FEN_B2_ISSUE
#define FEN_B2_ISSUE 2    // shape of the 36-MMA issue block of a tile: 0 fully unrolled, 1 a loop over the 3 tap rows (12 MMAs
#endif                    //   per iteration), 2 a loop over the 9 taps (4 MMAs per iteration)
#ifndef FEN_B2_ROTATE
#define FEN_B2_ROTATE 0   // 1: the issuers' tile shares rotate per pass (balances tiles + SE batches between the two issuers; measured
#endif                    //    SLOWER: 3.17 - 3.20 ms against 3.09 at batch 64)
#ifndef FEN_B2_TURN
#define FEN_B2_TURN 1      // 1: the issuers take turns, one whole tile each (see the tile loop)
#endif
#ifndef FEN_B2_SE_SELF
#define FEN_B2_SE_SELF 1   // 1: the SE warp issues the 36 MMAs of its mat-vec itself (no hand-over to issuer A)
#endif
#ifndef FEN_B2_NEAR_FLAGS
#define FEN_B2_NEAR_FLAGS 0   // 1: the TMA warp only waits for the CTAs whose rows its boxes read (c - 1, c, c + 1) instead of every CTA of the image
#endif
#ifndef FEN_B2_NI
#define FEN_B2_NI 2   // MMA issuer warps.  One thread sustains ~81 cycles per tcgen05.mma (tools/umma_probe3.cu), the pipe
#endif                //   takes one N = 64 MMA every ~48: at least two issuers must be inside their 36-MMA loops at any time
constexpr int kB2Issuers = FEN_B2_NI;
constexpr int kB2SeWarp = 0;                       // warp % 4 == 0: may read TMEM lanes 0..31 (the SE result rows)
constexpr int kB2TmaWarp = 1;
constexpr int kB2FirstMmaWarp = 2;                 // warps 2 .. 2 + kB2Issuers - 1
constexpr int kB2FirstEpiWarp = kB2FirstMmaWarp + kB2Issuers;   // 8 epilogue warps: lane quarter = warp & 3, column half = (warp - first) >> 2
constexpr int kB2EpiWarps = 8;
constexpr int kB2Threads = 32 * (kB2FirstEpiWarp + kB2EpiWarps);
#ifndef FEN_B2_SMEM_CONST
#define FEN_B2_SMEM_CONST (FEN_B2_NI > 2)   // 1: the epilogue reads bias / slope / SE scale from shared memory per tile instead of
#endif                                      //    caching 64 values in registers (more warps per CTA = fewer registers per thread)
#ifndef FEN_B2_ACCBUFS
#define FEN_B2_ACCBUFS 7
#endif
constexpr int kB2AccBufs = FEN_B2_ACCBUFS;         // 7 x 64 TMEM columns for conv tiles ...
constexpr uint32_t kB2SeCol = kB2AccBufs * kC;     // ... + 64 columns for the SE mat-vec
constexpr int kB2SBytes = 9 * 1024;                // SE operand: one 8-row SWIZZLE_128B atom per tap
#ifndef FEN_B2_STAGED_STORE
#define FEN_B2_STAGED_STORE 0   // 1: outputs go through a shared-memory transpose to coalesced stores (measured slower)
#endif
constexpr int kB2Slots = FEN_B2_STAGED_STORE ? 6 : 7;   // activation ring: two-row boxes, + 1 mirror slot (6: same time, 5: + 2 %)
constexpr int kB2RingPx = kB2Slots * kBBoxPx;      // 792 pixels, + the 132-pixel mirror slot
constexpr int kB2RingBytes = (kB2Slots + 1) * kBSlotBytes;
constexpr int kB2StageBytes = 2048;                // per epilogue warp: 32 pixels x 32 channels bf16, SWIZZLE_64B
constexpr int kB2DynBytes = kBodyWBytes + kB2SBytes + kB2RingBytes + (FEN_B2_STAGED_STORE ? kB2EpiWarps * kB2StageBytes : 0) + 1024;
constexpr int kB2MaxTiles = 64;
constexpr int kB2MaxBoxes = 96;

struct B2Tile { uint16_t m; uint8_t wait_upto, first_box, unit, t; uint16_t pad; };   // views of the tile lie in boxes [first_box, wait_upto)
struct B2Box { int16_t img; int8_t y0; uint8_t mirror; uint8_t last_tile; uint8_t pad[3]; };   // last_tile: last tile of the pass reading the box
struct B2Unit { int img, t0, t1, pad; };

// BodyParams::B is the whole batch; total_tiles / tiles_per_cta count the tiles of ONE set.
// Tensor maps (kernel parameters stay below 4 KB): per activation buffer one load map (box 64 ch x 66 px x
// 2 rows, SWIZZLE_128B) and one store map (box 32 ch x 32 px, SWIZZLE_64B); the packed weights.
// Tensor maps: ONE 5-D load map over all activation buffers of the workspace ([buffer][B][H][W][64] bf16, the buffers
// lie at a constant stride; box 64 ch x 66 px x 2 rows, SWIZZLE_128B) - a layer's input is a coordinate, not a map, so
// the training variant can address its ~190 saved tensors - and the packed weights.
struct Body2Maps {
  CUtensorMap act;
  CUtensorMap w;
};

struct Body2Params : BodyParams {
  int nset;      // 1 or 2 interleaved image sets
  int set_B;     // images per set (B = nset * set_B)
  long long* hsum64;   // [n_rcab][B][9][64] fixed-point (2^-24) channel sums of the bf16-rounded h: integer atomics make the
                       // SE pool - and with it the whole forward - independent of the order in which warps and CTAs arrive
  bf16* act_base;      // buffer i = act_base + i * act_elems
  int64_t act_elems;
  // training variant (body2_umma_kernel<true>): what fen_backward needs besides the activations themselves
  uint32_t* mask0;     // per RCAB [B][H][W][2] words: sign bits of conv1's pre-activation (ConvParams::mask_out layout)
  int64_t mask_stride; //   words between RCABs
  long long* pool_sums;  // [n_rcab][B][64] fixed point: sum over pixels of o = conv2(h) + b2 (the SE pool, as kEpiSum leaves it)
  const float* cvec;   // bias / slope table in the packed blob (BodyLayer::cv_bias / cv_slope index it), 512 B readable past any cv_bias
};

__device__ __forceinline__ float2 ld_cg_hs_x2(const long long* p) {     // two adjacent fixed-point sums -> float
  long long a, b;
  asm volatile("ld.global.cg.v2.s64 {%0, %1}, [%2];" : "=l"(a), "=l"(b) : "l"(p));
  return make_float2(__ll2float_rn(a) * (1.f / kHsScale), __ll2float_rn(b) * (1.f / kHsScale));
}
// Buffer indices.  Inference: the BodyBuf enum (F0, X0, X1, H, O, G0 ...: ping-pong buffers).  Training: every
// tensor the backward needs is kept - [f0][body][x' of RCAB 0..n-1][h ...][o ...][group outputs ...] (StepWs order).
struct B2Layer : BodyLayer { int out2; };       // out2: where conv2 also stores o (training), else -1
template <bool kTrain>
__device__ __forceinline__ B2Layer body2_layer(const Body2Params& p, int L) {
  B2Layer l;
  static_cast<BodyLayer&>(l) = body_layer(p, L);
  l.out2 = -1;
  if (!kTrain) return l;
  const int n = p.G * p.Bk;
  const int per_group = 2 * p.Bk + 1;
  const int g = L / per_group, r = L - g * per_group;
  auto xs = [&](int rc) { return 2 + rc; };
  auto gout = [&](int gg) { return 2 + 3 * n + gg; };
  l.last_use = 0;                                  // everything is read again by the backward
  if (g == p.G) { l.in = gout(p.G - 1); l.res = 0; l.out = 1; return l; }
  const int gin = (g == 0) ? 0 : gout(g - 1);
  if (r == 2 * p.Bk) { l.in = xs(g * p.Bk + p.Bk - 1); l.res = gin; l.out = gout(g); return l; }
  const int b = r >> 1, rc = g * p.Bk + b;
  const int xb = (b == 0) ? gin : xs(rc - 1);
  if ((r & 1) == 0) { l.in = xb; l.out = 2 + n + rc; }
  else { l.in = 2 + n + rc; l.res = xb; l.out = xs(rc); l.out2 = 2 + 2 * n + rc; }
  return l;
}

__device__ __forceinline__ void tma_load_5d_hint(const CUtensorMap* m, uint64_t* bar, uint32_t dst_smem, int c0, int c1,
                                                 int c2, int c3, int c4, uint64_t policy) {
  asm volatile(
      "cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint"
      " [%0], [%1, {%3, %4, %5, %6, %7}], [%2], %8;"
      :
      : "r"(dst_smem), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3),
        "r"(c4), "l"(policy)
      : "memory");
}

__device__ __forceinline__ float2 ld_cg_f32x2(const float* p) {
  float2 v;
  asm volatile("ld.global.cg.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "l"(p));
  return v;
}
__device__ __forceinline__ void st_shared_u128(uint32_t saddr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(saddr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ uint4 ld_shared_u128(uint32_t saddr) {
  uint4 v;
  asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(saddr));
  return v;
}
__device__ __forceinline__ void st_global_128(void* ptr, const uint4& v) {
  asm volatile("st.global.v4.b32 [%0], {%1, %2, %3, %4};" ::"l"(ptr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ void tma_store_4d(const CUtensorMap* m, uint32_t src_smem, int c0, int c1, int c2, int c3) {
  asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];"
               :
               : "l"(reinterpret_cast<uint64_t>(m)), "r"(src_smem), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
               : "memory");
}
// plain (non-tensor) bulk copy global -> shared, completion on an mbarrier; 16-byte aligned, size a multiple of 16
__device__ __forceinline__ void bulk_load_1d(uint32_t dst_smem, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               :
               : "r"(dst_smem), "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void st_release_cta_shared(uint32_t* p, uint32_t v) {
  asm volatile("st.release.cta.shared::cta.u32 [%0], %1;" ::"r"(smem_u32(p)), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_cta_shared(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.cta.shared::cta.u32 %0, [%1];" : "=r"(v) : "r"(smem_u32(p)) : "memory");
  return v;
}
__device__ __forceinline__ void st_shared_u32(uint32_t saddr, uint32_t v) {
  asm volatile("st.shared.u32 [%0], %1;" ::"r"(saddr), "r"(v) : "memory");
}

// (registers are allocated per group of 4 warps: 12 warps get 168 registers per thread, 13 .. 16 warps 128)
template <bool kTrain>
__global__ void __launch_bounds__(kB2Threads, 1)
body2_umma_kernel(const __grid_constant__ Body2Maps maps, const Body2Params p) {
  constexpr int N = kC;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* w_smem = smem;                              // [9][64][64] bf16, SWIZZLE_128B
  uint8_t* s_buf = smem + kBodyWBytes;                 // [9] atoms of 8 rows x 128 B (rows 2u, 2u+1: hi / lo of unit u)
  uint8_t* ring = s_buf + kB2SBytes;                   // 6 slots + mirror
  uint8_t* stage = ring + kB2RingBytes;                // 8 x 2 KB output staging (one per epilogue warp)
  // TWO mbarriers per ring slot, used alternately (box g -> slot g % kB2Slots, barrier g % (2 kB2Slots), parity of
  // g / (2 kB2Slots)).  An issuer only waits for the boxes its own tiles read (issuer B never sees the first box of a
  // pass arrive).  With one barrier per slot, B asking for box g while box g - 7 - same slot, a box only A's tile reads -
  // was still in flight found the barrier one phase behind its own count, and a parity wait cannot tell "phase n not
  // complete" from "phase n - 1 not complete": it passed, and B's tile read a slot whose box had not been requested
  // yet.  That was the rare single wrong tile of tools/soak2.py (always B's last tile of a pass, from its last lane
  // quarter on; tools/body2_protocol_sim.py reproduces it from the protocol alone).  With two barriers the previous
  // phase of box g's barrier belongs to box g - 14, which has landed before box g - 7 could be REQUESTED, and box g - 7
  // has been requested before any box B's previous tile read (boxes are requested in order) as long as two consecutive
  // tiles of an issuer lie less than 7 boxes apart.  s_issued makes it unconditional: the TMA warp publishes the running
  // count of boxes it has requested, and an issuer only trusts the parity wait for box g once s_issued > g - box g is
  // only requested after every reader of box g - 7 has completed, so all earlier phases of its barrier are complete.
  __shared__ uint64_t bar_w[9], bar_wfree[9], bar_full[2 * kB2Slots];
  __shared__ uint32_t s_issued;
  __shared__ int s_hist[kB2Slots];                     // TMA warp: running index of the last tile that reads the box now in each slot
  __shared__ uint64_t bar_acc_full[kB2AccBufs], bar_acc_empty[kB2AccBufs];
  __shared__ uint64_t bar_done, bar_s_ready, bar_s_free, bar_se_full, bar_se_empty, bar_scale[2];
  __shared__ uint64_t bar_turn[4];                     // FEN_B2_TURN: tile G's issuer arrives on [G % 4] once its MMAs are issued
  __shared__ uint64_t bar_cv[2];                       // per-layer bias / slope vectors staged by the TMA warp (slot L & 1)
  __shared__ __align__(16) float s_cv[2][128];
  __shared__ uint32_t tmem_slot;
  // per-pass tables, one set of them per image set (the two sets give a CTA runs of different length)
  __shared__ B2Tile tile_tab2[2][kB2MaxTiles];
  __shared__ B2Box box_tab2[2][kB2MaxBoxes];
  __shared__ B2Unit unit_tab2[2][kBodyMaxUnits];
  __shared__ int s_meta[2][4];                         // per set: tiles, boxes, units, -
  __shared__ int s_peer[2][2];                         // per set: first / last CTA sharing an image with this one
  __shared__ __align__(16) float s_scale[2][kBodyMaxUnits][kC];   // res_scale * s, per set and unit
  __shared__ __align__(16) float s_mean[kBodyMaxUnits][kC], s_hid[kBodyMaxUnits][kC];
  __shared__ __align__(16) float s_raw[2 * kBodyMaxUnits][kC];     // raw SE accumulator rows (hi / lo per unit)
  __shared__ __align__(16) uint32_t s_colx[kB2EpiWarps][2][16];   // per epilogue warp: one staged pixel (32 bf16) per border column

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  constexpr uint32_t kTmemCols = 512;

  // Tile runs.  The T tiles of a set are dealt to the C CTAs as evenly as possible (q or q + 1 tiles); set 0
  // gives the extra tile to the FIRST T % C CTAs, set 1 to the LAST ones, so that over a layer a CTA handles at
  // most ceil(2T / C) tiles instead of 2 ceil(T / C) (batch 64: 15 instead of 16, on 148 CTAs instead of 132).
  const int T = p.total_tiles, Cn = int(gridDim.x);
  const int rq = T / Cn, rr = T % Cn;
  auto run_begin = [&](int c, int s) { return (s == 0) ? c * rq + min(c, rr) : c * rq + max(0, c - (Cn - rr)); };
  auto cta_of = [&](int g, int s) {
    if (s == 0) return (g < rr * (rq + 1)) ? g / (rq + 1) : rr + (g - rr * (rq + 1)) / rq;
    const int h0 = Cn - rr;
    return (g < h0 * rq) ? g / rq : h0 + (g - h0 * rq) / (rq + 1);
  };

  if (warp == kB2FirstMmaWarp) tmem_alloc(&tmem_slot, kTmemCols);
  if (tid == 0) {
    int max_tiles = 0;
    for (int s = 0; s < p.nset; ++s) {
      const int g_begin = run_begin(blockIdx.x, s), g_end = run_begin(blockIdx.x + 1, s);
      B2Tile* tile_tab = tile_tab2[s]; B2Box* box_tab = box_tab2[s]; B2Unit* unit_tab = unit_tab2[s];
      // ---- per-pass tables (identical for every layer)
      int b_cum = 0, i = 0, u = 0;
      for (int g = g_begin; g < g_end; ++u) {
        const BUnit bu = body_unit(p, g, g_end);
        unit_tab[u].img = bu.n; unit_tab[u].t0 = bu.t0; unit_tab[u].t1 = bu.t1; unit_tab[u].pad = 0;
        for (int j = 0; j < bu.nboxes; ++j) {
          B2Box e;
          e.img = int16_t(bu.n); e.y0 = int8_t(bu.ra - 1 + j * kBBoxRows);
          e.mirror = uint8_t(j > 0);   // continues the previous box of its unit: mirrored when it lands in slot 0
          e.last_tile = uint8_t(i); e.pad[0] = e.pad[1] = e.pad[2] = 0;
          box_tab[b_cum + j] = e;
        }
        for (int t = bu.t0; t < bu.t1; ++t, ++i) {
          const int base = kTileM * t - kPitch * bu.ra;
          B2Tile e;
          e.pad = 0;
          e.first_box = uint8_t(b_cum + base / kBBoxPx);
          e.m = uint16_t(b_cum * kBBoxPx + base);   // pixel position of the tile's view, relative to the pass's first box
          e.wait_upto = uint8_t(b_cum + min((base + kTileM + kMaxShift - 1) / kBBoxPx, bu.nboxes - 1) + 1);
          e.unit = uint8_t(u); e.t = uint8_t(t);
          tile_tab[i] = e;
          for (int b = e.first_box; b < e.wait_upto; ++b) box_tab[b].last_tile = uint8_t(i);
        }
        b_cum += bu.nboxes;
        g += bu.t1 - bu.t0;
      }
      s_meta[s][0] = i; s_meta[s][1] = b_cum; s_meta[s][2] = u; s_meta[s][3] = 0;
      max_tiles = max(max_tiles, i);
      // peers: CTAs owning tiles of the images this CTA touches in this set (including itself)
      const int img0 = g_begin / p.tiles_per_seg, img1 = (g_end - 1) / p.tiles_per_seg;
      s_peer[s][0] = cta_of(img0 * p.tiles_per_seg, s);
      s_peer[s][1] = cta_of(min(T, (img1 + 1) * p.tiles_per_seg) - 1, s);
    }
    // ---- barriers.  Issuer w takes the tiles w, w + n, w + 2n ... of every pass; an issuer that never gets a tile
    // (fewer tiles per pass than issuer warps, in every set) stays out of the weight release protocol entirely.
    // Ring slots need no release barriers: a slot may be refilled once every tile that reads its box has COMPLETED,
    // and tile completion is what bar_acc_full already signals (one tcgen05.commit per tile) - the TMA warp follows
    // those barriers (s_hist, `frontier`).
    const uint32_t n_issuers = uint32_t(min(kB2Issuers, max(max_tiles, 1)));
    s_meta[0][3] = int(n_issuers);
    for (int k = 0; k < kB2Slots; ++k) s_hist[k] = -1;
    s_issued = 0;
    for (int k = 0; k < 9; ++k) { mbar_init(&bar_w[k], 1); mbar_init(&bar_wfree[k], n_issuers + (FEN_B2_SE_SELF ? 1u : 0u)); }
    for (int k = 0; k < 2 * kB2Slots; ++k) mbar_init(&bar_full[k], 1);
    for (int k = 0; k < kB2AccBufs; ++k) { mbar_init(&bar_acc_full[k], 1); mbar_init(&bar_acc_empty[k], kB2EpiWarps); }
    mbar_init(&bar_done, kB2EpiWarps);
    mbar_init(&bar_s_ready, 1); mbar_init(&bar_s_free, 1); mbar_init(&bar_se_full, 1); mbar_init(&bar_se_empty, 1);
    mbar_init(&bar_scale[0], 1); mbar_init(&bar_scale[1], 1);
    mbar_init(&bar_cv[0], 1); mbar_init(&bar_cv[1], 1);
    for (int k = 0; k < 4; ++k) mbar_init(&bar_turn[k], 1);
    fence_mbar_init();
    tma_prefetch_desc(&maps.w);
  }
  for (int k = tid; k < kB2SBytes / 16; k += kB2Threads) reinterpret_cast<uint4*>(s_buf)[k] = make_uint4(0, 0, 0, 0);
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_slot;
  const int n_issuers = s_meta[0][3];

  if (warp == kB2TmaWarp) {
    // ============================================================ TMA issuer + peer-flag poller
    // The ring is one running sequence of boxes over all passes (box g -> slot g % kB2Slots, use g / kB2Slots),
    // so the first boxes of a pass are prefetched while the previous pass still owns the other slots.
    uint32_t P = 0, slot = 0, fb = 0, n_req = 0, gbase = 0, frontier = 0;   // fb: barrier of the next box (running box index n_req mod 2 kB2Slots); frontier: every tile below it (running index) is complete
    constexpr int kPre = 4;                       // boxes of a new layer requested BEFORE its weights
    for (int L = 0; L < p.n_layers; ++L) {
      const B2Layer ly = body2_layer<kTrain>(p, L);
      const uint64_t pol = ly.last_use ? kPolicyEvictFirst : 0x1000000000000000ull;
      auto wait_flags = [&](int s) {
        // every peer must have finished layer L-1 of this set: their outputs are my inputs / halos
        B2W(1, 2, L, s, 0);
        if (L > 0) {
          const int* fl = p.flags + s * int(gridDim.x);
#if FEN_B2_NEAR_FLAGS
          // halo rows come from the neighbouring runs only (a run is at least one tile = 128 pixels, a halo row 66); the
          // buffers this CTA overwrites in layer L were read by the same neighbours in layer L - 1
          const int k0 = max(s_peer[s][0], int(blockIdx.x) - 1), k1 = min(s_peer[s][1], int(blockIdx.x) + 1);
#else
          const int k0 = s_peer[s][0], k1 = s_peer[s][1];
#endif
          for (int k = k0 + lane; k <= k1; k += 32)
            while (ld_acquire_gpu(fl + k) < L) { __nanosleep(20); }
          __syncwarp();
          fence_proxy_async_all();
        }
      };
      // Every wait below is executed by the WHOLE (converged) warp; lane 0 only issues, in straight-line blocks (rule 1
      // of DESIGN.md 4.2 - TMA issue is a uniform-datapath instruction like tcgen05: it must not sit in an elected-lane
      // block that also spins).
      auto issue_boxes = [&](int s, int b0, int b1, uint32_t gbase_pass) {
        const int img_base = s * p.set_B;
        for (int b = b0; b < b1; ++b) {
          const B2Box e = box_tab2[s][b];
          if (lane == 0) B2T2(P, 9, b);
          B2W(1, 3, L, s, b);
          // the box this one replaces must have been read by every tile that uses it: wait for those tiles to complete
          const int need = s_hist[slot];
          while (int(frontier) <= need) {
            mbar_wait(&bar_acc_full[frontier % kB2AccBufs], (frontier / kB2AccBufs) & 1u);
            ++frontier;
          }
          __syncwarp();                              // (every lane has read s_hist[slot] before lane 0 replaces it)
          const bool mirror = e.mirror && slot == 0;
          if (lane == 0) {
            s_hist[slot] = int(gbase_pass) + int(e.last_tile);
            mbar_expect_tx(&bar_full[fb], mirror ? 2 * kBSlotBytes : kBSlotBytes);
            tma_load_5d_hint(&maps.act, &bar_full[fb], smem_u32(ring + slot * kBSlotBytes), 0, -1, e.y0,
                             img_base + e.img, ly.in, pol);
            if (mirror)
              tma_load_5d_hint(&maps.act, &bar_full[fb], smem_u32(ring + kB2Slots * kBSlotBytes), 0, -1, e.y0,
                               img_base + e.img, ly.in, pol);
            st_release_cta_shared(&s_issued, n_req + 1);
            B2T2(P, 10, b);
          }
          __syncwarp();
          if (++slot == kB2Slots) slot = 0;
          if (++fb == 2 * kB2Slots) fb = 0;
          ++n_req;
        }
      };
      const int pre = min(kPre, s_meta[0][1]);
      wait_flags(0);
      if (lane == 0) B2TRACE(P, 0);
      issue_boxes(0, 0, pre, gbase);
      // bias (+ slope) of this layer -> s_cv[L & 1].  The slot was last read by the epilogue of layer L - 2, which
      // is complete: wait_flags saw this CTA's own flag reach L, i.e. its epilogue has finished layer L - 1.
      if (lane == 0) {
        mbar_expect_tx(&bar_cv[L & 1], 512);
        bulk_load_1d(smem_u32(&s_cv[L & 1][0]), p.cvec + ly.cv_bias, 512, &bar_cv[L & 1]);
      }
      __syncwarp();
      for (int tap = 0; tap < 9; ++tap) {     // weights, tap by tap, as soon as the previous layer released the tap
        B2W(1, 1, L, 0, tap);
        if (L > 0 && (!FEN_B2_WFREE1 || tap == 0)) mbar_wait(&bar_wfree[tap], (L - 1) & 1);
        __syncwarp();
        if (lane == 0) {
          mbar_expect_tx(&bar_w[tap], N * kC * 2);
          tma_load_2d(&maps.w, &bar_w[tap], w_smem + tap * N * 128, 0, ly.w_row + tap * N);
        }
        __syncwarp();
      }
      for (int s = 0; s < p.nset; ++s, ++P) {
        if (s > 0) {
          wait_flags(s);
          if (lane == 0) B2TRACE(P, 0);
        }
        issue_boxes(s, s == 0 ? pre : 0, s_meta[s][1], gbase);
        gbase += uint32_t(s_meta[s][0]);
        __syncwarp();
      }
    }
  } else if (warp >= kB2FirstMmaWarp && warp < kB2FirstMmaWarp + kB2Issuers) {
    // ============================================================ MMA issuers: warp 2 + w takes tiles w, w + n, ... (FEN_B2_SE_SELF = 0: warp 2 also the SE batches)
    constexpr uint32_t idesc = umma_idesc_bf16(kTileM, N);
    constexpr uint32_t kDescHi = (1024u >> 4) | (1u << 14) | (2u << 29);
    constexpr uint32_t kLbo = 1u << 16;
    const uint32_t ring_lo = (smem_u32(ring) >> 4) | kLbo;
    const uint32_t w_lo = (smem_u32(w_smem) >> 4) | kLbo;
    const uint32_t s_lo = (smem_u32(s_buf) >> 4) | kLbo;
    const bool leader = elect_one();
    const int wi = warp - kB2FirstMmaWarp;
    uint32_t P = 0, gbase = 0, gbox = 0, se_n = 0;
    for (int L = 0; L < ((wi < n_issuers) ? p.n_layers : 0); ++L) {
      const bool conv2 = body_layer(p, L).epi == kBEpiSeResidual;   // (the layer kinds do not depend on the variant)
      bool w_seen = false;
      int se_done = 0;
      for (int s = 0; s < p.nset; ++s, ++P) {
        const B2Tile* tile_tab = tile_tab2[s];
        const int n_tiles = s_meta[s][0], n_boxes = s_meta[s][1];
        // Which tiles of the pass this issuer takes: w0, w0 + n, ...  With a fixed w0 = wi issuer A gets 4 of 7 tiles in
        // every pass AND the SE batches (5 blocks of 36 MMAs against B's 3 in a conv2 pass), and whenever one issuer idles
        // the other feeds the pipe at the single-thread rate.  The start index therefore rotates: in conv2 passes the SE
        // issuer (A) takes the odd tiles - the smaller share -, in the other layers the extra tile alternates with the set.
#if FEN_B2_ROTATE
        const int rot = (n_issuers == 2 && n_tiles >= 2) ? (conv2 ? 1 : (s & 1)) : 0;
#else
        const int rot = 0;
#endif
        const int w0 = (wi + rot) % n_issuers;
        const int last_own = ((n_tiles - 1 - w0) >= 0) ? w0 + n_issuers * ((n_tiles - 1 - w0) / n_issuers) : -1;
        const uint32_t gb0 = gbox;                                  // running index of the pass's first box
        const uint32_t gbase_pass = gbase;
        gbox += uint32_t(n_boxes); gbase += uint32_t(n_tiles);
        const uint32_t start_px = (gb0 % kB2Slots) * kBBoxPx;
        const bool last_pass = (s == p.nset - 1);
        // SE batches of this layer still owed: the one of set s must run before this pass's epilogue can start,
        // those of later sets may run as soon as their operand is ready (W2 is in shared memory all layer long)
#if defined(FEN_B2_X2) || FEN_B2_SE_SELF
        const bool se_layer = false;
#else
        const bool se_layer = conv2 && (wi == 0);
#endif
        if (s == 0) se_done = 0;
        uint32_t waited = 0;
        // one SE batch: D[row, c] = sum_tap S_tap[row, :] . W2_tap[c, :] into the SE accumulator
        auto issue_se = [&]() {
          B2W(2 + wi, 5, L, s, se_n);
          if (se_n >= 1) mbar_wait(&bar_se_empty, (se_n - 1) & 1);
          tc_fence_after();
          // The batch only starts once ALL taps of this layer's weights have landed, i.e. once no MMA of the
          // previous layer is in flight any more.  Issued tap by tap behind the weight loads, interleaved
          // with the other issuer's last tile of the previous layer, it raised "warp out-of-range address"
          // on B200 for even tile counts (the only case where this warp enters a layer first).
          if (!w_seen) {
            for (int tap = 0; tap < 9; ++tap) mbar_wait(&bar_w[tap], L & 1);
          }
          __syncwarp();                            // converge after the spin-waits before any tcgen05 issue
          if (leader) {
#if FEN_B2_ISSUE == 0
#pragma unroll
            for (int tap = 0; tap < 9; ++tap) {
              const uint32_t a_lo = s_lo + tap * (1024 >> 4);
              const uint32_t b_lo = w_lo + tap * (N * 128 >> 4);
#pragma unroll
              for (int k = 0; k < 4; ++k) {
                umma_bf16_ss_lohi(tmem_base + kB2SeCol, a_lo + 2 * k, b_lo + 2 * k, kDescHi, idesc, (tap | k) != 0);
              }
            }
#else
            uint32_t a_lo = s_lo, b_lo = w_lo;      // (a real loop: see the tile loop below)
#pragma unroll 1
            for (int tap = 0; tap < 9; ++tap) {
#pragma unroll
              for (int k = 0; k < 4; ++k)
                umma_bf16_ss_lohi_p(tmem_base + kB2SeCol, a_lo + 2 * k, b_lo + 2 * k, kDescHi, idesc, k ? 1u : uint32_t(tap));
              a_lo += 1024 >> 4;
              b_lo += N * 128 >> 4;
            }
#endif
            umma_commit(&bar_se_full);
            umma_commit(&bar_s_free);
          }
          w_seen = true;
          ++se_n;
          ++se_done;
          __syncwarp();
        };
        if (last_own < 0) {
          // no tile for this issuer in this pass (fewer tiles than issuers): only the weights need its hand-back, if
          // this is the layer's last pass (the commit covers this warp's MMAs of the other set)
          __syncwarp();
          if (leader && last_pass)
            for (int tap = 0; tap < (FEN_B2_WFREE1 ? 1 : 9); ++tap) umma_commit(&bar_wfree[tap]);
          __syncwarp();
          continue;
        }
        for (int i = w0; i < n_tiles; i += n_issuers) {
          const B2Tile e = tile_tab[i];
          const uint32_t G = gbase_pass + i, acc = G % kB2AccBufs, aph = (G / kB2AccBufs) & 1;
          B2W(2 + wi, 1, L, s, i);
          if (se_layer && se_done <= s) {
            // The epilogue of this pass cannot free accumulators before the SE batch of this set has run: never
            // block on an accumulator while that batch is still owed.
            if (i == last_own) {
              mbar_wait(&bar_s_ready, se_n & 1);
              issue_se();
            } else {
              for (;;) {
                // (never as the first batch of a layer: see issue_se)
                if (w_seen && __any_sync(0xffffffffu, mbar_test_wait(&bar_s_ready, se_n & 1))) { issue_se(); break; }
                if (__any_sync(0xffffffffu, mbar_test_wait(&bar_acc_empty[acc], aph ^ 1))) break;
              }
            }
          }
#ifndef FEN_B2_NOEARLY
          if (se_layer && se_done > s && se_done < p.nset && w_seen &&
              __any_sync(0xffffffffu, mbar_test_wait(&bar_s_ready, se_n & 1)))
            issue_se();                          // a later set's batch, ahead of time
#endif
          B2W(2 + wi, 2, L, s, i);
          if (leader) B2T2(P, 4, i);
          mbar_wait(&bar_acc_empty[acc], aph ^ 1);
          if (leader) B2T2(P, 5, i);
          B2W(2 + wi, 3, L, s, i);
          // only the boxes this tile reads (earlier ones may already have been replaced: their readers are complete)
          if (waited < e.first_box) waited = e.first_box;
          while (waited < e.wait_upto) {
            const uint32_t g = gb0 + waited;
            uint64_t* bf = &bar_full[g % (2 * kB2Slots)];
            const uint32_t bph = (g / (2 * kB2Slots)) & 1u;
            while (ld_acquire_cta_shared(&s_issued) <= g) mbar_try_wait(bf, bph);   // (not requested yet: try_wait only as a pause)
            mbar_wait(bf, bph);
            ++waited;
          }
          if (leader && i == w0 && wi < 2) B2TRACE(P, 1 + wi);
          B2W(2 + wi, 4, L, s, i);
          if (leader) B2T2(P, 6, i);
          __syncwarp();                            // converge after the spin-waits (see the commits below)
          tc_fence_after();
          const uint32_t d_tmem = tmem_base + acc * N;
          const uint32_t m = (uint32_t(e.m) + start_px) % kB2RingPx;
          const bool w_rel = last_pass && (i == last_own);
          // Elected-lane blocks hold straight-line tcgen05 code only: every wait is executed by the whole
          // (converged) warp.  The first tile of a layer follows the weight loads tap by tap.
#define FEN_B2_ISSUE_TAP(tap)                                                                              \
  {                                                                                                        \
    const uint32_t off = ((tap) / 3) * kPitch + ((tap) % 3);                                               \
    uint32_t pos = m + off;                                                                                \
    if (pos >= uint32_t(kB2RingPx)) pos -= kB2RingPx;                                                      \
    const uint32_t a_lo = ring_lo + pos * 8;                                                               \
    const uint32_t b_lo = w_lo + (tap) * (N * 128 >> 4);                                                   \
    _Pragma("unroll") for (int k = 0; k < 4; ++k)                                                          \
        umma_bf16_ss_lohi(d_tmem, a_lo + 2 * k, b_lo + 2 * k, kDescHi, idesc, ((tap) | k) != 0);           \
    if (w_rel && (!FEN_B2_WFREE1 || (tap) == 8)) umma_commit(&bar_wfree[FEN_B2_WFREE1 ? 0 : (tap)]); /* the next layer's tap may overwrite once these MMAs finish */ \
  }
#if FEN_B2_TURN
          // The issuers take turns, a whole tile each, in tile order: while one feeds the pipe its 36 MMAs the other does
          // the bookkeeping of its next tile (the waits above, ~ 1 300 cycles).  Issuing concurrently they fall into
          // lock-step - both blocked by the same full pipe, both finishing together - and their gaps coincide.
          // Tile G - 1's issuer arrives on bar_turn[(G - 1) % 4]; tiles are issued strictly in order, so exactly
          // (G - 1) / 4 phases of that barrier are complete before the arrival: the parity wait is unambiguous.
          if (G > 0) mbar_wait(&bar_turn[(G - 1) & 3u], ((G - 1) >> 2) & 1u);
          __syncwarp();
#endif
          if (!w_seen) {
#pragma unroll
            for (int tap = 0; tap < 9; ++tap) {
              mbar_wait(&bar_w[tap], L & 1);
              __syncwarp();
              if (leader) FEN_B2_ISSUE_TAP(tap)
              __syncwarp();
            }
          } else if (leader) {
#if FEN_B2_ISSUE == 0 || !FEN_B2_WFREE1
#pragma unroll
            for (int tap = 0; tap < 9; ++tap) FEN_B2_ISSUE_TAP(tap)
#else
            // The fully unrolled form above lets ptxas hoist all 72 descriptor moves (R2UR) ahead of the first MMA; they do
            // not fit the uniform register file, and the block becomes a chain of uniform-register spills and fills: ~ 9
            // instructions and ~ 81 cycles per MMA from one thread, although one thread can issue an N = 64 MMA every
            // ~ 52 cycles (tools/umma_probe.cu T5).  A real loop bounds what can be hoisted.
            uint32_t b_lo = w_lo, row = m;
#if FEN_B2_ISSUE == 1
#pragma unroll 1
            for (int dy = 0; dy < 3; ++dy) {
#pragma unroll
              for (int dx = 0; dx < 3; ++dx) {
                uint32_t pos = row + dx;
                if (pos >= uint32_t(kB2RingPx)) pos -= kB2RingPx;
                const uint32_t a_lo = ring_lo + pos * 8;
#pragma unroll
                for (int k = 0; k < 4; ++k)
                  umma_bf16_ss_lohi_p(d_tmem, a_lo + 2 * k, b_lo + 2 * k, kDescHi, idesc, (dx | k) ? 1u : uint32_t(dy));
                b_lo += N * 128 >> 4;
              }
              row += kPitch;
              if (row >= uint32_t(kB2RingPx)) row -= kB2RingPx;
            }
#else
            uint32_t dx = 0;
#pragma unroll 1
            for (int tap = 0; tap < 9; ++tap) {
              uint32_t pos = row + dx;
              if (pos >= uint32_t(kB2RingPx)) pos -= kB2RingPx;
              const uint32_t a_lo = ring_lo + pos * 8;
#pragma unroll
              for (int k = 0; k < 4; ++k)
                umma_bf16_ss_lohi_p(d_tmem, a_lo + 2 * k, b_lo + 2 * k, kDescHi, idesc, k ? 1u : uint32_t(tap));
              b_lo += N * 128 >> 4;
              if (++dx == 3) {
                dx = 0;
                row += kPitch;
                if (row >= uint32_t(kB2RingPx)) row -= kB2RingPx;
              }
            }
#endif
            if (w_rel) umma_commit(&bar_wfree[0]);   // the next layer's weights may overwrite once these MMAs finish
#endif
          }
          w_seen = true;
          __syncwarp();
#if FEN_B2_TURN
          if (leader) mbar_arrive(&bar_turn[G & 3u]);
#endif
          if (leader) B2T2(P, 7, i);
          // ONE commit per tile: the accumulator is complete.  The epilogue reads it; the TMA warp learns from the same
          // barrier that the ring boxes this tile read may be replaced.
          if (leader) umma_commit(&bar_acc_full[acc]);
          if (leader) B2T2(P, 8, i);
          if (leader && i == last_own && wi < 2) B2TRACE(P, 3 + wi);
          __syncwarp();
        }
      }
    }
  } else if (warp == kB2SeWarp) {
    // ============================================================ SE warp
    uint32_t se_n = 0, m_cnt = 0;
    const uint32_t s_base = smem_u32(s_buf);
#if FEN_B2_SE_SELF
    constexpr uint32_t se_idesc = umma_idesc_bf16(kTileM, kC);
    constexpr uint32_t kSeDescHi = (1024u >> 4) | (1u << 14) | (2u << 29);
    const uint32_t se_s_lo = (smem_u32(s_buf) >> 4) | (1u << 16), se_w_lo = (smem_u32(w_smem) >> 4) | (1u << 16);
    const bool se_leader = elect_one();
#endif
#ifdef FEN_B2_X2
    for (int L = 0; L < 0; ++L) {
#else
    for (int L = 0; L < p.n_layers; ++L) {
#endif
      const B2Layer ly = body2_layer<kTrain>(p, L);
#if FEN_B2_SE_SELF
      // This warp reads the layer's weights too (its MMAs run against W2 in shared memory): it is a party of the weight
      // hand-back.  In a layer without SE it arrives at once - but never a phase ahead of the issuers.
      if (ly.epi != kBEpiSeResidual) {
        if (L > 0) mbar_wait(&bar_wfree[0], (L - 1) & 1);
        __syncwarp();
        if (lane == 0) mbar_arrive(&bar_wfree[0]);
        __syncwarp();
        continue;
      }
#else
      if (ly.epi != kBEpiSeResidual) continue;
#endif
      const uint8_t* rec = p.packed + p.k_rcab0 + int64_t(ly.rcab) * p.k_rcab_stride;
      const float* fc0 = reinterpret_cast<const float*>(rec + p.k_rcab_fc0);
      const float* fc2 = reinterpret_cast<const float*>(rec + p.k_rcab_fc2);
      // the two FC matrices (contiguous, 2 * R * 64 floats) into L1 long before they are needed
      for (int ofs = lane * 32; ofs < 2 * p.R * kC; ofs += 32 * 32)
        asm volatile("prefetch.global.L1 [%0];" ::"l"(fc0 + ofs));
      for (int s = 0; s < p.nset; ++s, ++se_n) {
        // the conv1 layer (L - 1) of this set must be complete on every peer: its epilogues own the sums
        B2W(0, 1, L, s, se_n);
        B2TS(L * p.nset + s, 0);
        {
          const int* fl = p.flags + s * int(gridDim.x);
          for (int k = s_peer[s][0] + lane; k <= s_peer[s][1]; k += 32)
            while (ld_acquire_gpu(fl + k) < L) { __nanosleep(20); }
          __syncwarp();
        }
        const B2Unit* unit_tab = unit_tab2[s];
        const int n_units = s_meta[s][2];
        const int img_base = s * p.set_B;
        // 9 sums x 2 channels per lane and unit, all requested before the first use (one L2 round trip)
        float2 qv[kBodyMaxUnits][kHsCount];
#pragma unroll
        for (int u = 0; u < kBodyMaxUnits; ++u) {
          if (u < n_units) {
            const long long* hs = p.hsum64 + (size_t(ly.rcab) * p.B + img_base + unit_tab[u].img) * (kHsCount * kC) + 2 * lane;
#pragma unroll
            for (int k = 0; k < kHsCount; ++k) qv[u][k] = ld_cg_hs_x2(hs + k * kC);
          }
        }
        B2W(0, 2, L, s, se_n);
        B2TS(L * p.nset + s, 1);
#if !FEN_B2_SE_SELF
        if (se_n >= 1) mbar_wait(&bar_s_free, (se_n - 1) & 1);   // the previous batch has consumed the operand
#endif                                                           // (SE_SELF: this warp has seen the previous batch complete)
#pragma unroll
        for (int u = 0; u < kBodyMaxUnits; ++u) {
          if (u < n_units) {
#pragma unroll
            for (int tap = 0; tap < 9; ++tap) {
              const int dy = tap / 3 - 1, dx = tap % 3 - 1;
              float2 v = qv[u][kHsTotal];
              if (dy == 1) { v.x -= qv[u][kHsRow0].x; v.y -= qv[u][kHsRow0].y; }
              if (dy == -1) { v.x -= qv[u][kHsRowL].x; v.y -= qv[u][kHsRowL].y; }
              if (dx == 1) { v.x -= qv[u][kHsCol0].x; v.y -= qv[u][kHsCol0].y; }
              if (dx == -1) { v.x -= qv[u][kHsColL].x; v.y -= qv[u][kHsColL].y; }
              if (dy == 1 && dx == 1) { v.x += qv[u][kHsC00].x; v.y += qv[u][kHsC00].y; }
              if (dy == 1 && dx == -1) { v.x += qv[u][kHsC0L].x; v.y += qv[u][kHsC0L].y; }
              if (dy == -1 && dx == 1) { v.x += qv[u][kHsCL0].x; v.y += qv[u][kHsCL0].y; }
              if (dy == -1 && dx == -1) { v.x += qv[u][kHsCLL].x; v.y += qv[u][kHsCLL].y; }
              // hi / lo bf16 split: hi + lo carries 16 mantissa bits of the fp32 sum
              const uint32_t hi = pack_bf16(v.x, v.y);
              const uint32_t lo = pack_bf16(v.x - bf16lo(hi), v.y - bf16hi(hi));
              const uint32_t chunk = uint32_t(lane >> 2), inner = uint32_t(lane & 3) * 4;   // channels 2*lane, 2*lane+1
              const uint32_t r0 = 2 * u, r1 = 2 * u + 1;
              st_shared_u32(s_base + tap * 1024 + r0 * 128 + ((chunk ^ r0) << 4) + inner, hi);
              st_shared_u32(s_base + tap * 1024 + r1 * 128 + ((chunk ^ r1) << 4) + inner, lo);
            }
          }
        }
        fence_proxy_async_smem();
        __syncwarp();
#if FEN_B2_SE_SELF
        {
          // D[row, c] = sum_tap S_tap[row, :] . W2_tap[c, :] into the SE accumulator, once all nine taps of THIS layer's
          // weights have landed (no MMA of the previous layer is in flight any more then)
          for (int tap = 0; tap < 9; ++tap) mbar_wait(&bar_w[tap], L & 1);
          __syncwarp();
          tc_fence_after();
          if (se_leader) {
            uint32_t a_lo = se_s_lo, b_lo = se_w_lo;
#pragma unroll 1
            for (int tap = 0; tap < 9; ++tap) {
#pragma unroll
              for (int k = 0; k < 4; ++k)
                umma_bf16_ss_lohi_p(tmem_base + kB2SeCol, a_lo + 2 * k, b_lo + 2 * k, kSeDescHi, se_idesc, k ? 1u : uint32_t(tap));
              a_lo += 1024 >> 4;
              b_lo += kC * 128 >> 4;
            }
            umma_commit(&bar_se_full);
            if (s == p.nset - 1) umma_commit(&bar_wfree[0]);   // the layer's last batch: the weights may go once it has run
          }
          __syncwarp();
        }
#else
        if (lane == 0) mbar_arrive(&bar_s_ready);
#endif
        B2TS(L * p.nset + s, 2);
        const bool fast_fc = (p.R == 16);   // lane l: hidden unit l & 15 of units (l >> 4), (l >> 4) + 2; channels l, l + 32
        // ---- result rows 0..7 of the SE accumulator (lane r = row r = hi / lo part of unit r / 2) -> smem
        B2W(0, 3, L, s, se_n);
        mbar_wait(&bar_se_full, se_n & 1);
        B2W(0, 4, L, s, se_n);
        B2TS(L * p.nset + s, 3);
        tc_fence_after();
#pragma unroll
        for (int hf = 0; hf < 2; ++hf) {
          uint32_t v[32];
          tmem_ld_32x32(tmem_base + kB2SeCol + 32 * hf, v);
          tmem_ld_wait();
          if (lane < 2 * kBodyMaxUnits) {
#pragma unroll
            for (int k = 0; k < 8; ++k)
              *reinterpret_cast<uint4*>(&s_raw[lane][32 * hf + 4 * k]) = make_uint4(v[4 * k], v[4 * k + 1], v[4 * k + 2], v[4 * k + 3]);
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&bar_se_empty);
        if (fast_fc) {
          // FC1 + ReLU: mean[c] = b2[c] + (hi + lo)[c] / HW, folded into the dot product
#pragma unroll
          for (int rep = 0; rep < 2; ++rep) {
            const int u = (lane >> 4) + 2 * rep;
            if (u < n_units) {
              float acc = 0.f;
#pragma unroll
              for (int k = 0; k < 16; ++k) {
                const float4 hi = *reinterpret_cast<const float4*>(&s_raw[2 * u][4 * k]);
                const float4 lo = *reinterpret_cast<const float4*>(&s_raw[2 * u + 1][4 * k]);
                const float4 b4 = __ldg(reinterpret_cast<const float4*>(p.cvec + ly.cv_bias) + k);
                const float4 w1 = __ldg(reinterpret_cast<const float4*>(fc0 + (lane & 15) * kC) + k);   // L1 hit (prefetched)
                acc = fmaf(w1.x, fmaf(hi.x + lo.x, p.inv_hw, b4.x), acc);
                acc = fmaf(w1.y, fmaf(hi.y + lo.y, p.inv_hw, b4.y), acc);
                acc = fmaf(w1.z, fmaf(hi.z + lo.z, p.inv_hw, b4.z), acc);
                acc = fmaf(w1.w, fmaf(hi.w + lo.w, p.inv_hw, b4.w), acc);
              }
              s_hid[u][lane & 15] = fmaxf(acc, 0.f);
            }
          }
          __syncwarp();
          // FC2 + sigmoid
          for (int u = 0; u < n_units; ++u) {
            float a0 = 0.f, a1 = 0.f;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              const float4 h4 = *reinterpret_cast<const float4*>(&s_hid[u][4 * k]);
              const float4 wa = __ldg(reinterpret_cast<const float4*>(fc2 + lane * 16) + k);
              const float4 wb = __ldg(reinterpret_cast<const float4*>(fc2 + (lane + 32) * 16) + k);
              a0 = fmaf(wa.x, h4.x, a0); a0 = fmaf(wa.y, h4.y, a0); a0 = fmaf(wa.z, h4.z, a0); a0 = fmaf(wa.w, h4.w, a0);
              a1 = fmaf(wb.x, h4.x, a1); a1 = fmaf(wb.y, h4.y, a1); a1 = fmaf(wb.z, h4.z, a1); a1 = fmaf(wb.w, h4.w, a1);
            }
            const float sv0 = 1.f / (1.f + expf(-a0)), sv1 = 1.f / (1.f + expf(-a1));
            s_scale[s][u][lane] = sv0 * p.res_scale;
            s_scale[s][u][lane + 32] = sv1 * p.res_scale;
            if (kTrain && unit_tab[u].t0 == 0) {     // sum_px o = HW b2 + (hi + lo): what se_bwd_apply_kernel reads back
              long long* ps = p.pool_sums + (size_t(ly.rcab) * p.B + img_base + unit_tab[u].img) * kC;
              const float hw = 1.f / p.inv_hw;
              ps[lane] = __float2ll_rn(kHsScale * fmaf(__ldg(p.cvec + ly.cv_bias + lane), hw, s_raw[2 * u][lane] + s_raw[2 * u + 1][lane]));
              ps[lane + 32] = __float2ll_rn(kHsScale * fmaf(__ldg(p.cvec + ly.cv_bias + lane + 32), hw, s_raw[2 * u][lane + 32] + s_raw[2 * u + 1][lane + 32]));
            }
#ifdef FEN_SE_DUMP3   /* developer: EVERY CTA sharing the image records the channel-0 pool total it read, slot = CTA % 64 */
            if (p.se_out && lane == 0)
              p.se_out[(size_t(img_base + unit_tab[u].img) * (p.G * p.Bk) + ly.rcab) * kC + (blockIdx.x & 63)] = qv[u][kHsTotal].x;
            if (false) {
#else
            if (p.se_out && unit_tab[u].t0 == 0) {   // the CTA owning tile 0 of the image publishes the attention vector
#endif
              float* so = p.se_out + (size_t(img_base + unit_tab[u].img) * (p.G * p.Bk) + ly.rcab) * kC;
              so[lane] = sv0;
              so[lane + 32] = sv1;
            }
          }
        } else {
          for (int idx = lane; idx < n_units * kC; idx += 32) {
            const int u = idx >> 6, c = idx & 63;
            s_mean[u][c] = __ldg(p.cvec + ly.cv_bias + c) + (s_raw[2 * u][c] + s_raw[2 * u + 1][c]) * p.inv_hw;
          }
          __syncwarp();
          for (int idx = lane; idx < n_units * p.R; idx += 32) {        // FC1 + ReLU
            const int u = idx / p.R, j = idx - u * p.R;
            float a = 0.f;
            for (int c = 0; c < kC; ++c) a = fmaf(__ldg(fc0 + j * kC + c), s_mean[u][c], a);
            s_hid[u][j] = fmaxf(a, 0.f);
          }
          __syncwarp();
          for (int idx = lane; idx < n_units * kC; idx += 32) {         // FC2 + sigmoid
            const int u = idx >> 6, c = idx & 63;
            float a = 0.f;
            for (int j = 0; j < p.R; ++j) a = fmaf(__ldg(fc2 + c * p.R + j), s_hid[u][j], a);
            const float sv = 1.f / (1.f + expf(-a));
            s_scale[s][u][c] = sv * p.res_scale;
            if (kTrain && unit_tab[u].t0 == 0)
              p.pool_sums[(size_t(ly.rcab) * p.B + img_base + unit_tab[u].img) * kC + c] = __float2ll_rn(kHsScale * (s_mean[u][c] / p.inv_hw));
            if (p.se_out && unit_tab[u].t0 == 0)
              p.se_out[(size_t(img_base + unit_tab[u].img) * (p.G * p.Bk) + ly.rcab) * kC + c] = sv;
          }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&bar_scale[s]);
        B2TS(L * p.nset + s, 4);
      }
      ++m_cnt;
    }
    (void)m_cnt;
  } else {
    // ============================================================ epilogue (8 warps)
    constexpr int CW = 32;
    const int q = warp & 3;
    const int ew = warp - kB2FirstEpiWarp;
    const int half = ew >> 2;
    const int col0 = half * CW;
    const int row_in_tile = q * 32 + lane;
    const bool flag_writer = (ew == 0);
    const uint32_t pair_u32 = smem_u32(stage + q * 2 * kB2StageBytes);   // 4 KB per lane quarter
    uint32_t G = 0, P = 0, m_cnt = 0;
    for (int L = 0; L < p.n_layers; ++L) {
      const B2Layer ly = body2_layer<kTrain>(p, L);
      bf16* outp = p.act_base + size_t(ly.out) * p.act_elems;
      const bf16* resp = ly.res >= 0 ? p.act_base + size_t(ly.res) * p.act_elems : nullptr;
      bf16* out2p = (kTrain && ly.out2 >= 0) ? p.act_base + size_t(ly.out2) * p.act_elems : nullptr;
      uint32_t* maskp = (kTrain && ly.epi == kBEpiPreluHsum) ? p.mask0 + size_t(ly.rcab) * p.mask_stride : nullptr;
      mbar_wait(&bar_cv[L & 1], (L >> 1) & 1);
      const float* s_bias_l = &s_cv[L & 1][col0];          // this layer's bias / PReLU slope (cv_slope = cv_bias + 64)
      const float* s_slope_l = &s_cv[L & 1][64 + col0];
#if !FEN_B2_SMEM_CONST
      float bias[CW], slope[CW];
#pragma unroll
      for (int j = 0; j < CW / 4; ++j) {
        const float4 b4 = *reinterpret_cast<const float4*>(s_bias_l + 4 * j);
        bias[4 * j] = b4.x; bias[4 * j + 1] = b4.y; bias[4 * j + 2] = b4.z; bias[4 * j + 3] = b4.w;
      }
      if (ly.epi == kBEpiPreluHsum) {
#pragma unroll
        for (int j = 0; j < CW / 4; ++j) {
          const float4 s4 = *reinterpret_cast<const float4*>(s_slope_l + 4 * j);
          slope[4 * j] = s4.x; slope[4 * j + 1] = s4.y; slope[4 * j + 2] = s4.z; slope[4 * j + 3] = s4.w;
        }
      }
#endif
      for (int s = 0; s < p.nset; ++s, ++P) {
        const B2Tile* tile_tab = tile_tab2[s];
        const B2Unit* unit_tab = unit_tab2[s];
        const int n_tiles = s_meta[s][0];
        const int img_base = s * p.set_B;
        if (ew == 0) B2W(4, 1, L, s, 0);
#ifndef FEN_B2_X2
        if (ly.epi == kBEpiSeResidual) mbar_wait(&bar_scale[s], m_cnt & 1);
#endif
        // The tile loop is instantiated once per epilogue kind: in one loop with run-time branches every kind's
        // loop-carried state (the 32 channel sums of conv1, the residual words of conv2 ...) stays live in all of
        // them, and the epilogue is register bound (168).
        auto run_tiles = [&](auto epi_tag) {
        constexpr int EPI = decltype(epi_tag)::value;
        int cur_unit = -1, img = 0;
        float csum[CW], col0sum = 0.f, colLsum = 0.f;
        long long* hs = nullptr;
        auto flush_unit = [&]() {
          if (EPI != kBEpiPreluHsum || cur_unit < 0) return;
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
          hs_add(hs + kHsTotal * kC + col0 + lane, csum[0]);
          hs_add(hs + kHsCol0 * kC + col0 + lane, col0sum);
          hs_add(hs + kHsColL * kC + col0 + lane, colLsum);
        };
        for (int i = 0; i < n_tiles; ++i, ++G) {
          const B2Tile e = tile_tab[i];
          if (int(e.unit) != cur_unit) {
            flush_unit();
            cur_unit = e.unit;
            img = img_base + unit_tab[cur_unit].img;
#pragma unroll
            for (int c = 0; c < CW; ++c) csum[c] = 0.f;
            col0sum = 0.f; colLsum = 0.f;
            if (EPI == kBEpiPreluHsum) hs = p.hsum64 + (size_t(ly.rcab) * p.B + img) * (kHsCount * kC);
#if !FEN_B2_SMEM_CONST
            if (EPI == kBEpiSeResidual) {             // `slope` doubles as the SE scale of this image
#pragma unroll
              for (int j = 0; j < CW / 4; ++j) {
                const float4 s4 = *reinterpret_cast<const float4*>(&s_scale[s][cur_unit][col0 + 4 * j]);
                slope[4 * j] = s4.x; slope[4 * j + 1] = s4.y; slope[4 * j + 2] = s4.z; slope[4 * j + 3] = s4.w;
              }
            }
#endif
          }
          const uint32_t acc = G % kB2AccBufs, aph = (G / kB2AccBufs) & 1;
          const int lin = kTileM * int(e.t) + row_in_tile;
          const int y = lin / kPitch, x = lin - y * kPitch;
          const bool valid = (x < kStripW) && (y < p.H);
          const size_t opix = (size_t(img) * p.H + y) * p.W + x;
          // residual / skip values of this pixel: requested before waiting for the accumulator
          uint32_t rv[16];
#ifdef FEN_EXP_NORES
          if (false) {
#else
          if (valid && EPI != kBEpiPreluHsum) {
#endif
            const bf16* rsd = resp + opix * kC + col0;
            ld_cg_256_hint(rsd, kPolicyEvictFirst, *reinterpret_cast<uint32_t(*)[8]>(&rv[0]));
            ld_cg_256_hint(rsd + 16, kPolicyEvictFirst, *reinterpret_cast<uint32_t(*)[8]>(&rv[8]));
          } else {
#pragma unroll
            for (int j = 0; j < 16; ++j) rv[j] = 0u;
          }
          if (ew == 0) B2W(4, 2, L, s, i);
          if (ew == 0 && lane == 0) B2T2(P, 0, i);
          mbar_wait(&bar_acc_full[acc], aph);
          if (ew == 0 && lane == 0 && i == 0) B2TRACE(P, 5);
          if (ew == 0 && lane == 0) B2T2(P, 1, i);
          tc_fence_after();
          uint32_t v[CW];
#ifdef FEN_EXP_NOEPI      // developer experiment: the accumulator is handed back unread, the tile's epilogue is skipped
#pragma unroll
          for (int c = 0; c < CW; ++c) v[c] = 0u;
          __syncwarp();
          if (lane == 0) mbar_arrive(&bar_acc_empty[acc]);
          if (v[0] == 0u) continue;
#else
          tmem_ld_32x32(tmem_base + acc * N + col0 + (uint32_t(q * 32) << 16), v);
          tmem_ld_wait();
          tc_fence_before();                       // accumulator read: hand it back to the MMA issuers
          __syncwarp();
          if (lane == 0) mbar_arrive(&bar_acc_empty[acc]);
#endif
          if (ew == 0 && lane == 0) B2T2(P, 2, i);
          float f[CW];
#if FEN_B2_SMEM_CONST
          // second operand of the epilogue (PReLU slope of the layer / SE scale of this image), broadcast reads
          const float* s_mul = (EPI == kBEpiSeResidual) ? &s_scale[s][cur_unit][col0] : s_slope_l;
#pragma unroll
          for (int j = 0; j < CW / 4; ++j) {
            const float4 b4 = *reinterpret_cast<const float4*>(s_bias_l + 4 * j);
            f[4 * j] = __uint_as_float(v[4 * j]) + b4.x; f[4 * j + 1] = __uint_as_float(v[4 * j + 1]) + b4.y;
            f[4 * j + 2] = __uint_as_float(v[4 * j + 2]) + b4.z; f[4 * j + 3] = __uint_as_float(v[4 * j + 3]) + b4.w;
          }
#define B2_FOR_MUL(OP)                                                                              \
  _Pragma("unroll") for (int j = 0; j < CW / 4; ++j) {                                               \
    const float4 m4 = *reinterpret_cast<const float4*>(s_mul + 4 * j);                               \
    OP(f[4 * j], m4.x) OP(f[4 * j + 1], m4.y) OP(f[4 * j + 2], m4.z) OP(f[4 * j + 3], m4.w)          \
  }
#else
#pragma unroll
          for (int c = 0; c < CW; ++c) f[c] = __uint_as_float(v[c]) + bias[c];
#define B2_FOR_MUL(OP) _Pragma("unroll") for (int c = 0; c < CW; ++c) { OP(f[c], slope[c]) }
#endif
#define B2_OP_PRELU(x, m) x = fmaxf(x, 0.f) + (m) * fminf(x, 0.f);
#define B2_OP_SCALE(x, m) x *= (m);
          if (EPI == kBEpiPreluHsum) {
            if (kTrain) {                           // sign bits of the pre-activation, for the PReLU backward
              uint32_t mbits = 0;
#pragma unroll
              for (int c = 0; c < CW; ++c) mbits |= (f[c] > 0.f ? 1u : 0u) << c;
              if (valid) maskp[opix * 2 + half] = mbits;
            }
            B2_FOR_MUL(B2_OP_PRELU)
          } else {
            if (EPI == kBEpiSeResidual) {
              if (kTrain) {                         // o = conv2(h) + b2 is kept: the SE backward needs sum dx' * o
                uint32_t o2[16];
#pragma unroll
                for (int k = 0; k < 16; ++k) o2[k] = pack_bf16(f[2 * k], f[2 * k + 1]);
                if (valid) {
                  st_global_256(out2p + opix * kC + col0, *reinterpret_cast<uint32_t(*)[8]>(&o2[0]));
                  st_global_256(out2p + opix * kC + col0 + 16, *reinterpret_cast<uint32_t(*)[8]>(&o2[8]));
                }
              }
              B2_FOR_MUL(B2_OP_SCALE)
            }
#pragma unroll
            for (int j = 0; j < 16; ++j) {          // + x (RCAB residual) or + skip (group / long skip)
              f[2 * j] += bf16lo(rv[j]);
              f[2 * j + 1] += bf16hi(rv[j]);
            }
          }
          uint32_t o[16];
#pragma unroll
          for (int k = 0; k < 16; ++k) o[k] = pack_bf16(f[2 * k], f[2 * k + 1]);
#if FEN_B2_STAGED_STORE
          {
            // ---- output: registers -> staging of this lane quarter (32 px x 128 B, shared by the two warps
            // that hold the two channel halves; 16 B chunks XOR-swizzled by the pixel index) -> coalesced
            // 128-bit global stores, 4 full 128 B lines per instruction.  (Writing straight from the TMEM
            // layout - one pixel per lane, 128 B apart - costs 16x the LSU wavefronts; TMA tensor stores
            // cannot be used: a row-straddling warp needs a negative start coordinate, which faults.)
            const uint32_t sw = uint32_t(lane) & 7u;
            const uint32_t prow = pair_u32 + lane * 128;
            st_shared_u128(prow + (((4u * half + 0u) ^ sw) << 4), o[0], o[1], o[2], o[3]);
            st_shared_u128(prow + (((4u * half + 1u) ^ sw) << 4), o[4], o[5], o[6], o[7]);
            st_shared_u128(prow + (((4u * half + 2u) ^ sw) << 4), o[8], o[9], o[10], o[11]);
            st_shared_u128(prow + (((4u * half + 3u) ^ sw) << 4), o[12], o[13], o[14], o[15]);
            named_bar_sync(1 + q, 64);
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              const int px = 16 * half + 4 * k + (lane >> 3);
              const uint32_t ch = uint32_t(lane) & 7u;
              const uint4 v4 = ld_shared_u128(pair_u32 + px * 128 + ((ch ^ (uint32_t(px) & 7u)) << 4));
              const int l2 = kTileM * int(e.t) + q * 32 + px;
              const int y2 = l2 / kPitch, x2 = l2 - y2 * kPitch;
#ifdef FEN_EXP_NOSTORE
              if (x2 < kStripW && y2 < p.H && v4.x == 0x12345678u)
#else
              if (x2 < kStripW && y2 < p.H)
#endif
                st_global_128(outp + ((size_t(img) * p.H + y2) * p.W + x2) * kC + ch * 8, v4);
            }
            named_bar_sync(1 + q, 64);             // the staging may be overwritten again
          }
#else
#ifdef FEN_EXP_NOSTORE
          if (valid && o[0] == 0x12345678u && o[9] == 0x9abcdef0u) {
#else
          if (valid) {
#endif
            st_global_256(outp + opix * kC + col0, *reinterpret_cast<uint32_t(*)[8]>(&o[0]));
            st_global_256(outp + opix * kC + col0 + 16, *reinterpret_cast<uint32_t(*)[8]>(&o[8]));
          }
#endif
          if (ew == 0 && lane == 0) B2T2(P, 3, i);
          if (EPI == kBEpiPreluHsum) {
            // ---- the 9 channel sums of the bf16-ROUNDED h (what conv2 will read) that the SE pool needs
            if (valid) {
#pragma unroll
              for (int k = 0; k < 16; ++k) {
                csum[2 * k] += bf16lo(o[k]);
                csum[2 * k + 1] += bf16hi(o[k]);
              }
            }
#if defined(FEN_B2_X1) || defined(FEN_B2_X4)
            const unsigned mc0 = 0, mcl = 0;
#else
            // border columns: at most one pixel of each per warp and tile (a strip row is 66 > 32 pixels)
            const unsigned mc0 = __ballot_sync(0xffffffffu, valid && x == 0);
            const unsigned mcl = __ballot_sync(0xffffffffu, valid && x == p.W - 1);
#endif
#pragma unroll
            for (int side = 0; side < 2; ++side) {
              const unsigned mk = side ? mcl : mc0;
              if (mk) {
                const int src = __ffs(mk) - 1;
                if (lane == src) {
                  uint4* d = reinterpret_cast<uint4*>(&s_colx[ew][side][0]);
                  d[0] = make_uint4(o[0], o[1], o[2], o[3]);    d[1] = make_uint4(o[4], o[5], o[6], o[7]);
                  d[2] = make_uint4(o[8], o[9], o[10], o[11]);  d[3] = make_uint4(o[12], o[13], o[14], o[15]);
                }
                __syncwarp();
                const uint32_t wv = s_colx[ew][side][lane >> 1];
                const float val = (lane & 1) ? bf16hi(wv) : bf16lo(wv);
                const int ys = __shfl_sync(0xffffffffu, y, src);
                if (side) colLsum += val; else col0sum += val;
                if (ys == 0) hs_add(hs + (side ? kHsC0L : kHsC00) * kC + col0 + lane, val);
                if (ys == p.H - 1) hs_add(hs + (side ? kHsCL0 + 1 : kHsCL0) * kC + col0 + lane, val);
                __syncwarp();
              }
            }
            // border rows: only the first and the last tiles of an image contain them
#if defined(FEN_B2_X1) || defined(FEN_B2_X3)
            const unsigned mr0 = 0, mrl = 0;
#else
            const unsigned mr0 = __ballot_sync(0xffffffffu, valid && y == 0);
            const unsigned mrl = __ballot_sync(0xffffffffu, valid && y == p.H - 1);
#endif
#pragma unroll
            for (int side = 0; side < 2; ++side) {
              const unsigned mk = side ? mrl : mr0;
              if (mk) {
                const bool in = (mk >> lane) & 1u;
#pragma unroll
                for (int k = 0; k < 16; ++k) {
                  f[2 * k] = in ? bf16lo(o[k]) : 0.f;
                  f[2 * k + 1] = in ? bf16hi(o[k]) : 0.f;
                }
#pragma unroll
                for (int d = 16, len = CW; d >= 1; d >>= 1, len >>= 1) {
                  const bool hi = (lane & d) != 0;
#pragma unroll
                  for (int k = 0; k < len / 2; ++k) {
                    const float send = hi ? f[k] : f[k + len / 2];
                    const float keep = hi ? f[k + len / 2] : f[k];
                    f[k] = keep + __shfl_xor_sync(0xffffffffu, send, d);
                  }
                }
                hs_add(hs + (side ? kHsRowL : kHsRow0) * kC + col0 + lane, f[0]);
              }
            }
          }
        }
        flush_unit();
        };
        if (ly.epi == kBEpiPreluHsum) run_tiles(std::integral_constant<int, kBEpiPreluHsum>{});
        else if (ly.epi == kBEpiSeResidual) run_tiles(std::integral_constant<int, kBEpiSeResidual>{});
        else run_tiles(std::integral_constant<int, kBEpiResidual>{});
        // ---- pass done for this warp: make its global writes visible, then publish the CTA's flag.
        // (Measured without the per-thread fence - one cumulative st.release.gpu by the flag writer behind the mbarrier,
        // valid in the PTX memory model: 3.085 -> 3.020 ms at batch 64, but tools/soak2.py then caught a quarter tile of
        // stale input, 8 times in 12 000 forwards, on one box.  The fence stays.)
        if (ew == 0) B2TS(P, 5);                  // (trace: this warp's tiles of the pass are stored)
#ifndef FEN_B2_NO_WRITER_PROXY_FENCE
        // the stores above went through the generic proxy; the readers are TMA loads (async proxy) of this and other CTAs
        asm volatile("fence.proxy.async.global;" ::: "memory");
#endif
#ifndef FEN_B2_NO_PASS_FENCE
        __threadfence();
#endif
        __syncwarp();
        // arrivals of pass P may only start once pass P-1 is complete (a warp running ahead over short
        // passes could otherwise complete a phase with two of its own arrivals)
        if (P > 0) mbar_wait(&bar_done, (P - 1) & 1);
        if (ew == 7) B2TS(P, 7);                  // (trace: the last epilogue warp arrives, fences done)
        if (lane == 0) mbar_arrive(&bar_done);
        if (flag_writer) {
          B2W(4, 3, L, s, 0);
          mbar_wait(&bar_done, P & 1);
          B2TS(P, 6);                             // (trace: all eight warps have arrived)
          if (lane == 0) {
            st_release_gpu(p.flags + s * int(gridDim.x) + blockIdx.x, L + 1);
            B2TRACE(P, 7);
          }
          __syncwarp();
        }
      }
      if (ly.epi == kBEpiSeResidual) ++m_cnt;
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == kB2FirstMmaWarp) tmem_dealloc(tmem_base, kTmemCols);
}

}  // namespace fen
