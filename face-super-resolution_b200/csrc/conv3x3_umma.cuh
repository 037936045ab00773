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
// contiguous across the ring wrap).  warp 1: single-thread tcgen05.mma issuer, 9 taps x 4 k-steps
// per tile, fp32 accumulators double-buffered in TMEM.  warps 2-5: epilogue, TMEM -> registers ->
// bias / PReLU / residual / SE partial sums / PixelShuffle / bicubic skip -> global, written
// straight from registers so shared-memory bandwidth (the binding resource: every MMA re-reads
// its A and B operands) is left to the tensor core.
#pragma once
#include "fen_common.cuh"
#include "ptx_sm100.cuh"

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
struct ConvSmem {
  static constexpr int kWBytes = 9 * N * kC * 2;
  static constexpr int kDynBytes = kWBytes + kRingBytes + 1024;  // + alignment slack
};

__device__ __forceinline__ float bf16lo(uint32_t v) { return __uint_as_float(v << 16); }
__device__ __forceinline__ float bf16hi(uint32_t v) { return __uint_as_float(v & 0xffff0000u); }
__device__ __forceinline__ uint32_t pack_bf16(float a, float b) {
  __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}

// Cubic-convolution phase filters of F.interpolate(scale_factor=4, mode='bicubic',
// align_corners=False) (A = -0.75), times 2048; output d = 4q + r reads q + off[r] - 1 .. + 2.
__device__ __constant__ float c_bicubic_w[4][4] = {{-135.f / 2048.f, 873.f / 2048.f, 1535.f / 2048.f, -225.f / 2048.f},
                                                   {-21.f / 2048.f, 235.f / 2048.f, 1981.f / 2048.f, -147.f / 2048.f},
                                                   {-147.f / 2048.f, 1981.f / 2048.f, 235.f / 2048.f, -21.f / 2048.f},
                                                   {-225.f / 2048.f, 1535.f / 2048.f, 873.f / 2048.f, -135.f / 2048.f}};

template <int N>
__global__ void __launch_bounds__(kConvThreads, 1)
conv3x3_umma_kernel(const __grid_constant__ CUtensorMap tm_in, const __grid_constant__ CUtensorMap tm_w,
                    const ConvParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* w_smem = smem;                                   // [9][N][64] bf16, SWIZZLE_128B
  uint8_t* ring = smem + ConvSmem<N>::kWBytes;              // (kRingSlots + 1) x kSlotBytes
  __shared__ uint64_t bar_w, bar_full[kRingSlots], bar_empty[kRingSlots], bar_acc_full[2], bar_acc_empty[2];
  __shared__ uint32_t tmem_slot;
  __shared__ float s_bias[N], s_slope[kC];

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  constexpr uint32_t kTmemCols = (2 * N < 32) ? 32 : 2 * N;  // two accumulators

  const int g_begin = blockIdx.x * p.tiles_per_cta;
  const int g_end = min(p.total_tiles, g_begin + p.tiles_per_cta);

  if (warp == 1) tmem_alloc(&tmem_slot, kTmemCols);
  if (tid == 0) {
    mbar_init(&bar_w, 1);
    for (int i = 0; i < kRingSlots; ++i) { mbar_init(&bar_full[i], 1); mbar_init(&bar_empty[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&bar_acc_full[i], 1); mbar_init(&bar_acc_empty[i], 4); }
    fence_mbar_init();
    tma_prefetch_desc(&tm_in);
    tma_prefetch_desc(&tm_w);
  }
  for (int i = tid; i < N; i += kConvThreads) s_bias[i] = p.bias ? p.bias[blockIdx.y * N + i] : 0.f;
  for (int i = tid; i < kC; i += kConvThreads) s_slope[i] = p.slope ? p.slope[i] : 1.f;
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_slot;

  if (warp == 0) {
    // ============================================================ TMA producer
    if (lane == 0 && g_begin < g_end) {
      mbar_expect_tx(&bar_w, ConvSmem<N>::kWBytes);
      for (int tap = 0; tap < 9; ++tap)
        tma_load_2d(&tm_w, &bar_w, w_smem + tap * N * 128, 0, (blockIdx.y * 9 + tap) * N);
      uint32_t gb = 0;  // running box counter of this CTA
      for (int g = g_begin; g < g_end;) {
        const Unit u = make_unit(p, g, g_end);
        const int x0 = u.strip * kStripW - 1;
        for (int j = 0; j < u.nboxes; ++j, ++gb) {
          const uint32_t slot = gb % kRingSlots, ph = (gb / kRingSlots) & 1;
          mbar_wait(&bar_empty[slot], ph ^ 1);
          const bool mirror = (slot == kRingSlots - 1) && (j + 1 < u.nboxes);
          mbar_expect_tx(&bar_full[slot], mirror ? 2 * kSlotBytes : kSlotBytes);
          const int y0 = u.ra - 1 + j * kBoxRows;
          tma_load_4d(&tm_in, &bar_full[slot], smem_u32(ring + slot * kSlotBytes), 0, x0, y0, u.n);
          if (mirror)
            tma_load_4d(&tm_in, &bar_full[slot], smem_u32(ring + kRingSlots * kSlotBytes), 0, x0,
                        y0 + kBoxRows, u.n);
        }
        g += u.t1 - u.t0;
      }
    }
  } else if (warp == 1) {
    // ============================================================ MMA issuer (one thread)
    if (lane == 0 && g_begin < g_end) {
      constexpr uint32_t idesc = umma_idesc_bf16(kTileM, N);
      const uint32_t ring_u32 = smem_u32(ring), w_u32 = smem_u32(w_smem);
      mbar_wait(&bar_w, 0);
      uint32_t gb_base = 0, tile_ctr = 0;
      for (int g = g_begin; g < g_end;) {
        const Unit u = make_unit(p, g, g_end);
        int waited = 0, released = 0;
        for (int t = u.t0; t < u.t1; ++t, ++tile_ctr) {
          const uint32_t acc = tile_ctr & 1;
          mbar_wait(&bar_acc_empty[acc], ((tile_ctr >> 1) & 1) ^ 1);
          const int base = kTileM * t - kPitch * u.ra;
          const int need_last = min((base + kTileM + kMaxShift - 1) / kBoxPx, u.nboxes - 1);
          while (waited <= need_last) {
            const uint32_t gb = gb_base + waited;
            mbar_wait(&bar_full[gb % kRingSlots], (gb / kRingSlots) & 1);
            ++waited;
          }
          tc_fence_after();
          const uint32_t d_tmem = tmem_base + acc * N;
#pragma unroll
          for (int tap = 0; tap < 9; ++tap) {
            const int px = base + (tap / 3) * kPitch + (tap % 3);
            const int lb = px / kBoxPx, within = px - lb * kBoxPx;
            const uint32_t slot = (gb_base + lb) % kRingSlots;
            const uint32_t a_addr = ring_u32 + slot * kSlotBytes + within * 128;
            const uint32_t b_addr = w_u32 + tap * N * 128;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              const uint64_t ad = umma_smem_desc(a_addr + k * 32, 16, 1024, UMMA_LAYOUT_SW128);
              const uint64_t bd = umma_smem_desc(b_addr + k * 32, 16, 1024, UMMA_LAYOUT_SW128);
              umma_bf16_ss(d_tmem, ad, bd, idesc, (tap | k) != 0);
            }
          }
          // boxes that no later tile of this unit will read can go back to the producer
          const int next_first = (t + 1 < u.t1) ? (base + kTileM) / kBoxPx : u.nboxes;
          while (released < next_first) {
            umma_commit(&bar_empty[(gb_base + released) % kRingSlots]);
            ++released;
          }
          umma_commit(&bar_acc_full[acc]);
        }
        gb_base += u.nboxes;
        g += u.t1 - u.t0;
      }
    }
  } else {
    // ============================================================ epilogue (4 warps, 128 threads)
    const int q = warp & 3;            // TMEM lane quarter this warp may access
    const int row_in_tile = q * 32 + lane;
    uint32_t tile_ctr = 0;
    for (int g = g_begin; g < g_end;) {
      const Unit u = make_unit(p, g, g_end);
      float csum[(N == kC) ? kC : 1];
      if (N == kC) {
#pragma unroll
        for (int c = 0; c < ((N == kC) ? kC : 1); ++c) csum[c] = 0.f;
      }
      for (int t = u.t0; t < u.t1; ++t, ++tile_ctr) {
        const uint32_t acc = tile_ctr & 1;
        mbar_wait(&bar_acc_full[acc], (tile_ctr >> 1) & 1);
        tc_fence_after();
        uint32_t v[N];
        const uint32_t taddr = tmem_base + acc * N + (uint32_t(q * 32) << 16);
        if constexpr (N == 16) {
          tmem_ld_32x16(taddr, *reinterpret_cast<uint32_t(*)[16]>(&v[0]));
        } else {
#pragma unroll
          for (int h = 0; h < N / 32; ++h)
            tmem_ld_32x32(taddr + h * 32, *reinterpret_cast<uint32_t(*)[32]>(&v[h * 32]));
        }
        tmem_ld_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&bar_acc_empty[acc]);

        const int lin = kTileM * t + row_in_tile;       // strip-linear output pixel
        const int y = lin / kPitch, xs = lin - y * kPitch;
        const bool valid = (xs < kStripW) && (y < p.H);
        const int x = u.strip * kStripW + xs;

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
              p.out_f32[((size_t(u.n) * 3 + c) * p.H + y) * p.W + x] = o;
            }
          }
        } else {
          float f[N];
#pragma unroll
          for (int c = 0; c < N; ++c) f[c] = __uint_as_float(v[c]) + s_bias[c];
          if (p.epi == kEpiSum) {
            if (valid) {
#pragma unroll
              for (int c = 0; c < N; ++c) csum[c & ((N == kC) ? 63 : 0)] += f[c];
            }
          } else if (p.epi == kEpiPrelu || p.epi == kEpiShuffle) {
#pragma unroll
            for (int c = 0; c < N; ++c) f[c] = f[c] > 0.f ? f[c] : f[c] * s_slope[c];
          }
          if (valid) {
            size_t opix;
            if (p.epi == kEpiShuffle) {
              const int sub = blockIdx.y;
              opix = (size_t(u.n) * (2 * p.H) + (2 * y + (sub >> 1))) * (2 * p.W) + (2 * x + (sub & 1));
            } else {
              opix = (size_t(u.n) * p.H + y) * p.W + x;
            }
            uint4* dst = reinterpret_cast<uint4*>(p.out + opix * kC);
            if (p.epi == kEpiResidual) {
              const uint4* rsd = reinterpret_cast<const uint4*>(p.residual + opix * kC);
#pragma unroll
              for (int j = 0; j < N / 8; ++j) {
                const uint4 r = __ldg(rsd + j);
                f[8 * j + 0] += bf16lo(r.x); f[8 * j + 1] += bf16hi(r.x);
                f[8 * j + 2] += bf16lo(r.y); f[8 * j + 3] += bf16hi(r.y);
                f[8 * j + 4] += bf16lo(r.z); f[8 * j + 5] += bf16hi(r.z);
                f[8 * j + 6] += bf16lo(r.w); f[8 * j + 7] += bf16hi(r.w);
              }
            }
#pragma unroll
            for (int j = 0; j < N / 8; ++j) {
              uint4 o;
              o.x = pack_bf16(f[8 * j + 0], f[8 * j + 1]);
              o.y = pack_bf16(f[8 * j + 2], f[8 * j + 3]);
              o.z = pack_bf16(f[8 * j + 4], f[8 * j + 5]);
              o.w = pack_bf16(f[8 * j + 6], f[8 * j + 7]);
              dst[j] = o;
            }
          }
        }
      }
      if constexpr (N == kC) {
        if (p.epi == kEpiSum) {
          // per-image channel sums for the squeeze-and-excitation pool: warp reduce, then atomics
#pragma unroll
          for (int c = 0; c < kC; ++c) {
            float s = csum[c];
            s += __shfl_xor_sync(0xffffffffu, s, 16);
            s += __shfl_xor_sync(0xffffffffu, s, 8);
            s += __shfl_xor_sync(0xffffffffu, s, 4);
            s += __shfl_xor_sync(0xffffffffu, s, 2);
            s += __shfl_xor_sync(0xffffffffu, s, 1);
            if (lane == (c & 31)) atomicAdd(p.sums + size_t(u.n) * kC + c, s);
          }
        }
      }
      g += u.t1 - u.t0;
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, kTmemCols);
}

}  // namespace fen
