// SSIM validation metric and SSIM loss (forward + gradient w.r.t. the prediction) as two tiled CUDA kernels.
// Reference: src/losses/ssim_loss.py:14-98 (create_gaussian_window, ssim) and :166-226 (SSIMLoss = 1 - ssim), used by
// Trainer._compute_ssim (src/training/trainer.py:630-634), evaluation/metrics.py:77 and the Stage-2 CombinedLoss
// (src/losses/combined.py:134-138).
//
//   window = g g^T, g = normalised 1-D Gaussian (size ws, sigma)          [the reference convolves with the 2-D product]
//   mu_p = w * p, mu_t = w * t, E_pp = w * p^2, E_tt = w * t^2, E_pt = w * (p t)      (zero padding ws / 2, per channel)
//   S = (2 mu_p mu_t + C1)(2 (E_pt - mu_p mu_t) + C2) / ((mu_p^2 + mu_t^2 + C1)(E_pp - mu_p^2 + E_tt - mu_t^2 + C2))
//   ssim = mean(S)
// The five filtered maps are never written: one CTA stages a (32 + ws - 1)^2 patch of p and t in shared memory, runs the
// separable filter (rows, then columns) on the five products and reduces S - one pass over p and t (HBM bound:
// 8 B per element read).  With a gradient requested it also writes the three maps the chain rule needs,
//   G_mu = dS/dmu_p, G_pp = dS/dE_pp, G_pt = dS/dE_pt,
// and a second kernel forms  d ssim / d p = 1/N [ w * G_mu + 2 p (w * G_pp) + t (w * G_pt) ]  (w is symmetric).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace fen {

constexpr int kSsimTile = 32;
constexpr int kSsimMaxWin = 11;
constexpr int kSsimPatch = kSsimTile + kSsimMaxWin - 1;     // 42
constexpr int kSsimPitch = kSsimPatch + 2;                   // 44 floats per staged row

struct SsimParams {
  int B, C, H, W;
  int ws;                       // odd window size <= 11
  float g[kSsimMaxWin];         // 1-D Gaussian, normalised (fp32, as create_gaussian_window computes it)
  float c1, c2;
  int tiles_x, tiles_y;         // per plane
};

// grid: (tiles_x * tiles_y, B * C).  partial[plane][tile] = sum of S over the tile (fixed order -> deterministic).
__global__ void __launch_bounds__(256)
ssim_fwd_kernel(const float* __restrict__ pred, const float* __restrict__ target, SsimParams P, float* __restrict__ partial,
                float* __restrict__ gmu, float* __restrict__ gpp, float* __restrict__ gpt) {
  __shared__ float sp[kSsimPatch][kSsimPitch], st[kSsimPatch][kSsimPitch];
  __shared__ float sh[5][kSsimPatch][kSsimTile + 1];       // row-filtered mu_p, mu_t, E_pp, E_tt, E_pt
  __shared__ float s_red[8], s_g[kSsimMaxWin];
  if (threadIdx.x < kSsimMaxWin) s_g[threadIdx.x] = P.g[threadIdx.x];
  const int tile = blockIdx.x, plane = blockIdx.y;
  const int ty = tile / P.tiles_x, tx = tile - ty * P.tiles_x;
  const int y0 = ty * kSsimTile, x0 = tx * kSsimTile, r = P.ws >> 1, patch = kSsimTile + P.ws - 1;
  const float* p = pred + size_t(plane) * P.H * P.W;
  const float* t = target + size_t(plane) * P.H * P.W;
  for (int i = threadIdx.x; i < patch * patch; i += 256) {
    const int py = i / patch, px = i - py * patch;
    const int y = y0 + py - r, x = x0 + px - r;
    const bool in = y >= 0 && y < P.H && x >= 0 && x < P.W;
    sp[py][px] = in ? __ldg(p + size_t(y) * P.W + x) : 0.f;
    st[py][px] = in ? __ldg(t + size_t(y) * P.W + x) : 0.f;
  }
  __syncthreads();
  // rows: every staged row, the 32 output columns
  for (int i = threadIdx.x; i < patch * kSsimTile; i += 256) {
    const int py = i / kSsimTile, ox = i - py * kSsimTile;
    float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f, a4 = 0.f;
    for (int k = 0; k < P.ws; ++k) {
      const float w = s_g[k], a = sp[py][ox + k], b = st[py][ox + k];
      a0 = fmaf(w, a, a0); a1 = fmaf(w, b, a1); a2 = fmaf(w, a * a, a2); a3 = fmaf(w, b * b, a3); a4 = fmaf(w, a * b, a4);
    }
    sh[0][py][ox] = a0; sh[1][py][ox] = a1; sh[2][py][ox] = a2; sh[3][py][ox] = a3; sh[4][py][ox] = a4;
  }
  __syncthreads();
  // columns + the SSIM map: 4 outputs per thread
  float acc = 0.f;
  for (int i = threadIdx.x; i < kSsimTile * kSsimTile; i += 256) {
    const int oy = i / kSsimTile, ox = i - oy * kSsimTile;
    const int y = y0 + oy, x = x0 + ox;
    if (y >= P.H || x >= P.W) continue;
    float m[5] = {0.f, 0.f, 0.f, 0.f, 0.f};
    for (int k = 0; k < P.ws; ++k) {
      const float w = s_g[k];
#pragma unroll
      for (int q = 0; q < 5; ++q) m[q] = fmaf(w, sh[q][oy + k][ox], m[q]);
    }
    const float mp = m[0], mt = m[1];
    const float vp = m[2] - mp * mp, vt = m[3] - mt * mt, cov = m[4] - mp * mt;
    const float A1 = 2.f * mp * mt + P.c1, A2 = 2.f * cov + P.c2;
    const float B1 = mp * mp + mt * mt + P.c1, B2 = vp + vt + P.c2;
    const float inv = 1.f / (B1 * B2);
    const float S = A1 * A2 * inv;
    acc += S;
    if (gmu) {
      const size_t o = size_t(plane) * P.H * P.W + size_t(y) * P.W + x;
      // dA1 = 2 mt, dA2 = -2 mt, dB1 = 2 mp, dB2 = -2 mp   (w.r.t. mu_p)
      gmu[o] = 2.f * mt * (A2 - A1) * inv - S * 2.f * mp * (1.f / B1 - 1.f / B2);
      gpp[o] = -S / B2;
      gpt[o] = 2.f * A1 * inv;
    }
  }
#pragma unroll
  for (int d = 16; d >= 1; d >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, d);
  if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float s = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) s += s_red[k];
    partial[size_t(plane) * gridDim.x + tile] = s;
  }
}

// per_image[b] = mean of S over the C planes of image b; mean[0] = mean over everything.  One block, fixed order.
__global__ void __launch_bounds__(256)
ssim_reduce_kernel(const float* __restrict__ partial, int B, int C, int tiles, double inv_chw, float* __restrict__ per_image,
                   float* __restrict__ mean) {
  __shared__ double sh[256];
  __shared__ double s_total;
  if (threadIdx.x == 0) s_total = 0.0;
  __syncthreads();
  for (int b = 0; b < B; ++b) {
    double a = 0.0;
    const float* src = partial + size_t(b) * C * tiles;
    for (int i = threadIdx.x; i < C * tiles; i += 256) a += double(src[i]);
    sh[threadIdx.x] = a;
    __syncthreads();
    for (int d = 128; d >= 1; d >>= 1) {
      if (threadIdx.x < d) sh[threadIdx.x] += sh[threadIdx.x + d];
      __syncthreads();
    }
    if (threadIdx.x == 0) {
      if (per_image) per_image[b] = float(sh[0] * inv_chw);
      s_total += sh[0];
    }
    __syncthreads();
  }
  if (threadIdx.x == 0 && mean) mean[0] = float(s_total * inv_chw / double(B));
}

// grad[b,c,y,x] = scale * ( (w * G_mu) + 2 p (w * G_pp) + t (w * G_pt) )      scale = 1 / (B C H W)
__global__ void __launch_bounds__(256)
ssim_bwd_kernel(const float* __restrict__ pred, const float* __restrict__ target, const float* __restrict__ gmu,
                const float* __restrict__ gpp, const float* __restrict__ gpt, SsimParams P, float scale,
                float* __restrict__ grad) {
  __shared__ float sg[3][kSsimPatch][kSsimPitch];
  __shared__ float sh[3][kSsimPatch][kSsimTile + 1];
  __shared__ float s_g[kSsimMaxWin];
  if (threadIdx.x < kSsimMaxWin) s_g[threadIdx.x] = P.g[threadIdx.x];
  const int tile = blockIdx.x, plane = blockIdx.y;
  const int ty = tile / P.tiles_x, tx = tile - ty * P.tiles_x;
  const int y0 = ty * kSsimTile, x0 = tx * kSsimTile, r = P.ws >> 1, patch = kSsimTile + P.ws - 1;
  const size_t base = size_t(plane) * P.H * P.W;
  for (int i = threadIdx.x; i < patch * patch; i += 256) {
    const int py = i / patch, px = i - py * patch;
    const int y = y0 + py - r, x = x0 + px - r;
    const bool in = y >= 0 && y < P.H && x >= 0 && x < P.W;
    const size_t o = base + size_t(in ? y : 0) * P.W + (in ? x : 0);
    sg[0][py][px] = in ? __ldg(gmu + o) : 0.f;
    sg[1][py][px] = in ? __ldg(gpp + o) : 0.f;
    sg[2][py][px] = in ? __ldg(gpt + o) : 0.f;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < patch * kSsimTile; i += 256) {
    const int py = i / kSsimTile, ox = i - py * kSsimTile;
    float a0 = 0.f, a1 = 0.f, a2 = 0.f;
    for (int k = 0; k < P.ws; ++k) {
      const float w = s_g[k];
      a0 = fmaf(w, sg[0][py][ox + k], a0); a1 = fmaf(w, sg[1][py][ox + k], a1); a2 = fmaf(w, sg[2][py][ox + k], a2);
    }
    sh[0][py][ox] = a0; sh[1][py][ox] = a1; sh[2][py][ox] = a2;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < kSsimTile * kSsimTile; i += 256) {
    const int oy = i / kSsimTile, ox = i - oy * kSsimTile;
    const int y = y0 + oy, x = x0 + ox;
    if (y >= P.H || x >= P.W) continue;
    float m0 = 0.f, m1 = 0.f, m2 = 0.f;
    for (int k = 0; k < P.ws; ++k) {
      const float w = s_g[k];
      m0 = fmaf(w, sh[0][oy + k][ox], m0); m1 = fmaf(w, sh[1][oy + k][ox], m1); m2 = fmaf(w, sh[2][oy + k][ox], m2);
    }
    const size_t o = base + size_t(y) * P.W + x;
    grad[o] = scale * (m0 + 2.f * __ldg(pred + o) * m1 + __ldg(target + o) * m2);
  }
}

}  // namespace fen
