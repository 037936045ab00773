// Host orchestration of the Stage-1 training step, network side: a forward pass that KEEPS what the backward
// needs (per-layer launches of the tcgen05 convolution kernel) and the backward pass itself.
// Included by fen_b200.cu inside namespace fen (uses its Layout / ConvArgs / launch_conv helpers).
#pragma once

// ---------------------------------------------------------------- transposed weights for the dgrad convolutions
struct BwdLayout {
  int64_t rcab0, rcab_stride;   // per RCAB: [w1T][w2T]
  int64_t gconv0;               // per group: [wT]
  int64_t after, up[2], last, zeros, total;
};
static void make_bwd_layout(const Layout& L, BwdLayout* K) {
  int64_t o = 0;
  K->rcab0 = o; K->rcab_stride = 2 * kConvWBytes; o += int64_t(L.n_rcab) * K->rcab_stride;
  K->gconv0 = o; o += int64_t(L.G) * kConvWBytes;
  K->after = o; o += kConvWBytes;
  for (int s = 0; s < 2; ++s) { K->up[s] = o; o += 4 * kConvWBytes; }
  K->last = o; o += kConvWBytes;            // conv_last's data gradient as a 64 -> 64 convolution (pack_last_T_kernel)
  K->zeros = o; o += 256;
  K->total = o;
}

// records 0 .. 2 n_rcab - 1: conv1 / conv2 of every RCAB, then the G group convolutions, then conv_after_body
__global__ void pack_T_all_kernel(const float* __restrict__ params, uint8_t* __restrict__ kb, Layout L, BwdLayout K) {
  const int i = blockIdx.y;
  const float* w;
  uint8_t* dst;
  if (i < 2 * L.n_rcab) {
    const int r = i >> 1, g = r / L.Bk, b = r % L.Bk;
    w = params + L.p_rcab0 + g * L.p_group_stride + b * L.p_rcab_stride + ((i & 1) ? kConvW + 64 + 64 : 0);
    dst = kb + K.rcab0 + int64_t(r) * K.rcab_stride + ((i & 1) ? kConvWBytes : 0);
  } else if (i < 2 * L.n_rcab + L.G) {
    const int g = i - 2 * L.n_rcab;
    w = params + L.p_rcab0 + g * L.p_group_stride + L.p_gconv_w_in_group;
    dst = kb + K.gconv0 + g * kConvWBytes;
  } else {
    w = params + L.p_after_w;
    dst = kb + K.after;
  }
  pack_conv64_dev(w, reinterpret_cast<bf16*>(dst), true);
}

// batched weight-gradient launches: after group G / 2 (the upper half, with conv_after_body), after group 1 and after
// group 0 (a small last batch keeps the gradient slice that finishes last - and whose all-reduce is exposed - small)
static bool bwd_is_flush(const Layout& L, int g) { return g == L.G / 2 || g == 1 || g == 0; }
static int bwd_prev_flush(const Layout& L, int g) {       // the flush group above g, or -1
  for (int q = g + 1; q < L.G; ++q)
    if (bwd_is_flush(L, q)) return q;
  return -1;
}
// ---------------------------------------------------------------- saved activations + gradient buffers
// All [B][H][W][64] bf16 tensors of the step lie at ONE stride (`act`) from the workspace base, so that a single 5-D
// tensor map addresses any of them by index (body2_umma_kernel<true>, wgrad_batch_umma_kernel):
//   0 f0 | 1 body | xs[n] (x' of every RCAB) | h[n] | o[n] | gout[G] |            <- forward, kept for the backward
//   dO[n] | dH[n] (the output gradients of conv2 / conv1 of every RCAB) | dG[G] (d gout[g]) | dBody | dF | dX0 | dX1
struct StepWs {
  int64_t act;                      // bytes of one [B][H][W][64] bf16 tensor
  int64_t f0, body, xs0, h0, o0, gout0;
  int64_t dO0, dH0, dG0, dBody, dF, dX0, dX1;
  int n_act;                        // buffers in the uniformly strided region
  int64_t u0, u1, sums;
  int64_t m_h0, m_stride, m_u0, m_u1;                     // PReLU sign masks (8 B per pixel): per RCAB conv1, the two stages
  int64_t hsum, flags;                                    // fused forward (body2_umma_kernel<true>)
  int64_t dy1, du0, dy0, dsum;                            // backward, upsample resolution
  int64_t d8, x8, wg3;                                    // backward, the 3-channel ends (narrow tensors + scratch)
  int64_t g64, wg_part;                                   // deterministic reductions: fixed-point shadow, partial weight gradients
  int64_t total;
};
static void make_step_ws(const Layout& L, int B, int H, int W, StepWs* w) {
  const int64_t act = align256(int64_t(B) * H * W * 64 * 2);
  w->act = act;
  int64_t o = 0;
  w->f0 = o; o += act;
  w->body = o; o += act;
  w->xs0 = o; o += act * L.n_rcab;
  w->h0 = o; o += act * L.n_rcab;
  w->o0 = o; o += act * L.n_rcab;
  w->gout0 = o; o += act * L.G;
  w->dO0 = o; o += act * L.n_rcab;
  w->dH0 = o; o += act * L.n_rcab;
  w->dG0 = o; o += act * L.G;
  w->dBody = o; o += act;
  w->dF = o; o += act;
  w->dX0 = o; o += act;
  w->dX1 = o; o += act;
  w->n_act = int(o / act);
  w->u0 = o; o += 4 * act;
  w->u1 = o; o += 16 * act;
  w->sums = o; o += align256(int64_t(L.n_rcab) * B * 64 * 8);   // fixed-point SE pool sums (hs_add)
  w->m_stride = align256(int64_t(B) * H * W * 8);
  w->m_h0 = o; o += w->m_stride * L.n_rcab;
  w->m_u0 = o; o += 4 * w->m_stride;
  w->m_u1 = o; o += 16 * w->m_stride;
  w->hsum = o; o += align256(int64_t(L.n_rcab) * B * 9 * 64 * 8);
  w->flags = o; o += 4096;
  w->dy1 = o; o += 16 * act;
  w->du0 = o; o += 4 * act;
  w->dy0 = o; o += 4 * act;
  w->dsum = o; o += align256(int64_t(L.n_rcab) * B * 64 * 8);   // sum_px dx' * o per RCAB, image, channel (fixed point, gs_add)
  w->g64 = o; o += align256(L.p_total * 8);                      // fixed-point shadow of the flat gradient (small tensors only)
  {
    // partial weight gradients (wgrad_reduce_kernel): the largest batched launch, or one partial per SM
    int max_jobs = 1;
    for (int g = L.G - 1; g >= 0; --g)
      if (bwd_is_flush(L, g)) {
        const int prev = bwd_prev_flush(L, g), per = 1 + 2 * L.Bk;
        const int jb = prev < 0 ? 0 : 1 + (L.G - prev) * per, je = 1 + (L.G - g) * per;
        if (je - jb > max_jobs) max_jobs = je - jb;
      }
    int64_t parts = int64_t(max_jobs) * 16;
    if (parts < num_sms()) parts = num_sms();
    w->wg_part = o; o += align256(parts * kWgPartFloats * 4);
  }
  w->d8 = o; o += 2 * act;                                       // d out as bf16 [B][4H][4W][8] (nchw3_to_nhwc8_kernel)
  w->x8 = o; o += align256(int64_t(B) * H * W * 8 * 2);          // the LR input likewise
  w->wg3 = o; o += align256((64 * 576 + 64) * 4);        // [64][576] + [64] fp32: weight-gradient scratch of the 3-channel ends
  w->total = o;
}

static int conv64(const bf16* in, const void* w, const float* bias, const float* slope, const bf16* res, float* sm,
                  bf16* out, int epi, int B, int h, int w_, cudaStream_t st, const bf16* aux = nullptr,
                  uint32_t* mask = nullptr, long long* sm64 = nullptr, int in_chans = 0, int unshuffle = 0) {
  ConvArgs a{};
  a.in_chans = in_chans; a.p.unshuffle = unshuffle;
  a.x = in; a.w = w; a.n = 64; a.groups = (epi == kEpiShuffle) ? 4 : 1;
  a.p.B = B; a.p.H = h; a.p.W = w_; a.p.epi = epi; a.p.bias = bias; a.p.slope = slope; a.p.residual = res;
  a.p.out = out; a.p.sums = sm; a.p.sums64 = sm64; a.p.aux = aux;
  if (epi == kEpiGate) a.p.mask_in = mask; else a.p.mask_out = mask;
  return launch_conv(a, st);
}

static int ew_blocks(size_t items) {
  size_t b = (items + 255) / 256;
  const size_t cap = size_t(num_sms()) * 8;
  if (b > cap) b = cap;
  if (b < 1) b = 1;
  return int(b);
}

// Forward in train() mode (no clamp, custom.py:187) that keeps every RCAB's x / h / o, the group outputs and both
// upsample stages for the backward pass.
static int step_forward(const fen_config* cfg, const Layout& L, const uint8_t* k, const float* x, float* out, int B,
                        int H, int W, uint8_t* wsb, const StepWs& ws, cudaStream_t st) {
  int rc;
  const RcabRec rr = rcab_rec(L.R);
  auto act = [&](int64_t off) { return reinterpret_cast<bf16*>(wsb + off); };
  long long* sums = reinterpret_cast<long long*>(wsb + ws.sums);
  conv_first_kernel<<<dim3(H, B), 256, 0, st>>>(x, reinterpret_cast<const float*>(k + L.k_first_w),
                                                reinterpret_cast<const float*>(k + L.k_first_b), act(ws.f0), H, W);
  FEN_CUDA(cudaGetLastError());
  ++g_launches;
  const int hw = H * W;
  const bf16* cur = act(ws.f0);
  const bool fused = body2_usable(L, B, H, W);
  if (fused) {
    // the whole residual body in ONE persistent launch that also keeps what the backward needs: every RCAB's x', h
    // and o, the PReLU sign masks, the SE pool sums (body2_umma_kernel<true>; StepWs buffer order = body2_layer's)
    Body2Bufs bufs{};
    bufs.act_base = wsb + ws.f0; bufs.act_stride = ws.act; bufs.nbuf = 2 + 3 * L.n_rcab + L.G;
    bufs.hsum64 = reinterpret_cast<long long*>(wsb + ws.hsum); bufs.flags = reinterpret_cast<int*>(wsb + ws.flags);
    bufs.mask0 = reinterpret_cast<uint32_t*>(wsb + ws.m_h0); bufs.mask_stride_bytes = ws.m_stride;
    bufs.pool_sums = sums;
    if ((rc = launch_body2<true>(cfg, L, bufs, k, B, H, W, nullptr, st))) return rc;
  }
  if (!fused) FEN_CUDA(cudaMemsetAsync(sums, 0, size_t(L.n_rcab) * B * 64 * 8, st));
  for (int g = 0; g < (fused ? 0 : L.G); ++g) {
    const bf16* gin = cur;
    for (int b = 0; b < L.Bk; ++b) {
      const int r = g * L.Bk + b;
      const uint8_t* kr = k + L.k_rcab0 + int64_t(r) * L.k_rcab_stride;
      long long* sm = sums + size_t(r) * B * 64;
      bf16* h = act(ws.h0 + r * ws.act);
      bf16* o = act(ws.o0 + r * ws.act);
      bf16* nxt = act(ws.xs0 + r * ws.act);
      if ((rc = conv64(cur, kr + rr.w1, reinterpret_cast<const float*>(kr + rr.b1),
                       reinterpret_cast<const float*>(kr + rr.slope), nullptr, nullptr, h, kEpiPrelu, B, H, W, st,
                       nullptr, reinterpret_cast<uint32_t*>(wsb + ws.m_h0 + r * ws.m_stride))))
        return rc;
      if ((rc = conv64(h, kr + rr.w2, reinterpret_cast<const float*>(kr + rr.b2), nullptr, nullptr, nullptr, o, kEpiSum, B,
                       H, W, st, nullptr, nullptr, sm)))
        return rc;
      se_residual_kernel<<<dim3(32, B), 256, 0, st>>>(cur, o, sm, reinterpret_cast<const float*>(kr + rr.fc0),
                                                      reinterpret_cast<const float*>(kr + rr.fc2), L.R,
                                                      1.f / float(hw), cfg->res_scale, nxt, nullptr, 0, hw);
      FEN_CUDA(cudaGetLastError());
      ++g_launches;
      cur = nxt;
    }
    const uint8_t* kg = k + L.k_gconv0 + g * L.k_gconv_stride;
    bf16* gout = act(ws.gout0 + g * ws.act);
    if ((rc = conv64(cur, kg, reinterpret_cast<const float*>(kg + kConvWBytes), nullptr, gin, nullptr, gout,
                     kEpiResidual, B, H, W, st)))
      return rc;
    cur = gout;
  }
  if (!fused &&
      (rc = conv64(cur, k + L.k_after, reinterpret_cast<const float*>(k + L.k_after + kConvWBytes), nullptr,
                   act(ws.f0), nullptr, act(ws.body), kEpiResidual, B, H, W, st)))
    return rc;
  const uint8_t* ku = k + L.k_up[0];
  if ((rc = conv64(act(ws.body), ku, reinterpret_cast<const float*>(ku + 4 * kConvWBytes),
                   reinterpret_cast<const float*>(ku + 4 * kConvWBytes + 1024), nullptr, nullptr, act(ws.u0),
                   kEpiShuffle, B, H, W, st, nullptr, reinterpret_cast<uint32_t*>(wsb + ws.m_u0))))
    return rc;
  ku = k + L.k_up[1];
  if ((rc = conv64(act(ws.u0), ku, reinterpret_cast<const float*>(ku + 4 * kConvWBytes),
                   reinterpret_cast<const float*>(ku + 4 * kConvWBytes + 1024), nullptr, nullptr, act(ws.u1),
                   kEpiShuffle, B, 2 * H, 2 * W, st, nullptr, reinterpret_cast<uint32_t*>(wsb + ws.m_u1))))
    return rc;
  return launch_conv_last(k + L.k_last, act(ws.u1), x, out, nullptr, 0, 1, B, 4 * H, 4 * W, st);
}

static int wgrad64(const bf16* dY, const bf16* X, float* dW, float* dB, int B, int H, int W, int co_mul, int co_off,
                   cudaStream_t st, float* parts, int y_chans = kC, int x_chans = kC) {
  const int bands = B * H * ((W + kStripW - 1) / kStripW);
  const int grid = bands < num_sms() ? bands : num_sms();
  // The product library carries the tcgen05 kernel (wgrad_umma.cuh).  Developer builds (-DFEN_DEV) also hold the
  // two earlier generations for A/B runs: FEN_WGRAD=0 fp32 FMA on the CUDA cores, 1 warp-level mma.sync.
  int version = 2;
#ifdef FEN_DEV
  version = env_int("FEN_WGRAD", 2);
  if (version == 0) {
    wgrad_c64_kernel<<<grid, 256, 0, st>>>(dY, X, dW, dB, B, H, W, co_mul, co_off);
  } else if (version == 1) {
    FEN_CUDA(ensure_smem_attr(kKWgMma, wgrad_c64_mma_kernel, kWgDynBytes));
    wgrad_c64_mma_kernel<<<grid, 256, kWgDynBytes, st>>>(dY, X, dW, dB, B, H, W, co_mul, co_off);
  }
#endif
  if (version == 2) {
    FEN_CUDA(ensure_smem_attr(kKWgUmma, wgrad_c64_umma_kernel, kWuDynBytes));
    CUtensorMap tm_y, tm_x;
    int rc = make_act_map(&tm_y, dY, B, H, W, 1, kStripW, y_chans);
    if (rc) return rc;
    if ((rc = make_act_map(&tm_x, X, B, H, W, 3, kPitch, x_chans))) return rc;
    // every CTA leaves its partial gradient in `parts`; they are added up in CTA order (deterministic)
    wgrad_c64_umma_kernel<<<grid, kWuThreads, kWuDynBytes, st>>>(tm_y, tm_x, parts, B, H, W);
    FEN_CUDA(cudaGetLastError());
    ++g_launches;
    wgrad_reduce_kernel<<<37, 256, 0, st>>>(parts, grid, dW, dB, co_mul, co_off);
  }
  FEN_CUDA(cudaGetLastError());
  ++g_launches;
  return FEN_OK;
}

// Backward of step_forward: dout [B,3,4H,4W] fp32 (d loss / d network output) -> grads (flat fp32, the layout of
// the flat parameter vector, zeroed by stage 0).
//
// Weight gradients of the 64 -> 64 body convolutions are DEFERRED: the chain of data-gradient convolutions keeps every
// layer's output gradient (dO / dH per RCAB, dG per group, dBody), and wgrad_batch_umma_kernel computes all weight
// gradients of a run of groups in one persistent launch (bwd_is_flush: after the upper half, after group 1, after group 0).
//
// The pass is cut into G + 2 STAGES, in the order the backward runs, so that a data-parallel trainer can all-reduce a
// finished slice of the flat gradient while later stages still compute:
//   stage 0        conv_last, both upsample stages, data gradient of conv_after_body
//   stage 1 + k    residual group G - 1 - k   (+ a batched weight-gradient launch where bwd_is_flush)
//   stage G + 1    long skip + conv_first
// fen_backward_stage_range gives the slice a stage COMPLETES (empty for a stage that only feeds a later batch).
// All state between stages lives in the step workspace.  [stage_begin, stage_end) runs a sub-range.
static int bwd_num_stages(const Layout& L) { return L.G + 2; }
static void bwd_stage_range(const Layout& L, int stage, int64_t* begin, int64_t* count) {
  *begin = 0; *count = 0;
  if (stage == 0) { *begin = L.p_up[0]; *count = L.p_total - L.p_up[0]; }
  else if (stage <= L.G) {
    const int g = L.G - stage;
    if (bwd_is_flush(L, g)) {
      const int prev = bwd_prev_flush(L, g);
      const int64_t end = prev < 0 ? L.p_up[0] : L.p_rcab0 + int64_t(prev) * L.p_group_stride;
      *begin = L.p_rcab0 + int64_t(g) * L.p_group_stride;
      *count = end - *begin;
    }
  } else { *begin = 0; *count = L.p_rcab0; }
}

// 5-D maps over the workspace's activation-strided region for the batched weight gradient
static int make_act5_map(CUtensorMap* m, const void* base, int64_t buf_stride_bytes, int nbuf, int B, int H, int W,
                         int box_px, int box_rows);

static int wgrad_batch(const Layout& L, uint8_t* wsb, const StepWs& ws, float* grads, int B, int H, int W, int job_begin,
                       int job_end, cudaStream_t st) {
  if (job_end <= job_begin) return FEN_OK;
  FEN_CUDA(ensure_smem_attr(kKWgBatch, wgrad_batch_umma_kernel, kWbDynBytes));
  CUtensorMap tm_y, tm_x;
  int rc = make_act5_map(&tm_y, wsb, ws.act, ws.n_act, B, H, W, kStripW, 1);
  if (rc) return rc;
  if ((rc = make_act5_map(&tm_x, wsb, ws.act, ws.n_act, B, H, W, kPitch, 3))) return rc;
  WgBatchParams p{};
  p.B = B; p.H = H; p.W = W; p.G = L.G; p.Bk = L.Bk;
  p.job_begin = job_begin; p.job_end = job_end;
  const int bands = B * H * ((W + kStripW - 1) / kStripW);
  p.chunks = bands >= 16 * 32 ? 16 : (bands >= 64 ? bands / 32 : 1);     // >= 32 bands per work item
  p.i_xs0 = int(ws.xs0 / ws.act); p.i_h0 = int(ws.h0 / ws.act); p.i_gout0 = int(ws.gout0 / ws.act);
  p.i_dO0 = int(ws.dO0 / ws.act); p.i_dH0 = int(ws.dH0 / ws.act); p.i_dG0 = int(ws.dG0 / ws.act);
  p.i_dBody = int(ws.dBody / ws.act);
  p.p_rcab0 = L.p_rcab0; p.p_rcab_stride = L.p_rcab_stride; p.p_group_stride = L.p_group_stride;
  p.p_gconv_w_in_group = L.p_gconv_w_in_group; p.p_after_w = L.p_after_w;
  p.grads = grads;
  p.parts = reinterpret_cast<float*>(wsb + ws.wg_part);
  const int items = (job_end - job_begin) * p.chunks;
  const int grid = items < num_sms() ? items : num_sms();
  wgrad_batch_umma_kernel<<<grid, kWbThreads, kWbDynBytes, st>>>(tm_y, tm_x, p);
  FEN_CUDA(cudaGetLastError());
  ++g_launches;
  wgrad_reduce_batch_kernel<<<dim3(10, job_end - job_begin), 256, 0, st>>>(p);     // chunk partials -> flat gradient, fixed order
  FEN_CUDA(cudaGetLastError());
  ++g_launches;
  return FEN_OK;
}

static int step_backward(const fen_config* cfg, const Layout& L, const uint8_t* k, const uint8_t* kb,
                         const BwdLayout& K, const float* x, const float* dout, float* grads, int B, int H, int W,
                         uint8_t* wsb, const StepWs& ws, cudaStream_t st, int stage_begin, int stage_end) {
  int rc;
  const RcabRec rr = rcab_rec(L.R);
  auto act = [&](int64_t off) { return reinterpret_cast<bf16*>(wsb + off); };
  const float* zeros = reinterpret_cast<const float*>(kb + K.zeros);
  const long long* sums = reinterpret_cast<const long long*>(wsb + ws.sums);
  long long* dsum = reinterpret_cast<long long*>(wsb + ws.dsum);     // fixed point (gs_add), like every small reduction below
  long long* g64 = reinterpret_cast<long long*>(wsb + ws.g64);       // shadow of `grads` for the slope / SE-matrix gradients
  float* parts = reinterpret_cast<float*>(wsb + ws.wg_part);
  const int Ho = 4 * H, Wo = 4 * W;
  const size_t n8 = size_t(B) * H * W * 8;   // 8-element groups of one body-resolution tensor
  bf16* dBody = act(ws.dBody);
  auto dG = [&](int g) { return act(ws.dG0 + g * ws.act); };      // d gout[g]
  const int per = 1 + 2 * L.Bk;              // deferred weight-gradient jobs per group (wgrad_umma.cuh: wg_job)
  if (stage_begin <= 0 && stage_end > 0) {
  FEN_CUDA(cudaMemsetAsync(grads, 0, size_t(L.p_total) * 4, st));
  FEN_CUDA(cudaMemsetAsync(dsum, 0, size_t(L.n_rcab) * B * 64 * 8, st));
  FEN_CUDA(cudaMemsetAsync(g64, 0, size_t(L.p_total) * 8, st));

  // ---- conv_last (64 -> 3) on the 64-channel tcgen05 kernels: d out goes to bf16 NHWC with 8 channels, which a narrow
  // tensor map zero-extends to 64 (make_act_map).  Weight / bias gradient = rows 0..2 of a 64 x 64 weight gradient;
  // data gradient = a 64 -> 64 convolution with the transposed weights whose epilogue is the backward of stage 1's
  // PReLU (kEpiGate) and PixelShuffle (unshuffle).  (The CUDA-core kernels these replace took 0.5 + 0.66 ms at batch 32.)
  {
    bf16* d8 = act(ws.d8);
    float* wg = reinterpret_cast<float*>(wsb + ws.wg3);
    nchw3_to_nhwc8_kernel<<<ew_blocks(size_t(B) * Ho * Wo), 256, 0, st>>>(dout, d8, B, size_t(Ho) * Wo);
    FEN_CUDA(cudaGetLastError());
    ++g_launches;
    if ((rc = wgrad64(d8, act(ws.u1), wg, wg + 64 * 576, B, Ho, Wo, 1, 0, st, parts, 8, kC))) return rc;
    FEN_CUDA(cudaMemcpyAsync(grads + L.p_last_w, wg, 3 * 576 * 4, cudaMemcpyDeviceToDevice, st));
    FEN_CUDA(cudaMemcpyAsync(grads + L.p_last_b, wg + 64 * 576, 3 * 4, cudaMemcpyDeviceToDevice, st));
    const float* slope1 = reinterpret_cast<const float*>(k + L.k_up[1] + 4 * kConvWBytes + 1024);
    if ((rc = conv64(d8, kb + K.last, zeros, slope1, act(ws.u1), nullptr, act(ws.dy1), kEpiGate, B, Ho, Wo, st, nullptr,
                     reinterpret_cast<uint32_t*>(wsb + ws.m_u1), g64 + L.p_up[1] + 4 * kConvW + 256, 8, 1)))
      return rc;
  }
  // ---- upsample stage 1 (conv 64 -> 256 on 2H x 2W, input u0): 4 sub-pixel planes
  {
    const int h2 = 2 * H, w2 = 2 * W;
    const int64_t plane = 4 * ws.act;   // one [B][2H][2W][64] plane of dy1
    bf16* acc[2] = {act(ws.du0), act(ws.dy0)};   // ping-pong; dy0 is free until the unshuffle below
    for (int sub = 0; sub < 4; ++sub) {
      const bf16* dy = act(ws.dy1 + sub * plane);
      if ((rc = wgrad64(dy, act(ws.u0), grads + L.p_up[1], grads + L.p_up[1] + 4 * kConvW, B, h2, w2, 4, sub, st, parts)))
        return rc;
      if ((rc = conv64(dy, kb + K.up[1] + sub * kConvWBytes, zeros, nullptr, sub ? acc[(sub - 1) & 1] : nullptr,
                       nullptr, acc[sub & 1], sub ? kEpiResidual : kEpiBias, B, h2, w2, st)))
        return rc;
    }
    // result in acc[1] = dy0 region -> PReLU + PixelShuffle backward of stage 0 writes the planes into du0 region
    const float* slope0 = reinterpret_cast<const float*>(k + L.k_up[0] + 4 * kConvWBytes + 1024);
    prelu_bwd_kernel<<<ew_blocks(size_t(B) * h2 * w2 * 8), 256, 0, st>>>(
        acc[1], act(ws.u0), reinterpret_cast<const uint32_t*>(wsb + ws.m_u0), slope0, acc[0],
        g64 + L.p_up[0] + 4 * kConvW + 256, B, h2, w2, 1);
    FEN_CUDA(cudaGetLastError());
    ++g_launches;
  }
  // ---- upsample stage 0 (conv 64 -> 256 on H x W, input body): planes in du0 region, result -> dBody
  {
    bf16* acc[2] = {act(ws.dX0), dBody};
    for (int sub = 0; sub < 4; ++sub) {
      const bf16* dy = act(ws.du0 + sub * ws.act);
      if ((rc = wgrad64(dy, act(ws.body), grads + L.p_up[0], grads + L.p_up[0] + 4 * kConvW, B, H, W, 4, sub, st, parts)))
        return rc;
      if ((rc = conv64(dy, kb + K.up[0] + sub * kConvWBytes, zeros, nullptr, sub ? acc[(sub - 1) & 1] : nullptr,
                       nullptr, acc[sub & 1], sub ? kEpiResidual : kEpiBias, B, H, W, st)))
        return rc;
    }
  }
  // ---- conv_after_body + long skip: body = conv(gout[G-1]) + f0.  (weight gradient: job 0 of the first batch)
  if ((rc = conv64(dBody, kb + K.after, zeros, nullptr, nullptr, nullptr, dG(L.G - 1), kEpiBias, B, H, W, st))) return rc;
  // the two upsample PReLU slope gradients: fixed point -> fp32
  grads_from_fixed_kernel<<<dim3(1, 1), 128, 0, st>>>(g64, grads, L.p_up[0] + 4 * kConvW + 256, 0, 0, 64,
                                                       L.p_up[1] - L.p_up[0], 64);
  FEN_CUDA(cudaGetLastError());
  ++g_launches;
  }   // stage 0
  // ---- residual groups, last to first
  const int hw = H * W;
  for (int g = L.G - 1; g >= 0; --g) {
    const int stage = L.G - g;
    if (stage < stage_begin || stage >= stage_end) continue;
    bf16* dCur = dG(g);
    bf16* dX = act(ws.dX0);        // (re)written from dCur at the start of every group: the roles need no carry-over
    bf16* dXn = act(ws.dX1);
    // group conv: d blocks_out, and with it sum_px dx' * o of the group's last RCAB (kEpiDot)
    if ((rc = conv64(dCur, kb + K.gconv0 + g * kConvWBytes, zeros, nullptr, nullptr, nullptr, dX, kEpiDot, B, H, W, st,
                     act(ws.o0 + (g * L.Bk + L.Bk - 1) * ws.act), nullptr, dsum + size_t(g * L.Bk + L.Bk - 1) * B * 64)))
      return rc;
    for (int b = L.Bk - 1; b >= 0; --b) {
      const int r = g * L.Bk + b;
      const uint8_t* kr = k + L.k_rcab0 + int64_t(r) * L.k_rcab_stride;
      long long* pr = g64 + L.p_rcab0 + g * L.p_group_stride + b * L.p_rcab_stride;   // (fixed-point shadow: gs_add)
      long long* d_sl = pr + kConvW + 64;
      long long* d_fc0 = d_sl + 64 + kConvW + 64;
      long long* d_fc2 = d_fc0 + L.R * 64;
      const bf16* h = act(ws.h0 + r * ws.act);
      bf16* dO = act(ws.dO0 + r * ws.act);      // kept: the operands of the deferred weight gradients
      bf16* dH = act(ws.dH0 + r * ws.act);
      // squeeze-and-excitation + scaled residual (the per-image sums of dx' * o came with dX)
      FEN_CUDA(launch_pdl(se_bwd_apply_kernel, dim3(32, B), dim3(256), size_t(L.R * 65 + 64 * (L.R + 1)) * sizeof(float), st,
                          dX, sums + size_t(r) * B * 64, dsum + size_t(r) * B * 64,
                          reinterpret_cast<const float*>(kr + rr.fc0), reinterpret_cast<const float*>(kr + rr.fc2), L.R,
                          1.f / float(hw), cfg->res_scale, dO, d_fc0, d_fc2, hw));
      FEN_CUDA(cudaGetLastError());
      ++g_launches;
      // conv2: data gradient with the PReLU backward in its epilogue (dH holds dA)
      if ((rc = conv64(dO, kb + K.rcab0 + r * K.rcab_stride + kConvWBytes, zeros,
                       reinterpret_cast<const float*>(kr + rr.slope), h, nullptr, dH, kEpiGate, B, H, W, st, nullptr,
                       reinterpret_cast<uint32_t*>(wsb + ws.m_h0 + r * ws.m_stride), d_sl)))
        return rc;
      // conv1 + the identity path of the RCAB; the result is dx' of the previous RCAB of the group
      if (b > 0) {
        if ((rc = conv64(dH, kb + K.rcab0 + r * K.rcab_stride, zeros, nullptr, dX, nullptr, dXn, kEpiDot, B, H, W, st,
                         act(ws.o0 + (r - 1) * ws.act), nullptr, dsum + size_t(r - 1) * B * 64)))
          return rc;
      } else {
        if ((rc = conv64(dH, kb + K.rcab0 + r * K.rcab_stride, zeros, nullptr, dX, nullptr, dXn, kEpiResidual, B, H,
                         W, st)))
          return rc;
      }
      bf16* t = dX; dX = dXn; dXn = t;
    }
    // group skip: d gin = dX + dCur  (-> d gout[g - 1], or the f0-level gradient after group 0)
    FEN_CUDA(launch_pdl(add_bf16_kernel, dim3(ew_blocks(n8)), dim3(256), 0, st, dX, dCur, g ? dG(g - 1) : act(ws.dF), n8));
    FEN_CUDA(cudaGetLastError());
    ++g_launches;
    // slope and SE-matrix gradients of the group's RCABs: fixed point -> fp32
    grads_from_fixed_kernel<<<dim3(2, L.Bk), 256, 0, st>>>(g64, grads, L.p_rcab0 + g * L.p_group_stride, L.p_rcab_stride,
                                                            kConvW + 64, 64, 2 * (kConvW + 64) + 64, 2 * L.R * 64);
    FEN_CUDA(cudaGetLastError());
    ++g_launches;
    // deferred weight gradients: everything since the previous batch (job 0 = conv_after_body rides with the first)
    if (bwd_is_flush(L, g)) {
      const int prev = bwd_prev_flush(L, g);
      const int job_begin = prev < 0 ? 0 : 1 + (L.G - prev) * per;
      const int job_end = 1 + (L.G - g) * per;
      if ((rc = wgrad_batch(L, wsb, ws, grads, B, H, W, job_begin, job_end, st))) return rc;
    }
  }
  // ---- long skip + conv_first
  if (stage_begin > L.G + 1 || stage_end <= L.G + 1) return FEN_OK;
  FEN_CUDA(launch_pdl(add_bf16_kernel, dim3(ew_blocks(n8)), dim3(256), 0, st, act(ws.dF), dBody, act(ws.dF), n8));
  FEN_CUDA(cudaGetLastError());
  ++g_launches;
  // conv_first (3 -> 64): its weight gradient is columns ci < 3 of a 64 x 64 weight gradient against the LR input,
  // zero-extended from 8 channels by the tensor map
  {
    bf16* x8 = act(ws.x8);
    float* wg = reinterpret_cast<float*>(wsb + ws.wg3);
    nchw3_to_nhwc8_kernel<<<ew_blocks(size_t(B) * H * W), 256, 0, st>>>(x, x8, B, size_t(H) * W);
    FEN_CUDA(cudaGetLastError());
    ++g_launches;
    if ((rc = wgrad64(act(ws.dF), x8, wg, grads + L.p_first_b, B, H, W, 1, 0, st, parts, kC, 8))) return rc;
    FEN_CUDA(cudaMemcpy2DAsync(grads + L.p_first_w, 27 * 4, wg, 576 * 4, 27 * 4, 64, cudaMemcpyDeviceToDevice, st));
  }
  return FEN_OK;
}
