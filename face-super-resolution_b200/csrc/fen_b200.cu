// C-ABI implementation (include/fen_b200.h) of the B200-native FaceEnhanceNet forward path.
// Host orchestration + the small CUDA-core kernels; the tensor-core convolution lives in
// conv3x3_umma.cuh.  Build: see face-super-resolution_b200/build.py (nvcc, sm_100a only).
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <string>
#include <unordered_map>
#include <vector>

#include "../../include/fen_b200.h"
#include "conv3x3_umma.cuh"
#include "conv3x3_umma2.cuh"
#include "conv_last_umma.cuh"
#include "body_umma.cuh"
#include "body2_umma.cuh"
#include "fen_backward.cuh"
#include "wgrad_umma.cuh"
#include "ssim.cuh"

namespace fen {

// ===================================================================== error plumbing
static thread_local std::string g_err;
static thread_local int g_launches = 0;
static long long* g_dbg = nullptr;
static int g_time_body = 0;           // fen_profile_body(1): CUDA events around the persistent body kernel

static int fail(int code, const std::string& msg) {
  g_err = msg;
  return code;
}
#define FEN_CUDA(expr)                                                                     \
  do {                                                                                     \
    cudaError_t e_ = (expr);                                                               \
    if (e_ != cudaSuccess)                                                                 \
      return fail(FEN_ECUDA, std::string(#expr) + ": " + cudaGetErrorString(e_));          \
  } while (0)

static int check_device() {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) {
    cudaGetLastError();
    return fail(FEN_ENODEV, "no CUDA device (the FaceEnhanceNet kernels need an sm_100 GPU; there is no CPU fallback)");
  }
  int major = 0;
  if (cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev) != cudaSuccess || major != 10) {
    cudaGetLastError();
    return fail(FEN_ENODEV, "device is not sm_100 (Blackwell B200); kernels are built for sm_100a only");
  }
  return FEN_OK;
}

// Developer aid: FEN_SYNC_EACH=1 synchronises after every stage of fen_forward and names the stage
// whose kernel faulted.
static int stage_check(const char* name, cudaStream_t st) {
  static int on = -1;
  if (on < 0) { const char* e = getenv("FEN_SYNC_EACH"); on = (e && e[0] == '1') ? 1 : 0; }
  if (!on) return FEN_OK;
  cudaError_t e = cudaStreamSynchronize(st);
  if (e != cudaSuccess) return fail(FEN_ECUDA, std::string("stage '") + name + "': " + cudaGetErrorString(e));
  return FEN_OK;
}

// ===================================================================== tensor maps
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                                  CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                  CUtensorMapFloatOOBfill);
static EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    cudaDriverEntryPointQueryResult q;
    void* p = nullptr;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

// ---- cache of encoded tensor maps.  Encoding costs ~1 us on the host and a forward needs ~25 of them (a training
// step ~1 300): maps are keyed by everything that goes into them and live for the life of the process.
struct MapKey {
  const void* ptr; int kind, a, b, c, d, e;
  bool operator==(const MapKey& o) const {
    return ptr == o.ptr && kind == o.kind && a == o.a && b == o.b && c == o.c && d == o.d && e == o.e;
  }
};
struct MapKeyHash {
  size_t operator()(const MapKey& k) const {
    size_t h = std::hash<const void*>()(k.ptr);
    for (int v : {k.kind, k.a, k.b, k.c, k.d, k.e}) h = h * 1000003u ^ std::hash<int>()(v);
    return h;
  }
};
static std::mutex g_map_mutex;
static std::unordered_map<MapKey, CUtensorMap, MapKeyHash> g_map_cache;
static bool map_lookup(const MapKey& k, CUtensorMap* m) {
  std::lock_guard<std::mutex> lock(g_map_mutex);
  auto it = g_map_cache.find(k);
  if (it == g_map_cache.end()) return false;
  *m = it->second;
  return true;
}
static void map_store(const MapKey& k, const CUtensorMap& m) {
  std::lock_guard<std::mutex> lock(g_map_mutex);
  if (g_map_cache.size() > 16384) g_map_cache.clear();   // bounded: workspaces that came and went
  g_map_cache[k] = m;
}

// NHWC bf16 activation [B][H][W][64]: box = 64 ch x 66 px x 4 rows, 128B swizzle, OOB -> zeros.
// chans < 64: a narrow tensor [B][H][W][chans] (chans a multiple of 8: 16-byte pixels at least) seen through the SAME
// 64-channel box - channels chans .. 63 are out of bounds and arrive as zeros, so the 64-channel kernels run on it
// unchanged (the 3-channel ends of the network in the backward pass: 16 B per pixel from HBM instead of 128).
static int make_act_map(CUtensorMap* m, const void* ptr, int B, int H, int W, int box_rows = kBoxRows,
                        int box_px = kPitch, int chans = kC) {
  const MapKey key{ptr, chans << 8, B, H, W, box_rows, box_px};
  if (map_lookup(key, m)) return FEN_OK;
  EncodeTiledFn enc = get_encode();
  if (!enc) return fail(FEN_ECUDA, "cuTensorMapEncodeTiled not available from the driver");
  cuuint64_t dims[4] = {cuuint64_t(chans), cuuint64_t(W), cuuint64_t(H), cuuint64_t(B)};
  cuuint64_t strides[3] = {cuuint64_t(chans) * 2, cuuint64_t(W) * chans * 2, cuuint64_t(H) * W * chans * 2};
  cuuint32_t box[4] = {cuuint32_t(kC), cuuint32_t(box_px), cuuint32_t(box_rows), 1};
  cuuint32_t es[4] = {1, 1, 1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(ptr), dims, strides, box, es,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(FEN_ECUDA, "cuTensorMapEncodeTiled(activation) failed: " + std::to_string(int(r)));
  map_store(key, *m);
  return FEN_OK;
}
// packed weights [rows][64] bf16, box = 64 x N rows.
static int make_w_map(CUtensorMap* m, const void* ptr, int rows, int n) {
  const MapKey key{ptr, 1, rows, n, 0, 0, 0};
  if (map_lookup(key, m)) return FEN_OK;
  EncodeTiledFn enc = get_encode();
  if (!enc) return fail(FEN_ECUDA, "cuTensorMapEncodeTiled not available from the driver");
  cuuint64_t dims[2] = {cuuint64_t(kC), cuuint64_t(rows)};
  cuuint64_t strides[1] = {cuuint64_t(kC) * 2};
  cuuint32_t box[2] = {cuuint32_t(kC), cuuint32_t(n)};
  cuuint32_t es[2] = {1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), dims, strides, box, es,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(FEN_ECUDA, "cuTensorMapEncodeTiled(weights) failed: " + std::to_string(int(r)));
  map_store(key, *m);
  return FEN_OK;
}

// ---- per-device state: SM count and the "max dynamic shared memory" function attributes are properties of a
// device, not of the process (one process may drive several GPUs).
constexpr int kMaxDevices = 64;
enum KernelId : int { kKConv64 = 0, kKConv16, kKConv2_64, kKConv2_16, kKBody, kKBody2, kKBody2Train, kKWgMma, kKWgUmma, kKWgBatch, kKConvLast, kKernelIds };
struct DevState {
  int sms = 0;
  bool attr[kKernelIds] = {};
  cudaEvent_t body_ev[2] = {nullptr, nullptr};
};
static DevState g_dev[kMaxDevices];
static DevState& dev_state() {
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 0 || dev >= kMaxDevices) dev = 0;
  return g_dev[dev];
}
static int num_sms() {
  DevState& d = dev_state();
  if (!d.sms) {
    int dev = 0, n = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    d.sms = n > 0 ? n : 148;
  }
  return d.sms;
}
// cudaFuncAttributeMaxDynamicSharedMemorySize, once per device and kernel
template <typename K>
static cudaError_t ensure_smem_attr(KernelId id, K kernel, int bytes) {
  DevState& d = dev_state();
  if (d.attr[id]) return cudaSuccess;
  cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
  if (e == cudaSuccess) d.attr[id] = true;
  return e;
}

// Launch with programmatic stream serialisation (ptx_sm100.cuh: pdl_wait): ONLY for kernels that execute pdl_wait()
// before touching data of earlier launches.  FEN_PDL=0 launches them the plain way (A/B timing).
static bool pdl_enabled() {
  static int on = -1;
  if (on < 0) { const char* e = getenv("FEN_PDL"); on = (e && e[0] == '0') ? 0 : 1; }
  return on != 0;
}
template <typename... KArgs, typename... Args>
static cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at; cfg.numAttrs = pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

// ===================================================================== conv launcher
struct ConvArgs {
  const void* x;        // NHWC bf16 [B][H][W][64]
  const void* w;        // packed bf16 [groups*9*N][64]
  int n;                // 64 or 16
  int groups;           // gridDim.y (4 for the PixelShuffle convs)
  ConvParams p;
  int in_chans;         // 0 / 64: x is [B][H][W][64]; 8: a narrow [B][H][W][8] tensor, zero-extended by TMA (make_act_map)
};

template <int N>
static int launch_conv_n(const ConvArgs& a, cudaStream_t st) {
  FEN_CUDA(ensure_smem_attr(N == 64 ? kKConv64 : kKConv16, conv3x3_umma_kernel<N>, ConvCfg<N>::kDynBytes));
  CUtensorMap tm_in, tm_w;
  int rc = make_act_map(&tm_in, a.x, a.p.B, a.p.H, a.p.W, kBoxRows, kPitch, a.in_chans ? a.in_chans : kC);
  if (rc) return rc;
  rc = make_w_map(&tm_w, a.w, a.groups * 9 * N, N);
  if (rc) return rc;
  ConvParams p = a.p;
  p.strips = (p.W + kStripW - 1) / kStripW;
  p.tiles_per_seg = (p.H * kPitch + kTileM - 1) / kTileM;
  p.total_tiles = p.B * p.strips * p.tiles_per_seg;
  const int ctas_y = a.groups;
  int ctas_x = num_sms() / ctas_y;
  if (ctas_x < 1) ctas_x = 1;
  p.tiles_per_cta = (p.total_tiles + ctas_x - 1) / ctas_x;
  ctas_x = (p.total_tiles + p.tiles_per_cta - 1) / p.tiles_per_cta;
  dim3 grid(ctas_x, ctas_y);
  // The table-driven second generation serves conv_last (N = 16: 220 vs 240 us at batch 64) when a CTA's run fits its
  // tables; the 64-wide convolutions measured faster on the first generation (285 vs 320 us for the upsample stages).
  int conv_version = (N == 16) ? 2 : 1;
#ifdef FEN_DEV
  { const char* e = getenv("FEN_CONV_KERNEL"); if (e && e[0] == '1') conv_version = 1; else if (e && e[0] == '3') conv_version = 3; }
#endif
  const int boxes_bound = p.tiles_per_cta * kTileM / kBoxPx + 3 * ((p.tiles_per_cta + p.tiles_per_seg - 1) / p.tiles_per_seg + 1);
  if ((conv_version == 2 && N == 16 || conv_version == 3) && p.epi < kEpiGate && !p.mask_out && !p.sums64 && p.tiles_per_cta <= kC2MaxTiles && boxes_bound <= kC2MaxBoxes) {
    FEN_CUDA(ensure_smem_attr(N == 64 ? kKConv2_64 : kKConv2_16, conv3x3_umma2_kernel<N>, ConvCfg<N>::kDynBytes));
    FEN_CUDA(launch_pdl(conv3x3_umma2_kernel<N>, grid, dim3(ConvCfg<N>::kThreads), ConvCfg<N>::kDynBytes, st, tm_in, tm_w, p));
  } else {
    FEN_CUDA(launch_pdl(conv3x3_umma_kernel<N>, grid, dim3(ConvCfg<N>::kThreads), ConvCfg<N>::kDynBytes, st, tm_in, tm_w, p));
  }
  FEN_CUDA(cudaGetLastError());
  ++g_launches;
  return FEN_OK;
}

static int launch_conv(const ConvArgs& a, cudaStream_t st) {
  if (a.p.W <= 0 || a.p.H <= 0 || a.p.B <= 0) return fail(FEN_EINVAL, "conv: empty tensor");
  if (a.n == 64) return launch_conv_n<64>(a, st);
  if (a.n == 16) return launch_conv_n<16>(a, st);
  return fail(FEN_EINVAL, "conv: unsupported N");
}

// packed conv_last record: [w bf16 9*16*64][b 16f][taps-in-N weights bf16 32*64 (conv_last_umma.cuh)]
static constexpr int64_t kLastBiasOff = 9 * 16 * 64 * 2;
static constexpr int64_t kLastNOff = kLastBiasOff + 256;
static constexpr int64_t kLastRec = kLastNOff + kCLWBytes;
// conv_last + bicubic skip + clamp (conv_last_umma.cuh): `rec` = the packed conv_last record, u1 [B][H][W][64] bf16.
// FEN_LAST_KERNEL=0 (developer builds) runs the generic N = 16 convolution instead.
static int launch_conv_last(const uint8_t* rec, const void* u1, const float* lr, float* out_f32, uint8_t* out_u8, int bgr,
                            int training, int B, int H, int W, cudaStream_t st) {
#ifdef FEN_DEV
  { const char* e = getenv("FEN_LAST_KERNEL");
    if (e && e[0] == '0') {
      ConvArgs a{};
      a.x = u1; a.w = rec; a.n = 16; a.groups = 1;
      a.p.B = B; a.p.H = H; a.p.W = W; a.p.epi = kEpiLast; a.p.training = training;
      a.p.bias = reinterpret_cast<const float*>(rec + kLastBiasOff);
      a.p.lr = lr; a.p.out_f32 = out_f32; a.p.out_u8 = out_u8; a.p.bgr = bgr;
      return launch_conv(a, st);
    } }
#endif
  FEN_CUDA(ensure_smem_attr(kKConvLast, conv_last_umma_kernel, kCLDynBytes));
  CUtensorMap tm_in, tm_w;
  int rc = make_act_map(&tm_in, u1, B, H, W, kCLBoxRows, kPitch);
  if (rc) return rc;
  if ((rc = make_w_map(&tm_w, rec + kLastNOff, kCLN, kCLN))) return rc;
  ConvLastParams p{};
  p.B = B; p.H = H; p.W = W;
  p.strips = (W + kStripW - 1) / kStripW;
  p.nblk = (H + kCLRows - 1) / kCLRows;
  p.items = B * p.strips * p.nblk;
  int grid = p.items < num_sms() ? p.items : num_sms();
  p.items_per_cta = (p.items + grid - 1) / grid;
  grid = (p.items + p.items_per_cta - 1) / p.items_per_cta;
  p.training = training; p.bgr = bgr;
  p.bias = reinterpret_cast<const float*>(rec + kLastBiasOff);
  p.lr = lr; p.out_f32 = out_f32; p.out_u8 = out_u8;
  FEN_CUDA(launch_pdl(conv_last_umma_kernel, dim3(grid), dim3(kCLThreads), kCLDynBytes, st, tm_in, tm_w, p));
  FEN_CUDA(cudaGetLastError());
  ++g_launches;
  return FEN_OK;
}

// ===================================================================== small kernels
// fp32 OIHW [cout][64][3][3] -> bf16 [grp][tap][n_grp][64].  perm = 0: one group of cout_pad rows
// (row r <- cout r, zero if r >= cout).  perm = 1: PixelShuffle grouping, 4 groups of 64 rows,
// group `sub` row c <- cout 4c + sub  (out[c, 2h+i, 2w+j] = in[4c + 2i + j, h, w], blocks.py:215).
__global__ void pack_conv_kernel(const float* __restrict__ w, bf16* __restrict__ dst, int cout, int n_grp,
                                 int groups, int perm) {
  const int total = groups * 9 * n_grp * kC;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int ci = i % kC;
    const int r = (i / kC) % n_grp;
    const int tap = (i / (kC * n_grp)) % 9;
    const int grp = i / (kC * n_grp * 9);
    const int co = perm ? 4 * r + grp : r;
    float v = 0.f;
    if (co < cout) v = w[(size_t(co) * kC + ci) * 9 + tap];
    dst[i] = __float2bfloat16(v);
  }
}
// dst[g*n + r] = src[perm ? 4r + g : r] (0 when out of range)
__global__ void pack_vec_kernel(const float* __restrict__ src, float* __restrict__ dst, int count, int n_grp,
                                int groups, int perm) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= groups * n_grp) return;
  const int r = i % n_grp, g = i / n_grp;
  const int s = perm ? 4 * r + g : r;
  dst[i] = s < count ? src[s] : 0.f;
}
// conv_first weights OIHW [64][3][3][3] -> [27][64] (k = ci*9 + tap major, cout minor)
__global__ void pack_first_kernel(const float* __restrict__ w, float* __restrict__ dst) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= 27 * kC) return;
  const int co = i % kC, k = i / kC;
  dst[i] = w[co * 27 + k];
}

// conv_first (custom.py:91-94,164): 3 -> 64 channels, fp32 math on CUDA cores (K = 27 is a poor MMA
// shape and 0.03 % of the FLOPs).  x fp32 NCHW -> out bf16 NHWC.  One block = one image row.
__global__ void __launch_bounds__(256) conv_first_kernel(const float* __restrict__ x, const float* __restrict__ wk,
                                                         const float* __restrict__ bias, bf16* __restrict__ out,
                                                         int H, int W) {
  __shared__ float sw[27 * kC];
  __shared__ float sb[kC];
  for (int i = threadIdx.x; i < 27 * kC; i += blockDim.x) sw[i] = wk[i];
  if (threadIdx.x < kC) sb[threadIdx.x] = bias[threadIdx.x];
  __syncthreads();
  const int n = blockIdx.y, y = blockIdx.x;
  const int q = threadIdx.x >> 6;  // 16-channel quarter, warp-uniform
  for (int x0 = 0; x0 < W; x0 += 64) {
    const int xx = x0 + (threadIdx.x & 63);
    float in[27];
#pragma unroll
    for (int ci = 0; ci < 3; ++ci)
#pragma unroll
      for (int ky = 0; ky < 3; ++ky)
#pragma unroll
        for (int kx = 0; kx < 3; ++kx) {
          const int yy = y + ky - 1, xc = xx + kx - 1;
          in[ci * 9 + ky * 3 + kx] =
              (yy >= 0 && yy < H && xc >= 0 && xc < W) ? __ldg(x + ((size_t(n) * 3 + ci) * H + yy) * W + xc) : 0.f;
        }
    float acc[16];
#pragma unroll
    for (int c = 0; c < 16; ++c) acc[c] = sb[q * 16 + c];
#pragma unroll
    for (int k = 0; k < 27; ++k) {
      const float4* wp = reinterpret_cast<const float4*>(sw + k * kC + q * 16);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float4 w4 = wp[j];
        acc[4 * j + 0] = fmaf(in[k], w4.x, acc[4 * j + 0]);
        acc[4 * j + 1] = fmaf(in[k], w4.y, acc[4 * j + 1]);
        acc[4 * j + 2] = fmaf(in[k], w4.z, acc[4 * j + 2]);
        acc[4 * j + 3] = fmaf(in[k], w4.w, acc[4 * j + 3]);
      }
    }
    if (xx >= W) continue;   // ragged width (no barrier inside this loop)
    uint4* dst = reinterpret_cast<uint4*>(out + ((size_t(n) * H + y) * W + xx) * kC + q * 16);
    uint4 o0, o1;
    o0.x = pack_bf16(acc[0], acc[1]);   o0.y = pack_bf16(acc[2], acc[3]);
    o0.z = pack_bf16(acc[4], acc[5]);   o0.w = pack_bf16(acc[6], acc[7]);
    o1.x = pack_bf16(acc[8], acc[9]);   o1.y = pack_bf16(acc[10], acc[11]);
    o1.z = pack_bf16(acc[12], acc[13]); o1.w = pack_bf16(acc[14], acc[15]);
    dst[0] = o0;
    dst[1] = o1;
  }
}

// Squeeze-and-excitation + residual (blocks.py:86-92,153):
//   s = sigmoid(W2 relu(W0 mean_hw(o)));  x' = o * s * res_scale + x
// `sums` holds the per-image channel sums produced by the conv2 epilogue.  grid (chunks, B).
__global__ void __launch_bounds__(256) se_residual_kernel(const bf16* __restrict__ x, const bf16* __restrict__ o,
                                                          const long long* __restrict__ sums,
                                                          const float* __restrict__ fc0, const float* __restrict__ fc2,
                                                          int R, float inv_hw, float res_scale, bf16* __restrict__ xout,
                                                          float* __restrict__ se_out, int se_stride, int hw) {
  __shared__ float s_mean[kC], s_hid[kC], s_scale[kC];
  const int n = blockIdx.y, tid = threadIdx.x;
  if (tid < kC) s_mean[tid] = hs_to_float(sums[n * kC + tid]) * inv_hw;
  __syncthreads();
  if (tid < R) {
    float a = 0.f;
    for (int c = 0; c < kC; ++c) a = fmaf(fc0[tid * kC + c], s_mean[c], a);
    s_hid[tid] = fmaxf(a, 0.f);
  }
  __syncthreads();
  if (tid < kC) {
    float a = 0.f;
    for (int j = 0; j < R; ++j) a = fmaf(fc2[tid * R + j], s_hid[j], a);
    const float s = 1.f / (1.f + expf(-a));
    s_scale[tid] = s * res_scale;
    if (se_out && blockIdx.x == 0) se_out[size_t(n) * se_stride + tid] = s;
  }
  __syncthreads();
  const size_t base = size_t(n) * hw * (kC / 8);
  const int total = hw * (kC / 8);
  const uint4* xv = reinterpret_cast<const uint4*>(x) + base;
  const uint4* ov = reinterpret_cast<const uint4*>(o) + base;
  uint4* dv = reinterpret_cast<uint4*>(xout) + base;
  for (int i = blockIdx.x * blockDim.x + tid; i < total; i += gridDim.x * blockDim.x) {
    const int c0 = (i & 7) * 8;
    const uint4 a = __ldg(xv + i), b = __ldg(ov + i);
    uint4 r;
    r.x = pack_bf16(fmaf(bf16lo(b.x), s_scale[c0 + 0], bf16lo(a.x)), fmaf(bf16hi(b.x), s_scale[c0 + 1], bf16hi(a.x)));
    r.y = pack_bf16(fmaf(bf16lo(b.y), s_scale[c0 + 2], bf16lo(a.y)), fmaf(bf16hi(b.y), s_scale[c0 + 3], bf16hi(a.y)));
    r.z = pack_bf16(fmaf(bf16lo(b.z), s_scale[c0 + 4], bf16lo(a.z)), fmaf(bf16hi(b.z), s_scale[c0 + 5], bf16hi(a.z)));
    r.w = pack_bf16(fmaf(bf16lo(b.w), s_scale[c0 + 6], bf16lo(a.w)), fmaf(bf16hi(b.w), s_scale[c0 + 7], bf16hi(a.w)));
    dv[i] = r;
  }
}

// Bicubic /4 LR generator, integer and bit-exact against cv2.resize(INTER_CUBIC) on uint8
// (dataset.py:296, prepare_data.py:38): u = sum a_i a_j hr[4y+i][4x+j][c], a = [-3,19,19,-3],
// lr = clamp(round_half_even(u / 1024), 0, 255).  One thread per output pixel (all C channels).
template <int C>
__global__ void __launch_bounds__(256) lr_from_hr_kernel(const uint8_t* __restrict__ hr, uint8_t* __restrict__ lr_u8,
                                                         float* __restrict__ lr_f32, int B, int H, int W) {
  const int w = W >> 2, h = H >> 2;
  const size_t total = size_t(B) * h * w;
  for (size_t i = size_t(blockIdx.x) * blockDim.x + threadIdx.x; i < total; i += size_t(gridDim.x) * blockDim.x) {
    const int x = int(i % w);
    const int y = int((i / w) % h);
    const int n = int(i / (size_t(w) * h));
    const uint8_t* src = hr + ((size_t(n) * H + 4 * y) * W + 4 * x) * C;
    int u[C];
#pragma unroll
    for (int c = 0; c < C; ++c) u[c] = 0;
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      const int ar = (r == 0 || r == 3) ? -3 : 19;
      uint8_t px[4 * C];
      if constexpr ((4 * C) % 4 == 0) {
        const uint32_t* s32 = reinterpret_cast<const uint32_t*>(src + size_t(r) * W * C);  // 4*C*x is 4-aligned
#pragma unroll
        for (int k = 0; k < C; ++k) {
          const uint32_t v = __ldg(s32 + k);
          px[4 * k + 0] = v & 0xff; px[4 * k + 1] = (v >> 8) & 0xff;
          px[4 * k + 2] = (v >> 16) & 0xff; px[4 * k + 3] = v >> 24;
        }
      }
#pragma unroll
      for (int c = 0; c < C; ++c) {
        const int rowv = -3 * int(px[c]) + 19 * int(px[C + c]) + 19 * int(px[2 * C + c]) - 3 * int(px[3 * C + c]);
        u[c] += ar * rowv;
      }
    }
#pragma unroll
    for (int c = 0; c < C; ++c) {
      int qv = (u[c] + 511 + ((u[c] >> 10) & 1)) >> 10;
      qv = min(max(qv, 0), 255);
      if (lr_u8) lr_u8[i * C + c] = uint8_t(qv);
      if (lr_f32) lr_f32[((size_t(n) * C + c) * h + y) * w + x] = float(qv) / 255.0f;
    }
  }
}

// The RGB case the pipeline runs (C = 3, W a multiple of 16, 16-byte aligned rows): one thread makes FOUR output pixels
// from 4 rows x 48 contiguous bytes (three 128-bit loads per row) and writes 12 bytes of u8 (three 32-bit stores) and
// three float4 of the normalised NCHW tensor.  Same integer arithmetic as lr_from_hr_kernel, bit for bit; the point is
// bytes in flight per thread and whole-sector stores: the kernel is HBM bound (208 896 B per image).
__global__ void __launch_bounds__(256) lr_from_hr_rgb4_kernel(const uint8_t* __restrict__ hr, uint8_t* __restrict__ lr_u8,
                                                              float* __restrict__ lr_f32, int B, int H, int W) {
  const int w = W >> 2, h = H >> 2, wq = w >> 2;
  const size_t total = size_t(B) * h * wq;
  for (size_t i = size_t(blockIdx.x) * blockDim.x + threadIdx.x; i < total; i += size_t(gridDim.x) * blockDim.x) {
    const int xq = int(i % wq);
    const int y = int((i / wq) % h);
    const int n = int(i / (size_t(wq) * h));
    const uint8_t* src = hr + ((size_t(n) * H + 4 * y) * W + 16 * xq) * 3;
    int u[4][3];
#pragma unroll
    for (int p = 0; p < 4; ++p)
#pragma unroll
      for (int c = 0; c < 3; ++c) u[p][c] = 0;
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      const int ar = (r == 0 || r == 3) ? -3 : 19;
      const uint4* s128 = reinterpret_cast<const uint4*>(src + size_t(r) * W * 3);
      uint32_t wds[12];
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        const uint4 v = __ldcs(s128 + k);      // streamed once: evict-first, the 201 MB batch must not sweep L2
        wds[4 * k] = v.x; wds[4 * k + 1] = v.y; wds[4 * k + 2] = v.z; wds[4 * k + 3] = v.w;
      }
#pragma unroll
      for (int p = 0; p < 4; ++p) {           // output pixel p reads bytes 12 p .. 12 p + 11 = words 3 p .. 3 p + 2
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          int rowv = 0;
#pragma unroll
          for (int t = 0; t < 4; ++t) {
            const int byte = 3 * t + c;       // within the 12 bytes of this output pixel
            const int v = int((wds[3 * p + (byte >> 2)] >> (8 * (byte & 3))) & 0xffu);
            rowv += ((t == 0 || t == 3) ? -3 : 19) * v;
          }
          u[p][c] += ar * rowv;
        }
      }
    }
    int q[4][3];
#pragma unroll
    for (int p = 0; p < 4; ++p)
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        const int qv = (u[p][c] + 511 + ((u[p][c] >> 10) & 1)) >> 10;
        q[p][c] = min(max(qv, 0), 255);
      }
    if (lr_u8) {
      uint32_t o[3] = {0u, 0u, 0u};
#pragma unroll
      for (int p = 0; p < 4; ++p)
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          const int byte = 3 * p + c;
          o[byte >> 2] |= uint32_t(q[p][c]) << (8 * (byte & 3));
        }
      uint32_t* dst = reinterpret_cast<uint32_t*>(lr_u8 + ((size_t(n) * h + y) * w + 4 * xq) * 3);
      dst[0] = o[0]; dst[1] = o[1]; dst[2] = o[2];
    }
    if (lr_f32) {
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        float4 f;
        f.x = float(q[0][c]) / 255.0f; f.y = float(q[1][c]) / 255.0f;
        f.z = float(q[2][c]) / 255.0f; f.w = float(q[3][c]) / 255.0f;
        *reinterpret_cast<float4*>(lr_f32 + ((size_t(n) * 3 + c) * h + y) * w + 4 * xq) = f;
      }
    }
  }
}

// Float bicubic /4 (trainer.py:416-421, scripts/test_model.py:139-156): out = sum_i w_i (sum_j w_j in[4y+i][4x+j]),
// w = [-3, 19, 19, -3] / 32 (cubic convolution, A = -0.75, t = 0.5; exact in fp32), rows first like ATen.
// One thread per output pixel and channel; the 4 x 4 source block is four 16-byte loads.
__global__ void __launch_bounds__(256) lr_from_hr_f32_kernel(const float* __restrict__ hr, float* __restrict__ lr_f32,
                                                             uint8_t* __restrict__ lr_u8, int B, int C, int H, int W,
                                                             int bgr) {
  const int w = W >> 2, h = H >> 2;
  const size_t total = size_t(B) * C * h * w;
  const float w0 = -0.09375f, w1 = 0.59375f;
  for (size_t i = size_t(blockIdx.x) * blockDim.x + threadIdx.x; i < total; i += size_t(gridDim.x) * blockDim.x) {
    const int x = int(i % w);
    const int y = int((i / w) % h);
    const int c = int((i / (size_t(w) * h)) % C);
    const int n = int(i / (size_t(w) * h * C));
    const float* src = hr + ((size_t(n) * C + c) * H + 4 * y) * W + 4 * x;
    float r[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float4 v = __ldg(reinterpret_cast<const float4*>(src + size_t(k) * W));
      r[k] = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(v.x, w0), __fmul_rn(v.y, w1)), __fmul_rn(v.z, w1)), __fmul_rn(v.w, w0));
    }
    const float o = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(r[0], w0), __fmul_rn(r[1], w1)), __fmul_rn(r[2], w1)), __fmul_rn(r[3], w0));
    if (lr_f32) lr_f32[i] = o;
    if (lr_u8) {
      const float q = fminf(fmaxf(__fmul_rn(o, 255.0f), 0.f), 255.f);
      const int cc = bgr ? C - 1 - c : c;
      lr_u8[((size_t(n) * h + y) * w + x) * C + cc] = uint8_t(int(q));   // truncation, as numpy astype(uint8)
    }
  }
}

// to_numpy of the scripts (test_model.py:176-190): fp32 NCHW -> uint8 HWC, trunc(clip(v * 255, 0, 255)).
__global__ void __launch_bounds__(256) sr_to_u8_kernel(const float* __restrict__ sr, uint8_t* __restrict__ out, int B, int C,
                                                       int H, int W, int bgr) {
  const size_t hw = size_t(H) * W, total = size_t(B) * hw;
  for (size_t i = size_t(blockIdx.x) * blockDim.x + threadIdx.x; i < total; i += size_t(gridDim.x) * blockDim.x) {
    const size_t n = i / hw, pix = i - n * hw;
    for (int c = 0; c < C; ++c) {
      const float q = fminf(fmaxf(__fmul_rn(__ldg(sr + (n * C + c) * hw + pix), 255.0f), 0.f), 255.f);
      out[i * C + (bgr ? C - 1 - c : c)] = uint8_t(int(q));
    }
  }
}

// ===================================================================== loss / optimiser kernels (Stage-1 step)
constexpr int kRedBlocks = 1024;   // partial sums of the two-stage reductions (fixed -> deterministic)

__device__ __forceinline__ float block_sum_256(float v, float* sh) {
#pragma unroll
  for (int d = 16; d >= 1; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
  __syncthreads();
  float t = (threadIdx.x < 8) ? sh[threadIdx.x] : 0.f;
  if (threadIdx.x < 32) {
#pragma unroll
    for (int d = 4; d >= 1; d >>= 1) t += __shfl_xor_sync(0xffffffffu, t, d);
  }
  return t;   // valid in thread 0
}
// mode 0: sum |a - b| (and dsr = sign(a - b) * inv_n); mode 1: sum a^2; mode 2: sum (a - b)^2
__global__ void __launch_bounds__(256) reduce_partial_kernel(const float* __restrict__ a, const float* __restrict__ b,
                                                            float* __restrict__ dsr, int64_t n, float inv_n, int mode,
                                                            float* __restrict__ partial) {
  __shared__ float sh[8];
  float acc = 0.f;
  for (int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += int64_t(gridDim.x) * blockDim.x) {
    if (mode == 0) {
      const float d = a[i] - b[i];
      acc += fabsf(d);
      if (dsr) dsr[i] = (d > 0.f) ? inv_n : (d < 0.f ? -inv_n : 0.f);
    } else if (mode == 1) {
      acc = fmaf(a[i], a[i], acc);
    } else {
      const float d = a[i] - b[i];
      acc = fmaf(d, d, acc);
    }
  }
  const float t = block_sum_256(acc, sh);
  if (threadIdx.x == 0) partial[blockIdx.x] = t;
}
// out[0] = scale * sum(partial)  (or sqrt of it), accumulated in double by one block
__global__ void __launch_bounds__(256) reduce_final_kernel(const float* __restrict__ partial, int count, float scale,
                                                          int take_sqrt, float* __restrict__ out) {
  __shared__ double sh[256];
  double acc = 0.0;
  for (int i = threadIdx.x; i < count; i += 256) acc += double(partial[i]);
  sh[threadIdx.x] = acc;
  __syncthreads();
  for (int d = 128; d >= 1; d >>= 1) {
    if (threadIdx.x < d) sh[threadIdx.x] += sh[threadIdx.x + d];
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    if (take_sqrt == 1) out[0] = float(sqrt(sh[0]));
    else if (take_sqrt == 2) out[0] = float(10.0 * log10(1.0 / (sh[0] * double(scale))));   // PSNR; mse == 0 -> +inf
    else out[0] = float(sh[0] * double(scale));
  }
}
// clip_grad_norm_ + AdamW (torch.optim.AdamW single-tensor formulas, trainer.py:217-221,490-503)
__global__ void __launch_bounds__(256) clip_adamw_kernel(float* __restrict__ p, const float* __restrict__ g,
                                                        float* __restrict__ m, float* __restrict__ v, int64_t n,
                                                        const float* __restrict__ total_norm, float max_norm, float lr,
                                                        float beta1, float beta2, float eps, float wd, float step_size,
                                                        float bc2_sqrt) {
  float coef = 1.f;
  if (max_norm > 0.f) coef = fminf(max_norm / (__ldg(total_norm) + 1e-6f), 1.f);
  for (int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += int64_t(gridDim.x) * blockDim.x) {
    const float gi = g[i] * coef;
    float pi = p[i] * (1.f - lr * wd);
    const float mi = m[i] + (gi - m[i]) * (1.f - beta1);          // exp_avg.lerp_(grad, 1 - beta1)
    const float vi = v[i] * beta2 + (1.f - beta2) * gi * gi;      // mul_(beta2).addcmul_(grad, grad, 1 - beta2)
    const float denom = sqrtf(vi) / bc2_sqrt + eps;
    pi -= step_size * (mi / denom);
    p[i] = pi; m[i] = mi; v[i] = vi;
  }
}

// ===================================================================== parameter / blob layout
struct Layout {
  int C, G, Bk, R, n_rcab;
  // flat fp32 parameter offsets (elements)
  int64_t p_first_w, p_first_b, p_rcab0, p_rcab_stride, p_group_stride, p_gconv_w_in_group, p_after_w, p_after_b,
      p_up[2], p_last_w, p_last_b, p_total;
  // packed blob offsets (bytes)
  int64_t k_first_w, k_first_b, k_rcab0, k_rcab_stride, k_gconv0, k_gconv_stride, k_after, k_up[2], k_last, k_cvec,
      k_total;
  // bias / slope table for the persistent body kernel (copied to __constant__ c_vec): per RCAB
  // [b1 64][slope 64][b2 64], then per group conv [b 64], then conv_after_body [b 64]
  int cv_rcab0, cv_gconv0, cv_after, cv_total;
};
// per-RCAB flat params: conv1.w conv1.b prelu conv2.w conv2.b fc0 fc2
static constexpr int64_t kConvW = 64 * 64 * 9;
static constexpr int64_t kConvWBytes = 9 * 64 * 64 * 2;
static int64_t align256(int64_t v) { return (v + 255) & ~int64_t(255); }

// packed RCAB record: [w1 bf16][w2 bf16][b1 64f][slope 64f][b2 64f][fc0 R*64 f][fc2 64*R f]
struct RcabRec { int64_t w1, w2, b1, slope, b2, fc0, fc2, size; };
static RcabRec rcab_rec(int R) {
  RcabRec r;
  r.w1 = 0; r.w2 = kConvWBytes; r.b1 = 2 * kConvWBytes; r.slope = r.b1 + 256; r.b2 = r.slope + 256;
  r.fc0 = r.b2 + 256; r.fc2 = r.fc0 + int64_t(R) * 64 * 4; r.size = align256(r.fc2 + int64_t(R) * 64 * 4);
  return r;
}
// packed plain conv record: [w bf16][b 64f]
static constexpr int64_t kPlainRec = kConvWBytes + 256;
// packed upsample record: [w bf16 4 groups][b 256f permuted][slope 64f]
static constexpr int64_t kUpRec = 4 * kConvWBytes + 1024 + 256;


static int make_layout(const fen_config* cfg, Layout* L) {
  if (!cfg) return fail(FEN_EINVAL, "null config");
  if (cfg->num_channels != 64) return fail(FEN_EINVAL, "unsupported config: num_channels must be 64 (no fallback path)");
  if (cfg->scale_factor != 4) return fail(FEN_EINVAL, "unsupported config: scale_factor must be 4");
  if (cfg->num_groups < 1 || cfg->blocks_per_group < 1) return fail(FEN_EINVAL, "num_groups and blocks_per_group must be >= 1");
  if (cfg->reduction_ratio < 1) return fail(FEN_EINVAL, "reduction_ratio must be >= 1");
  L->C = 64; L->G = cfg->num_groups; L->Bk = cfg->blocks_per_group;
  L->R = 64 / cfg->reduction_ratio; if (L->R < 8) L->R = 8;
  if (L->R > 64) return fail(FEN_EINVAL, "SE hidden width > 64 unsupported");
  L->n_rcab = L->G * L->Bk;
  const int64_t rcab_p = 2 * (kConvW + 64) + 64 + 2 * int64_t(L->R) * 64;
  int64_t o = 0;
  L->p_first_w = o; o += 64 * 27; L->p_first_b = o; o += 64;
  L->p_rcab0 = o; L->p_rcab_stride = rcab_p;
  L->p_gconv_w_in_group = L->Bk * rcab_p;
  L->p_group_stride = L->Bk * rcab_p + kConvW + 64;
  o += L->G * L->p_group_stride;
  L->p_after_w = o; o += kConvW; L->p_after_b = o; o += 64;
  for (int s = 0; s < 2; ++s) { L->p_up[s] = o; o += 4 * kConvW + 256 + 64; }
  L->p_last_w = o; o += 3 * 64 * 9; L->p_last_b = o; o += 3;
  L->p_total = o;
  const RcabRec rr = rcab_rec(L->R);
  int64_t k = 0;
  L->k_first_w = k; k += align256(27 * 64 * 4); L->k_first_b = k; k += 256;
  L->k_rcab0 = k; L->k_rcab_stride = rr.size; k += int64_t(L->n_rcab) * rr.size;
  L->k_gconv0 = k; L->k_gconv_stride = kPlainRec; k += int64_t(L->G) * kPlainRec;
  L->k_after = k; k += kPlainRec;
  for (int s = 0; s < 2; ++s) { L->k_up[s] = k; k += kUpRec; }
  L->k_last = k; k += kLastRec;
  L->cv_rcab0 = 0; L->cv_gconv0 = L->n_rcab * 192; L->cv_after = L->cv_gconv0 + L->G * 64;
  L->cv_total = L->cv_after + 64;
  L->k_cvec = align256(k); k = L->k_cvec + int64_t(L->cv_total) * 4 + 512;   // (the body kernel stages 512 B from any cv_bias)
  L->k_total = align256(k);
  return FEN_OK;
}

static int pack_conv(const float* w, int cout, int n_grp, int groups, int perm, void* dst, cudaStream_t st) {
  const int total = groups * 9 * n_grp * kC;
  pack_conv_kernel<<<(total + 255) / 256, 256, 0, st>>>(w, reinterpret_cast<bf16*>(dst), cout, n_grp, groups, perm);
  FEN_CUDA(cudaGetLastError());
  return FEN_OK;
}
static int pack_vec(const float* src, int count, int n_grp, int groups, int perm, void* dst, cudaStream_t st) {
  pack_vec_kernel<<<(groups * n_grp + 255) / 256, 256, 0, st>>>(src, reinterpret_cast<float*>(dst), count, n_grp, groups, perm);
  FEN_CUDA(cudaGetLastError());
  return FEN_OK;
}

// Batched packing (one launch for all RCABs, one for all plain convolutions): a training step repacks every weight
// after the optimiser update, and one launch per tensor (~750 of them for the 6 x 10 model) cost more than the
// forward pass itself.  blockIdx.y selects the record; element order as pack_conv_kernel (perm = 0).
__device__ __forceinline__ void pack_conv64_dev(const float* __restrict__ w, bf16* __restrict__ dst, bool transposed) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < 9 * kC * kC; i += gridDim.x * blockDim.x) {
    const int c = i % kC, r = (i / kC) % kC, t = i / (kC * kC);
    // forward: dst[t][co = r][ci = c] = W[r][c][t];   transposed (dgrad): dst[t][ci = r][co = c] = W[c][r][8 - t]
    const float v = transposed ? w[(size_t(c) * kC + r) * 9 + (8 - t)] : w[(size_t(r) * kC + c) * 9 + t];
    dst[i] = __float2bfloat16(v);
  }
}
__global__ void pack_rcab_all_kernel(const float* __restrict__ params, uint8_t* __restrict__ k, Layout L, RcabRec rr) {
  const int r = blockIdx.y, g = r / L.Bk, b = r % L.Bk;
  const float* pr = params + L.p_rcab0 + g * L.p_group_stride + b * L.p_rcab_stride;
  uint8_t* kr = k + L.k_rcab0 + int64_t(r) * L.k_rcab_stride;
  const float* c1w = pr; const float* c1b = c1w + kConvW; const float* sl = c1b + 64;
  const float* c2w = sl + 64; const float* c2b = c2w + kConvW; const float* fc0 = c2b + 64;
  const float* fc2 = fc0 + L.R * 64;
  pack_conv64_dev(c1w, reinterpret_cast<bf16*>(kr + rr.w1), false);
  pack_conv64_dev(c2w, reinterpret_cast<bf16*>(kr + rr.w2), false);
  if (blockIdx.x == 0) {
    float* cv = reinterpret_cast<float*>(k + L.k_cvec) + L.cv_rcab0 + r * 192;
    for (int i = threadIdx.x; i < 64; i += blockDim.x) {
      reinterpret_cast<float*>(kr + rr.b1)[i] = c1b[i]; cv[i] = c1b[i];
      reinterpret_cast<float*>(kr + rr.slope)[i] = sl[i]; cv[64 + i] = sl[i];
      reinterpret_cast<float*>(kr + rr.b2)[i] = c2b[i]; cv[128 + i] = c2b[i];
    }
    for (int i = threadIdx.x; i < L.R * 64; i += blockDim.x) {
      reinterpret_cast<float*>(kr + rr.fc0)[i] = fc0[i];
      reinterpret_cast<float*>(kr + rr.fc2)[i] = fc2[i];
    }
  }
}
// records 0 .. G-1: the group convolutions, record G: conv_after_body
__global__ void pack_plain_all_kernel(const float* __restrict__ params, uint8_t* __restrict__ k, Layout L) {
  const int g = blockIdx.y;
  const float* w = g < L.G ? params + L.p_rcab0 + g * L.p_group_stride + L.p_gconv_w_in_group : params + L.p_after_w;
  uint8_t* kg = g < L.G ? k + L.k_gconv0 + g * L.k_gconv_stride : k + L.k_after;
  pack_conv64_dev(w, reinterpret_cast<bf16*>(kg), false);
  if (blockIdx.x == 0) {
    float* cv = reinterpret_cast<float*>(k + L.k_cvec) + (g < L.G ? L.cv_gconv0 + g * 64 : L.cv_after);
    for (int i = threadIdx.x; i < 64; i += blockDim.x) {
      reinterpret_cast<float*>(kg + kConvWBytes)[i] = w[kConvW + i]; cv[i] = w[kConvW + i];
    }
  }
}

// ===================================================================== workspace
struct Workspace {
  int64_t f0, x[2], h, o, grp0, grp_stride, u0, u1, sums, hsum, flags, total;
};
static void make_workspace(const Layout& L, int B, int H, int W, Workspace* ws) {
  const int64_t act = align256(int64_t(B) * H * W * 64 * 2);
  int64_t o = 0;
  ws->f0 = o; o += act;
  ws->x[0] = o; o += act; ws->x[1] = o; o += act;
  ws->h = o; o += act; ws->o = o; o += act;
  ws->grp0 = o; ws->grp_stride = act; o += act * L.G;
  ws->u0 = o; o += 4 * act;
  ws->u1 = o; o += 16 * act;
  ws->sums = o; o += align256(int64_t(L.n_rcab) * B * 64 * 8);   // per-layer path: fixed-point SE pool sums
  ws->hsum = o; o += align256(int64_t(L.n_rcab) * B * 9 * 64 * 8);   // body kernel: 9 channel sums of h per RCAB, image (int64 fixed point)
  ws->flags = o; o += 4096;   // one int per CTA of the persistent body kernel
  ws->total = o;
}


#ifdef FEN_DEV
// ===================================================================== persistent body kernel launcher
static bool body_kernel_usable(const Layout& L, int B, int H, int W) {
  static int disabled = -1;
  if (disabled < 0) {
    const char* e = getenv("FEN_DISABLE_BODY_KERNEL");
    disabled = (e && e[0] == '1') ? 1 : 0;
  }
  if (disabled || W != kStripW || L.cv_total > kConstVecFloats || L.G > kBodyMaxBufs - 5) return false;
  // a CTA must not touch more than kBodyMaxUnits images per layer
  const int tps = (H * kPitch + kTileM - 1) / kTileM;
  const int tpc = (B * tps + num_sms() - 1) / num_sms();
  return (tpc + tps - 2) / tps + 1 <= kBodyMaxUnits;
}

static int launch_body(const fen_config* cfg, const Layout& L, const Workspace& ws, uint8_t* wsb, const uint8_t* k,
                       int B, int H, int W, float* se_out, cudaStream_t st) {
  FEN_CUDA(ensure_smem_attr(kKBody, body_umma_kernel, kBodyDynBytes));
  const RcabRec rr = rcab_rec(L.R);
  BodyMaps maps;
  BodyParams p{};
  p.B = B; p.H = H; p.W = W; p.G = L.G; p.Bk = L.Bk; p.R = L.R;
  p.n_layers = L.G * (2 * L.Bk + 1) + 1;
  p.tiles_per_seg = (H * kPitch + kTileM - 1) / kTileM;
  p.total_tiles = B * p.tiles_per_seg;
  p.tiles_per_cta = (p.total_tiles + num_sms() - 1) / num_sms();
  const int ctas = (p.total_tiles + p.tiles_per_cta - 1) / p.tiles_per_cta;
  p.res_scale = cfg->res_scale; p.inv_hw = 1.f / float(H * W);
  const int nbuf = 5 + L.G;
  int64_t offs[kBodyMaxBufs];
  offs[kBufF0] = ws.f0; offs[kBufX0] = ws.x[0]; offs[kBufX1] = ws.x[1]; offs[kBufH] = ws.h; offs[kBufO] = ws.o;
  for (int g = 0; g < L.G; ++g) offs[kBufG0 + g] = ws.grp0 + g * ws.grp_stride;
  for (int i = 0; i < nbuf; ++i) {
    p.buf[i] = reinterpret_cast<bf16*>(wsb + offs[i]);
    int rc = make_act_map(&maps.act[i], p.buf[i], B, H, W, kBBoxRows);
    if (rc) return rc;
  }
  for (int i = nbuf; i < kBodyMaxBufs; ++i) maps.act[i] = maps.act[0];
  int rc = make_w_map(&maps.w, k, int(L.k_total / 128), kC);
  if (rc) return rc;
  p.packed = k;
  p.k_rcab0 = L.k_rcab0; p.k_rcab_stride = L.k_rcab_stride; p.k_rcab_w2 = rr.w2; p.k_rcab_fc0 = rr.fc0; p.k_rcab_fc2 = rr.fc2;
  p.k_gconv0 = L.k_gconv0; p.k_gconv_stride = L.k_gconv_stride; p.k_after = L.k_after;
  p.cv_rcab0 = L.cv_rcab0; p.cv_gconv0 = L.cv_gconv0; p.cv_after = L.cv_after;
  p.hsum = reinterpret_cast<float*>(wsb + ws.hsum);
  p.se_out = se_out;
  p.flags = reinterpret_cast<int*>(wsb + ws.flags);
  p.dbg = g_dbg;
  FEN_CUDA(cudaMemsetAsync(p.flags, 0, 4096, st));
  FEN_CUDA(cudaMemsetAsync(p.hsum, 0, size_t(L.n_rcab) * B * 9 * 64 * 4, st));
  FEN_CUDA(cudaMemcpyToSymbolAsync(c_vec, k + L.k_cvec, size_t(L.cv_total) * 4, 0, cudaMemcpyDeviceToDevice, st));
  void* args[] = {&maps, &p};
  if (g_time_body) {
    if (!dev_state().body_ev[0]) { FEN_CUDA(cudaEventCreate(&dev_state().body_ev[0])); FEN_CUDA(cudaEventCreate(&dev_state().body_ev[1])); }
    FEN_CUDA(cudaEventRecord(dev_state().body_ev[0], st));
  }
  FEN_CUDA(cudaLaunchCooperativeKernel(reinterpret_cast<void*>(body_umma_kernel), dim3(ctas), dim3(kBodyThreads), args,
                                       kBodyDynBytes, st));
  if (g_time_body) FEN_CUDA(cudaEventRecord(dev_state().body_ev[1], st));
  ++g_launches;
  return FEN_OK;
}

#endif  // FEN_DEV

// ===================================================================== second-generation body kernel launcher
static int env_int(const char* name, int dflt) {
  const char* e = getenv(name);
  return (e && e[0]) ? atoi(e) : dflt;
}
// 2 interleaved image sets when the batch is even (FEN_BODY_NSET=1 forces one set)
static int body2_nset(int B) {
  static int forced = -1;
  if (forced < 0) forced = env_int("FEN_BODY_NSET", 0);
  if (forced == 1 || (B & 1)) return 1;
  return 2;
}
static bool body2_usable(const Layout& L, int B, int H, int W) {
  int version = 2;
#ifdef FEN_DEV
  version = env_int("FEN_BODY_KERNEL", 2);
#endif
  if (version != 2 || W != kStripW) return false;
  const int tps = (H * kPitch + kTileM - 1) / kTileM;
  if (tps > 255) return false;
  const int set_tiles = (B / body2_nset(B)) * tps;
  const int tpc = (set_tiles + num_sms() - 1) / num_sms();   // longest run of a CTA in one set
  if (tpc > kB2MaxTiles - 8) return false;
  return (tpc + tps - 2) / tps + 1 <= kBodyMaxUnits;
}

// [nbuf][B][H][W][64] bf16 activation buffers at a constant stride: one 5-D map, box 64 ch x 66 px x 2 rows.
static int make_act5_map(CUtensorMap* m, const void* base, int64_t buf_stride_bytes, int nbuf, int B, int H, int W,
                         int box_px, int box_rows) {
  const MapKey key{base, 2 + 16 * box_rows + 256 * box_px, B, H, W, nbuf, int(buf_stride_bytes >> 8)};
  if (map_lookup(key, m)) return FEN_OK;
  EncodeTiledFn enc = get_encode();
  if (!enc) return fail(FEN_ECUDA, "cuTensorMapEncodeTiled not available from the driver");
  cuuint64_t dims[5] = {cuuint64_t(kC), cuuint64_t(W), cuuint64_t(H), cuuint64_t(B), cuuint64_t(nbuf)};
  cuuint64_t strides[4] = {cuuint64_t(kC) * 2, cuuint64_t(W) * kC * 2, cuuint64_t(H) * W * kC * 2,
                           cuuint64_t(buf_stride_bytes)};
  cuuint32_t box[5] = {cuuint32_t(kC), cuuint32_t(box_px), cuuint32_t(box_rows), 1, 1};
  cuuint32_t es[5] = {1, 1, 1, 1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, const_cast<void*>(base), dims, strides, box, es,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(FEN_ECUDA, "cuTensorMapEncodeTiled(activation buffers) failed: " + std::to_string(int(r)));
  map_store(key, *m);
  return FEN_OK;
}

// Where the body kernel finds its buffers inside a workspace (inference: Workspace, training: StepWs).
struct Body2Bufs {
  uint8_t* act_base; int64_t act_stride; int nbuf;       // activation buffers (indices: body2_layer)
  long long* hsum64; int* flags;
  uint32_t* mask0; int64_t mask_stride_bytes; long long* pool_sums;   // training only
};

template <bool kTrain>
static int launch_body2(const fen_config* cfg, const Layout& L, const Body2Bufs& bufs, const uint8_t* k,
                        int B, int H, int W, float* se_out, cudaStream_t st) {
  FEN_CUDA(ensure_smem_attr(kTrain ? kKBody2Train : kKBody2, body2_umma_kernel<kTrain>, kB2DynBytes));
  const RcabRec rr = rcab_rec(L.R);
  Body2Maps maps;
  Body2Params p{};
  p.nset = body2_nset(B);
  p.set_B = B / p.nset;
  p.B = B; p.H = H; p.W = W; p.G = L.G; p.Bk = L.Bk; p.R = L.R;
  p.n_layers = L.G * (2 * L.Bk + 1) + 1;
  p.tiles_per_seg = (H * kPitch + kTileM - 1) / kTileM;
  p.total_tiles = p.set_B * p.tiles_per_seg;
  p.tiles_per_cta = (p.total_tiles + num_sms() - 1) / num_sms();   // (upper bound; the kernel deals q or q + 1 tiles)
  const int ctas = p.total_tiles < num_sms() ? p.total_tiles : num_sms();
  p.res_scale = cfg->res_scale; p.inv_hw = 1.f / float(H * W);
  p.act_base = reinterpret_cast<bf16*>(bufs.act_base);
  p.act_elems = bufs.act_stride / 2;
  int rc = make_act5_map(&maps.act, bufs.act_base, bufs.act_stride, bufs.nbuf, B, H, W, kPitch, kBBoxRows);
  if (rc) return rc;
  if ((rc = make_w_map(&maps.w, k, int(L.k_total / 128), kC))) return rc;
  p.packed = k;
  p.k_rcab0 = L.k_rcab0; p.k_rcab_stride = L.k_rcab_stride; p.k_rcab_w2 = rr.w2; p.k_rcab_fc0 = rr.fc0; p.k_rcab_fc2 = rr.fc2;
  p.k_gconv0 = L.k_gconv0; p.k_gconv_stride = L.k_gconv_stride; p.k_after = L.k_after;
  p.cv_rcab0 = L.cv_rcab0; p.cv_gconv0 = L.cv_gconv0; p.cv_after = L.cv_after;
  p.hsum64 = bufs.hsum64;
  p.se_out = se_out;
  p.flags = bufs.flags;
  p.mask0 = bufs.mask0; p.mask_stride = bufs.mask_stride_bytes / 4; p.pool_sums = bufs.pool_sums;
  p.dbg = g_dbg;
  FEN_CUDA(cudaMemsetAsync(p.flags, 0, 4096, st));
  FEN_CUDA(cudaMemsetAsync(p.hsum64, 0, size_t(L.n_rcab) * B * 9 * 64 * 8, st));
  p.cvec = reinterpret_cast<const float*>(k + L.k_cvec);
  void* args[] = {&maps, &p};
  if (g_time_body) {
    if (!dev_state().body_ev[0]) { FEN_CUDA(cudaEventCreate(&dev_state().body_ev[0])); FEN_CUDA(cudaEventCreate(&dev_state().body_ev[1])); }
    FEN_CUDA(cudaEventRecord(dev_state().body_ev[0], st));
  }
  FEN_CUDA(cudaLaunchCooperativeKernel(reinterpret_cast<void*>(body2_umma_kernel<kTrain>), dim3(ctas), dim3(kB2Threads),
                                       args, kB2DynBytes, st));
  if (g_time_body) FEN_CUDA(cudaEventRecord(dev_state().body_ev[1], st));
  ++g_launches;
  return FEN_OK;
}

#include "fen_step_host.cuh"

}  // namespace fen

using namespace fen;

// ===================================================================== C ABI
extern "C" {

int fen_abi_version(void) { return FEN_ABI_VERSION; }
const char* fen_last_error(void) { return g_err.c_str(); }
int fen_last_launch_count(void) { return g_launches; }
int fen_profile_body(int enable) { g_time_body = enable ? 1 : 0; return FEN_OK; }
float fen_last_body_ms(void) {
  if (!dev_state().body_ev[0] || !dev_state().body_ev[1]) return -1.f;
  if (cudaEventSynchronize(dev_state().body_ev[1]) != cudaSuccess) { cudaGetLastError(); return -1.f; }
  float ms = -1.f;
  if (cudaEventElapsedTime(&ms, dev_state().body_ev[0], dev_state().body_ev[1]) != cudaSuccess) { cudaGetLastError(); return -1.f; }
  return ms;
}
// developer hook (not in the public header): device buffer [ctas][8] of int64 cycle counters, or null
void fen_debug_set_counters(void* dev_buf) { g_dbg = static_cast<long long*>(dev_buf); }

int64_t fen_param_count(const fen_config* cfg) {
  Layout L;
  if (make_layout(cfg, &L)) return FEN_EINVAL;
  return L.p_total;
}
int64_t fen_packed_bytes(const fen_config* cfg) {
  Layout L;
  if (make_layout(cfg, &L)) return FEN_EINVAL;
  return L.k_total;
}

int fen_pack_conv3x3(const float* w_oihw, int cout, int cout_pad, void* w_packed, void* stream) {
  int rc = check_device();
  if (rc) return rc;
  if (!w_oihw || !w_packed || cout < 1 || cout_pad < cout) return fail(FEN_EINVAL, "fen_pack_conv3x3: bad arguments");
  return pack_conv(w_oihw, cout, cout_pad, 1, 0, w_packed, static_cast<cudaStream_t>(stream));
}

int fen_pack_weights(const fen_config* cfg, const float* params, void* packed, void* stream) {
  int rc = check_device();
  if (rc) return rc;
  Layout L;
  if ((rc = make_layout(cfg, &L))) return rc;
  if (!params || !packed) return fail(FEN_EINVAL, "fen_pack_weights: null pointer");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  uint8_t* k = static_cast<uint8_t*>(packed);
  const RcabRec rr = rcab_rec(L.R);
  pack_first_kernel<<<(27 * 64 + 255) / 256, 256, 0, st>>>(params + L.p_first_w, reinterpret_cast<float*>(k + L.k_first_w));
  FEN_CUDA(cudaGetLastError());
  if ((rc = pack_vec(params + L.p_first_b, 64, 64, 1, 0, k + L.k_first_b, st))) return rc;
  pack_rcab_all_kernel<<<dim3(36, L.n_rcab), 256, 0, st>>>(params, k, L, rr);
  FEN_CUDA(cudaGetLastError());
  pack_plain_all_kernel<<<dim3(36, L.G + 1), 256, 0, st>>>(params, k, L);
  FEN_CUDA(cudaGetLastError());
  for (int s = 0; s < 2; ++s) {
    const float* pu = params + L.p_up[s];
    uint8_t* ku = k + L.k_up[s];
    if ((rc = pack_conv(pu, 256, 64, 4, 1, ku, st))) return rc;
    if ((rc = pack_vec(pu + 4 * kConvW, 256, 64, 4, 1, ku + 4 * kConvWBytes, st))) return rc;
    if ((rc = pack_vec(pu + 4 * kConvW + 256, 64, 64, 1, 0, ku + 4 * kConvWBytes + 1024, st))) return rc;
  }
  if ((rc = pack_conv(params + L.p_last_w, 3, 16, 1, 0, k + L.k_last, st))) return rc;
  if ((rc = pack_vec(params + L.p_last_b, 3, 16, 1, 0, k + L.k_last + kLastBiasOff, st))) return rc;
  pack_last_n_kernel<<<(kCLN * kC + 255) / 256, 256, 0, st>>>(params + L.p_last_w, reinterpret_cast<bf16*>(k + L.k_last + kLastNOff));
  FEN_CUDA(cudaGetLastError());
  return FEN_OK;
}

int64_t fen_forward_workspace_bytes(const fen_config* cfg, int B, int H, int W) {
  Layout L;
  if (make_layout(cfg, &L)) return FEN_EINVAL;
  if (B < 1 || H < 1 || W < 1) { fail(FEN_EINVAL, "bad batch / size"); return FEN_EINVAL; }
  Workspace ws;
  make_workspace(L, B, H, W, &ws);
  return ws.total;
}

int fen_conv3x3_c64(const void* x, const void* w_packed, const float* bias, const float* slope,
                    const void* residual, float* sums, void* out, int B, int H, int W, int epilogue,
                    void* stream) {
  int rc = check_device();
  if (rc) return rc;
  g_launches = 0;
  if (!x || !w_packed || !out) return fail(FEN_EINVAL, "fen_conv3x3_c64: null pointer");
  if (epilogue == kEpiShuffle || epilogue == kEpiLast || epilogue < 0 || epilogue > kEpiBias)
    return fail(FEN_EINVAL, "fen_conv3x3_c64: unsupported epilogue");
  if (epilogue == kEpiResidual && !residual) return fail(FEN_EINVAL, "fen_conv3x3_c64: residual required");
  if (epilogue == kEpiSum && !sums) return fail(FEN_EINVAL, "fen_conv3x3_c64: sums required");
  ConvArgs a{};
  a.x = x; a.w = w_packed; a.n = 64; a.groups = 1;
  a.p.B = B; a.p.H = H; a.p.W = W; a.p.epi = epilogue; a.p.bias = bias; a.p.slope = slope;
  a.p.residual = static_cast<const bf16*>(residual); a.p.out = static_cast<bf16*>(out); a.p.sums = sums;
  a.p.dbg = g_dbg;
  return launch_conv(a, static_cast<cudaStream_t>(stream));
}

static int forward_impl(const fen_config* cfg, const void* packed, const float* x, float* out, uint8_t* out_u8, int bgr,
                        int B, int H, int W, int training, void* workspace, int64_t workspace_bytes, float* se_out,
                        void* stream) {
  int rc = check_device();
  if (rc) return rc;
  g_launches = 0;
  Layout L;
  if ((rc = make_layout(cfg, &L))) return rc;
  if (!packed || !x || (!out && !out_u8) || !workspace) return fail(FEN_EINVAL, "fen_forward: null pointer");
  if (B < 1 || H < 1 || W < 1) return fail(FEN_EINVAL, "fen_forward: empty input");
  if (H > 16384 || W > 16384) return fail(FEN_EINVAL, "fen_forward: H and W must be at most 16384");
  Workspace ws;
  make_workspace(L, B, H, W, &ws);
  if (workspace_bytes < ws.total) return fail(FEN_ENOMEM, "fen_forward: workspace too small");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  uint8_t* wsb = static_cast<uint8_t*>(workspace);
  const uint8_t* k = static_cast<const uint8_t*>(packed);
  const RcabRec rr = rcab_rec(L.R);
  auto act = [&](int64_t off) { return reinterpret_cast<bf16*>(wsb + off); };
  long long* sums = reinterpret_cast<long long*>(wsb + ws.sums);

  conv_first_kernel<<<dim3(H, B), 256, 0, st>>>(x, reinterpret_cast<const float*>(k + L.k_first_w),
                                                reinterpret_cast<const float*>(k + L.k_first_b), act(ws.f0), H, W);
  FEN_CUDA(cudaGetLastError());
  ++g_launches;

  auto conv = [&](const bf16* in, const uint8_t* w, const float* bias, const float* slope, const bf16* res,
                  long long* sm, bf16* o, int epi, int h, int w_) -> int {
    ConvArgs a{};
    a.x = in; a.w = w; a.n = 64; a.groups = (epi == kEpiShuffle) ? 4 : 1;
    a.p.B = B; a.p.H = h; a.p.W = w_; a.p.epi = epi; a.p.bias = bias; a.p.slope = slope; a.p.residual = res;
    a.p.out = o; a.p.sums64 = sm;
    return launch_conv(a, st);
  };

  if ((rc = stage_check("conv_first", st))) return rc;
  if (body2_usable(L, B, H, W)) {
    Body2Bufs bufs{};
    bufs.act_base = wsb + ws.f0; bufs.act_stride = ws.grp_stride; bufs.nbuf = 5 + L.G;   // f0 x0 x1 h o g0.. (BodyBuf order)
    bufs.hsum64 = reinterpret_cast<long long*>(wsb + ws.hsum); bufs.flags = reinterpret_cast<int*>(wsb + ws.flags);
    if ((rc = launch_body2<false>(cfg, L, bufs, k, B, H, W, se_out, st))) return rc;
    if ((rc = stage_check("body kernel", st))) return rc;
#ifdef FEN_DEV
  } else if (body_kernel_usable(L, B, H, W)) {
    if ((rc = launch_body(cfg, L, ws, wsb, k, B, H, W, se_out, st))) return rc;
    if ((rc = stage_check("body kernel", st))) return rc;
#endif
  } else {
    FEN_CUDA(cudaMemsetAsync(sums, 0, size_t(L.n_rcab) * B * 64 * 8, st));
    const int hw = H * W;
    const int se_chunks = 32;
    const bf16* cur = act(ws.f0);
    for (int g = 0; g < L.G; ++g) {
      const bf16* gin = cur;
      for (int b = 0; b < L.Bk; ++b) {
        const int r = g * L.Bk + b;
        const uint8_t* kr = k + L.k_rcab0 + int64_t(r) * L.k_rcab_stride;
        long long* sm = sums + size_t(r) * B * 64;
        if ((rc = conv(cur, kr + rr.w1, reinterpret_cast<const float*>(kr + rr.b1),
                       reinterpret_cast<const float*>(kr + rr.slope), nullptr, nullptr, act(ws.h), kEpiPrelu, H, W)))
          return rc;
        if ((rc = conv(act(ws.h), kr + rr.w2, reinterpret_cast<const float*>(kr + rr.b2), nullptr, nullptr, sm,
                       act(ws.o), kEpiSum, H, W)))
          return rc;
        bf16* nxt = act(ws.x[b & 1]);
        se_residual_kernel<<<dim3(se_chunks, B), 256, 0, st>>>(
            cur, act(ws.o), sm, reinterpret_cast<const float*>(kr + rr.fc0), reinterpret_cast<const float*>(kr + rr.fc2),
            L.R, 1.f / float(hw), cfg->res_scale, nxt, se_out ? se_out + size_t(r) * 64 : nullptr, L.n_rcab * 64, hw);
        FEN_CUDA(cudaGetLastError());
        ++g_launches;
        cur = nxt;
      }
      const uint8_t* kg = k + L.k_gconv0 + g * L.k_gconv_stride;
      bf16* gout = act(ws.grp0 + g * ws.grp_stride);
      if ((rc = conv(cur, kg, reinterpret_cast<const float*>(kg + kConvWBytes), nullptr, gin, nullptr, gout,
                     kEpiResidual, H, W)))
        return rc;
      cur = gout;
    }
    // conv_after_body + long skip -> x[0]
    if ((rc = conv(cur, k + L.k_after, reinterpret_cast<const float*>(k + L.k_after + kConvWBytes), nullptr,
                   act(ws.f0), nullptr, act(ws.x[0]), kEpiResidual, H, W)))
      return rc;
  }
  // upsample stages
  {
    const uint8_t* ku = k + L.k_up[0];
    if ((rc = conv(act(ws.x[0]), ku, reinterpret_cast<const float*>(ku + 4 * kConvWBytes),
                   reinterpret_cast<const float*>(ku + 4 * kConvWBytes + 1024), nullptr, nullptr, act(ws.u0),
                   kEpiShuffle, H, W)))
      return rc;
    ku = k + L.k_up[1];
    if ((rc = conv(act(ws.u0), ku, reinterpret_cast<const float*>(ku + 4 * kConvWBytes),
                   reinterpret_cast<const float*>(ku + 4 * kConvWBytes + 1024), nullptr, nullptr, act(ws.u1),
                   kEpiShuffle, 2 * H, 2 * W)))
      return rc;
  }
  if ((rc = stage_check("upsample convs", st))) return rc;
  // conv_last + bicubic skip + clamp
  if ((rc = launch_conv_last(k + L.k_last, act(ws.u1), x, out, out_u8, bgr, training, B, 4 * H, 4 * W, st))) return rc;
  if ((rc = stage_check("conv_last", st))) return rc;
  return FEN_OK;
}

int fen_forward(const fen_config* cfg, const void* packed, const float* x, float* out, int B, int H,
                int W, int training, void* workspace, int64_t workspace_bytes, float* se_out,
                void* stream) {
  if (!out) return fail(FEN_EINVAL, "fen_forward: null pointer");
  return forward_impl(cfg, packed, x, out, nullptr, 0, B, H, W, training, workspace, workspace_bytes, se_out, stream);
}

int fen_forward_u8(const fen_config* cfg, const void* packed, const float* x, uint8_t* out_u8, int bgr, int B, int H,
                   int W, void* workspace, int64_t workspace_bytes, void* stream) {
  if (!out_u8) return fail(FEN_EINVAL, "fen_forward_u8: null pointer");
  return forward_impl(cfg, packed, x, nullptr, out_u8, bgr ? 1 : 0, B, H, W, 0, workspace, workspace_bytes, nullptr,
                      stream);
}

int64_t fen_forward_tap(const fen_config* cfg, const void* workspace, int B, int H, int W, int which,
                        int index, const void** ptr) {
  Layout L;
  if (make_layout(cfg, &L)) return FEN_EINVAL;
  if (!workspace || !ptr) return fail(FEN_EINVAL, "fen_forward_tap: null pointer");
  Workspace ws;
  make_workspace(L, B, H, W, &ws);
  const uint8_t* b = static_cast<const uint8_t*>(workspace);
  const int64_t act = int64_t(B) * H * W * 64 * 2;
  switch (which) {
    case 0: *ptr = b + ws.f0; return act;
    case 1: *ptr = b + ws.x[0]; return act;
    case 2: *ptr = b + ws.u0; return 4 * act;
    case 3: *ptr = b + ws.u1; return 16 * act;
    case 4:
      if (index < 0 || index >= L.G) return fail(FEN_EINVAL, "fen_forward_tap: bad group index");
      *ptr = b + ws.grp0 + index * ws.grp_stride;
      return act;
    default: return fail(FEN_EINVAL, "fen_forward_tap: unknown tap");
  }
}

int64_t fen_train_workspace_bytes(int64_t n) {
  (void)n;
  return int64_t(kRedBlocks) * 4;
}

static int reduce_launch(const float* a, const float* b, float* dsr, int64_t n, float inv_n, int mode, float scale,
                         int take_sqrt, float* out, void* workspace, int64_t ws_bytes, cudaStream_t st) {
  if (ws_bytes < int64_t(kRedBlocks) * 4) return fail(FEN_ENOMEM, "training workspace too small");
  float* partial = static_cast<float*>(workspace);
  int64_t blocks = (n + 255) / 256;
  if (blocks > kRedBlocks) blocks = kRedBlocks;
  if (blocks < 1) blocks = 1;
  reduce_partial_kernel<<<unsigned(blocks), 256, 0, st>>>(a, b, dsr, n, inv_n, mode, partial);
  FEN_CUDA(cudaGetLastError());
  reduce_final_kernel<<<1, 256, 0, st>>>(partial, int(blocks), scale, take_sqrt, out);
  FEN_CUDA(cudaGetLastError());
  g_launches += 2;
  return FEN_OK;
}

int fen_l1_loss(const float* sr, const float* hr, int64_t n, float* loss, float* dsr, void* workspace,
                int64_t workspace_bytes, void* stream) {
  int rc = check_device();
  if (rc) return rc;
  g_launches = 0;
  if (!sr || !hr || !loss || !workspace) return fail(FEN_EINVAL, "fen_l1_loss: null pointer");
  if (n < 1) return fail(FEN_EINVAL, "fen_l1_loss: empty input");
  return reduce_launch(sr, hr, dsr, n, 1.f / float(n), 0, 1.f / float(n), 0, loss, workspace, workspace_bytes,
                       static_cast<cudaStream_t>(stream));
}

int fen_psnr(const float* pred, const float* target, int64_t n, float data_range, float* psnr, void* workspace,
             int64_t workspace_bytes, void* stream) {
  int rc = check_device();
  if (rc) return rc;
  g_launches = 0;
  if (!pred || !target || !psnr || !workspace) return fail(FEN_EINVAL, "fen_psnr: null pointer");
  if (n < 1 || !(data_range > 0.f)) return fail(FEN_EINVAL, "fen_psnr: n >= 1 and data_range > 0 required");
  // scale = 1 / (n * range^2): the final kernel evaluates 10 log10(1 / (sum * scale))
  return reduce_launch(pred, target, nullptr, n, 0.f, 2, float(1.0 / (double(n) * double(data_range) * double(data_range))),
                       2, psnr, workspace, workspace_bytes, static_cast<cudaStream_t>(stream));
}

int fen_grad_norm(const float* grads, int64_t n, float* norm_out, void* workspace, int64_t workspace_bytes,
                  void* stream) {
  int rc = check_device();
  if (rc) return rc;
  g_launches = 0;
  if (!grads || !norm_out || !workspace) return fail(FEN_EINVAL, "fen_grad_norm: null pointer");
  if (n < 1) return fail(FEN_EINVAL, "fen_grad_norm: empty input");
  return reduce_launch(grads, nullptr, nullptr, n, 0.f, 1, 1.f, 1, norm_out, workspace, workspace_bytes,
                       static_cast<cudaStream_t>(stream));
}

int fen_clip_adamw_step(float* params, const float* grads, float* exp_avg, float* exp_avg_sq, int64_t n,
                        const float* total_norm, float max_norm, float lr, float beta1, float beta2, float eps,
                        float weight_decay, int step, void* stream) {
  int rc = check_device();
  if (rc) return rc;
  g_launches = 0;
  if (!params || !grads || !exp_avg || !exp_avg_sq) return fail(FEN_EINVAL, "fen_clip_adamw_step: null pointer");
  if (max_norm > 0.f && !total_norm) return fail(FEN_EINVAL, "fen_clip_adamw_step: total_norm required for clipping");
  if (n < 1 || step < 1) return fail(FEN_EINVAL, "fen_clip_adamw_step: n and step must be >= 1");
  const double bc1 = 1.0 - pow(double(beta1), double(step)), bc2 = 1.0 - pow(double(beta2), double(step));
  int64_t blocks = (n + 255) / 256;
  const int64_t cap = int64_t(num_sms()) * 16;
  if (blocks > cap) blocks = cap;
  clip_adamw_kernel<<<unsigned(blocks), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      params, grads, exp_avg, exp_avg_sq, n, total_norm, max_norm, lr, beta1, beta2, eps, weight_decay,
      float(double(lr) / bc1), float(sqrt(bc2)));
  FEN_CUDA(cudaGetLastError());
  ++g_launches;
  return FEN_OK;
}

int fen_lr_from_hr_f32(const float* hr, float* lr_f32, uint8_t* lr_u8, int B, int C, int H, int W, int bgr,
                       void* stream) {
  int rc = check_device();
  if (rc) return rc;
  g_launches = 0;
  if (B == 0) return FEN_OK;
  if (!hr || (!lr_u8 && !lr_f32)) return fail(FEN_EINVAL, "fen_lr_from_hr_f32: null pointer");
  if (B < 0 || C < 1 || H < 4 || W < 4 || (H & 3) || (W & 3))
    return fail(FEN_EINVAL, "fen_lr_from_hr_f32: H and W must be multiples of 4");
  if (reinterpret_cast<uintptr_t>(hr) & 15) return fail(FEN_EINVAL, "fen_lr_from_hr_f32: hr must be 16-byte aligned");
  const size_t total = size_t(B) * C * (H / 4) * (W / 4);
  size_t blocks = (total + 255) / 256;
  const size_t cap = size_t(num_sms()) * 32;
  if (blocks > cap) blocks = cap;
  lr_from_hr_f32_kernel<<<unsigned(blocks), 256, 0, static_cast<cudaStream_t>(stream)>>>(hr, lr_f32, lr_u8, B, C, H, W, bgr);
  FEN_CUDA(cudaGetLastError());
  ++g_launches;
  return FEN_OK;
}

int fen_sr_to_u8(const float* sr, uint8_t* out, int B, int C, int H, int W, int bgr, void* stream) {
  int rc = check_device();
  if (rc) return rc;
  g_launches = 0;
  if (B == 0) return FEN_OK;
  if (!sr || !out) return fail(FEN_EINVAL, "fen_sr_to_u8: null pointer");
  if (B < 0 || C < 1 || H < 1 || W < 1) return fail(FEN_EINVAL, "fen_sr_to_u8: bad shape");
  const size_t total = size_t(B) * H * W;
  size_t blocks = (total + 255) / 256;
  const size_t cap = size_t(num_sms()) * 32;
  if (blocks > cap) blocks = cap;
  sr_to_u8_kernel<<<unsigned(blocks), 256, 0, static_cast<cudaStream_t>(stream)>>>(sr, out, B, C, H, W, bgr);
  FEN_CUDA(cudaGetLastError());
  ++g_launches;
  return FEN_OK;
}

int fen_lr_from_hr_u8(const uint8_t* hr, uint8_t* lr_u8, float* lr_f32, int B, int H, int W, int C,
                      void* stream) {
  int rc = check_device();
  if (rc) return rc;
  g_launches = 0;
  if (B == 0) return FEN_OK;  // empty batch: nothing to do (pointers may be null)
  if (!hr || (!lr_u8 && !lr_f32)) return fail(FEN_EINVAL, "fen_lr_from_hr_u8: null pointer");
  if (B < 0 || H < 4 || W < 4 || (H & 3) || (W & 3)) return fail(FEN_EINVAL, "fen_lr_from_hr_u8: H and W must be multiples of 4");
  if (C < 1 || C > 4) return fail(FEN_EINVAL, "fen_lr_from_hr_u8: C must be in 1..4");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const size_t total = size_t(B) * (H / 4) * (W / 4);
  size_t blocks = (total + 255) / 256;
  const size_t cap = size_t(num_sms()) * 32;
  if (blocks > cap) blocks = cap;
  const bool aligned = !(reinterpret_cast<uintptr_t>(hr) & 15) && !(reinterpret_cast<uintptr_t>(lr_u8) & 3) &&
                       !(reinterpret_cast<uintptr_t>(lr_f32) & 15);
  if (C == 3 && (W & 15) == 0 && aligned) {
    const size_t quads = total / 4;
    size_t qb = (quads + 255) / 256;
    if (qb > cap) qb = cap;
    lr_from_hr_rgb4_kernel<<<unsigned(qb), 256, 0, st>>>(hr, lr_u8, lr_f32, B, H, W);
    FEN_CUDA(cudaGetLastError());
    ++g_launches;
    return FEN_OK;
  }
  switch (C) {
    case 1: lr_from_hr_kernel<1><<<unsigned(blocks), 256, 0, st>>>(hr, lr_u8, lr_f32, B, H, W); break;
    case 2: lr_from_hr_kernel<2><<<unsigned(blocks), 256, 0, st>>>(hr, lr_u8, lr_f32, B, H, W); break;
    case 3: lr_from_hr_kernel<3><<<unsigned(blocks), 256, 0, st>>>(hr, lr_u8, lr_f32, B, H, W); break;
    default: lr_from_hr_kernel<4><<<unsigned(blocks), 256, 0, st>>>(hr, lr_u8, lr_f32, B, H, W); break;
  }
  FEN_CUDA(cudaGetLastError());
  ++g_launches;
  return FEN_OK;
}

// ===================================================================== SSIM (validation metric, Stage-2 loss term)
static int64_t ssim_ws(int B, int C, int H, int W, int want_grad, int64_t* partial_bytes) {
  const int64_t tiles = int64_t((H + kSsimTile - 1) / kSsimTile) * ((W + kSsimTile - 1) / kSsimTile);
  *partial_bytes = align256(int64_t(B) * C * tiles * 4);
  return *partial_bytes + (want_grad ? 3 * align256(int64_t(B) * C * H * W * 4) : 0);
}
int64_t fen_ssim_workspace_bytes(int B, int C, int H, int W, int want_grad) {
  if (B < 1 || C < 1 || H < 1 || W < 1) { fail(FEN_EINVAL, "fen_ssim_workspace_bytes: bad shape"); return FEN_EINVAL; }
  int64_t pb;
  return ssim_ws(B, C, H, W, want_grad, &pb);
}

int fen_ssim(const float* pred, const float* target, int B, int C, int H, int W, const float* window_1d, int window_size,
             float c1, float c2, float* per_image, float* mean, float* grad_pred, void* workspace,
             int64_t workspace_bytes, void* stream) {
  int rc = check_device();
  if (rc) return rc;
  g_launches = 0;
  if (!pred || !target || !window_1d || !workspace || (!per_image && !mean && !grad_pred))
    return fail(FEN_EINVAL, "fen_ssim: null pointer");
  if (B < 1 || C < 1 || H < 1 || W < 1 || int64_t(B) * C > 65535) return fail(FEN_EINVAL, "fen_ssim: bad shape");
  if (window_size < 1 || window_size > kSsimMaxWin || !(window_size & 1))
    return fail(FEN_EINVAL, "fen_ssim: window_size must be odd and at most 11");
  int64_t pb;
  if (workspace_bytes < ssim_ws(B, C, H, W, grad_pred != nullptr, &pb)) return fail(FEN_ENOMEM, "fen_ssim: workspace too small");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  SsimParams P{};
  P.B = B; P.C = C; P.H = H; P.W = W; P.ws = window_size; P.c1 = c1; P.c2 = c2;
  for (int i = 0; i < window_size; ++i) P.g[i] = window_1d[i];
  P.tiles_x = (W + kSsimTile - 1) / kSsimTile; P.tiles_y = (H + kSsimTile - 1) / kSsimTile;
  const int tiles = P.tiles_x * P.tiles_y;
  uint8_t* wsb = static_cast<uint8_t*>(workspace);
  float* partial = reinterpret_cast<float*>(wsb);
  const int64_t plane_bytes = align256(int64_t(B) * C * H * W * 4);
  float* gmu = grad_pred ? reinterpret_cast<float*>(wsb + pb) : nullptr;
  float* gpp = grad_pred ? reinterpret_cast<float*>(wsb + pb + plane_bytes) : nullptr;
  float* gpt = grad_pred ? reinterpret_cast<float*>(wsb + pb + 2 * plane_bytes) : nullptr;
  ssim_fwd_kernel<<<dim3(tiles, B * C), 256, 0, st>>>(pred, target, P, partial, gmu, gpp, gpt);
  FEN_CUDA(cudaGetLastError());
  ++g_launches;
  if (per_image || mean) {
    ssim_reduce_kernel<<<1, 256, 0, st>>>(partial, B, C, tiles, 1.0 / (double(C) * H * W), per_image, mean);
    FEN_CUDA(cudaGetLastError());
    ++g_launches;
  }
  if (grad_pred) {
    ssim_bwd_kernel<<<dim3(tiles, B * C), 256, 0, st>>>(pred, target, gmu, gpp, gpt, P,
                                                        float(1.0 / (double(B) * C * H * W)), grad_pred);
    FEN_CUDA(cudaGetLastError());
    ++g_launches;
  }
  return FEN_OK;
}

// ===================================================================== Stage-1 step, network side
int64_t fen_packed_bwd_bytes(const fen_config* cfg) {
  Layout L;
  if (make_layout(cfg, &L)) return FEN_EINVAL;
  BwdLayout K;
  make_bwd_layout(L, &K);
  return K.total;
}

int fen_pack_weights_bwd(const fen_config* cfg, const float* params, void* packed_bwd, void* stream) {
  int rc = check_device();
  if (rc) return rc;
  Layout L;
  if ((rc = make_layout(cfg, &L))) return rc;
  if (!params || !packed_bwd) return fail(FEN_EINVAL, "fen_pack_weights_bwd: null pointer");
  BwdLayout K;
  make_bwd_layout(L, &K);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  uint8_t* kb = static_cast<uint8_t*>(packed_bwd);
  auto packT = [&](const float* w, int groups, int64_t off) -> int {
    const int total = groups * 9 * 64 * 64;
    pack_conv_T_kernel<<<(total + 255) / 256, 256, 0, st>>>(w, reinterpret_cast<bf16*>(kb + off), groups);
    FEN_CUDA(cudaGetLastError());
    return FEN_OK;
  };
  pack_T_all_kernel<<<dim3(36, 2 * L.n_rcab + L.G + 1), 256, 0, st>>>(params, kb, L, K);
  FEN_CUDA(cudaGetLastError());
  for (int s = 0; s < 2; ++s)
    if ((rc = packT(params + L.p_up[s], 4, K.up[s]))) return rc;
  pack_last_T_kernel<<<(9 * 64 * 64 + 255) / 256, 256, 0, st>>>(params + L.p_last_w, reinterpret_cast<bf16*>(kb + K.last));
  FEN_CUDA(cudaGetLastError());
  FEN_CUDA(cudaMemsetAsync(kb + K.zeros, 0, 256, st));
  return FEN_OK;
}

int64_t fen_step_workspace_bytes(const fen_config* cfg, int B, int H, int W) {
  Layout L;
  if (make_layout(cfg, &L)) return FEN_EINVAL;
  if (B < 1 || H < 1 || W < 1) { fail(FEN_EINVAL, "bad batch / size"); return FEN_EINVAL; }
  StepWs ws;
  make_step_ws(L, B, H, W, &ws);
  return ws.total;
}

static int step_args(const fen_config* cfg, Layout* L, StepWs* ws, int B, int H, int W, int64_t bytes, const char* who) {
  int rc = check_device();
  if (rc) return rc;
  g_launches = 0;
  if ((rc = make_layout(cfg, L))) return rc;
  if (B < 1 || H < 1 || W < 1) return fail(FEN_EINVAL, std::string(who) + ": empty input");
  if (W > 336 || H > 4096)   // (the CUDA-core row kernels of the backward stage 4 W + 4 floats x 9 rows in 48 KB of shared memory)
    return fail(FEN_EINVAL, std::string(who) + ": the training path takes LR inputs up to 336 columns x 4096 rows");
  make_step_ws(*L, B, H, W, ws);
  if (bytes < ws->total) return fail(FEN_ENOMEM, std::string(who) + ": workspace too small");
  return FEN_OK;
}

int fen_forward_train(const fen_config* cfg, const void* packed, const float* x, float* out, int B, int H, int W,
                      void* step_workspace, int64_t step_workspace_bytes, void* stream) {
  Layout L;
  StepWs ws;
  int rc = step_args(cfg, &L, &ws, B, H, W, step_workspace_bytes, "fen_forward_train");
  if (rc) return rc;
  if (!packed || !x || !out || !step_workspace) return fail(FEN_EINVAL, "fen_forward_train: null pointer");
  return step_forward(cfg, L, static_cast<const uint8_t*>(packed), x, out, B, H, W,
                      static_cast<uint8_t*>(step_workspace), ws, static_cast<cudaStream_t>(stream));
}

int fen_backward_stages(const fen_config* cfg, const void* packed, const void* packed_bwd, const float* x,
                        const float* dout, float* grads, int B, int H, int W, void* step_workspace,
                        int64_t step_workspace_bytes, int stage_begin, int stage_end, void* stream) {
  Layout L;
  StepWs ws;
  int rc = step_args(cfg, &L, &ws, B, H, W, step_workspace_bytes, "fen_backward");
  if (rc) return rc;
  if (!packed || !packed_bwd || !x || !dout || !grads || !step_workspace)
    return fail(FEN_EINVAL, "fen_backward: null pointer");
  if (stage_begin < 0 || stage_end > bwd_num_stages(L) || stage_begin >= stage_end)
    return fail(FEN_EINVAL, "fen_backward_stages: bad stage range");
  BwdLayout K;
  make_bwd_layout(L, &K);
  return step_backward(cfg, L, static_cast<const uint8_t*>(packed), static_cast<const uint8_t*>(packed_bwd), K, x, dout,
                       grads, B, H, W, static_cast<uint8_t*>(step_workspace), ws, static_cast<cudaStream_t>(stream),
                       stage_begin, stage_end);
}

int fen_backward(const fen_config* cfg, const void* packed, const void* packed_bwd, const float* x, const float* dout,
                 float* grads, int B, int H, int W, void* step_workspace, int64_t step_workspace_bytes, void* stream) {
  Layout L;
  if (make_layout(cfg, &L)) return FEN_EINVAL;
  return fen_backward_stages(cfg, packed, packed_bwd, x, dout, grads, B, H, W, step_workspace, step_workspace_bytes, 0,
                             bwd_num_stages(L), stream);
}

int fen_backward_num_stages(const fen_config* cfg) {
  Layout L;
  if (make_layout(cfg, &L)) return FEN_EINVAL;
  return bwd_num_stages(L);
}

int fen_backward_stage_range(const fen_config* cfg, int stage, int64_t* begin, int64_t* count) {
  Layout L;
  if (make_layout(cfg, &L)) return FEN_EINVAL;
  if (!begin || !count || stage < 0 || stage >= bwd_num_stages(L)) return fail(FEN_EINVAL, "fen_backward_stage_range: bad arguments");
  bwd_stage_range(L, stage, begin, count);
  return FEN_OK;
}

}  // extern "C"
