// Hardware probe (test infrastructure, not product): checks on a real B200 the tcgen05 / TMA
// addressing facts the convolution kernel design relies on, and times the MMA issue rate.
//
//  T1  SW128 K-major A tile written by TMA, UMMA descriptor start shifted by j pixel rows (j*128 B,
//      NOT 1024-aligned), base_offset = 0.
//  T2  same, base_offset = (start >> 7) & 7.
//  T3  no-swizzle "chunk-major" A tile ([8 chunks][npix][16 B]) written by generic stores, start
//      shifted by j*16 B.
//  T4  same layout written by a 3-D TMA box (dims 8ch x npix x 8chunks).
//  T5  cycles for 36 back-to-back MMAs (9 taps x K=64) at N = 64 / 128 / 256, both layouts.
//
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o umma_probe umma_probe.cu
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>
#include <cuda_bf16.h>
#include "../face-super-resolution_b200/csrc/ptx_sm100.cuh"

using namespace fen;

#define CK(x)                                                                      \
  do {                                                                             \
    cudaError_t e_ = (x);                                                          \
    if (e_ != cudaSuccess) {                                                       \
      printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); \
      exit(1);                                                                     \
    }                                                                              \
  } while (0)

constexpr int NPIX = 256;
constexpr int NSHIFT = 8;
__constant__ int c_shifts[NSHIFT] = {0, 1, 2, 3, 7, 8, 66, 67};
static const int h_shifts[NSHIFT] = {0, 1, 2, 3, 7, 8, 66, 67};
constexpr int NTEST = 4 * NSHIFT;
constexpr int NTIME = 8;

constexpr int OFF_A_SW = 0;                       // 256 px * 128 B = 32 KB
constexpr int OFF_B_SW = 32768 + 1024;                  // up to 256 rows * 128 B = 32 KB
constexpr int OFF_A_NS = 65536 + 1024;                  // 8 * 256 * 16 = 32 KB (generic)
constexpr int OFF_A_NT = 98304 + 1024;                  // 32 KB (TMA 3-D)
constexpr int OFF_B_NS = 131072 + 1024;                 // 8 * 256 * 16 = 32 KB
constexpr int SMEM_BYTES = 163840 + 3072;

__global__ void __launch_bounds__(128, 1)
probe_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmW,
             const __grid_constant__ CUtensorMap tmX3, const __nv_bfloat16* __restrict__ X,
             const __nv_bfloat16* __restrict__ W, float* __restrict__ out,
             long long* __restrict__ cycles) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t tma_bar, mma_bar;
  __shared__ uint32_t tmem_slot;
  const int tid = threadIdx.x, warp = tid >> 5;

  if (warp == 0) tmem_alloc(&tmem_slot, 256);
  if (tid == 0) {
    mbar_init(&tma_bar, 1);
    mbar_init(&mma_bar, 1);
    fence_mbar_init();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_slot;

  if (tid == 0) {
    mbar_expect_tx(&tma_bar, 32768 + 8192 + 32768);
    tma_load_2d(&tmX, &tma_bar, smem + OFF_A_SW + 512, 0, 0);
    tma_load_2d(&tmX, &tma_bar, smem + OFF_A_SW + 512 + 16384, 0, 128);
    tma_load_2d(&tmW, &tma_bar, smem + OFF_B_SW, 0, 0);
    tma_load_3d(&tmX3, &tma_bar, smem + OFF_A_NT, 0, 0, 0);
  }
  // generic-proxy fill of the chunk-major copies: [chunk][pix][8 bf16]
  for (int i = tid; i < NPIX * 8; i += 128) {
    int pix = i >> 3, ch = i & 7;
    uint4 v = *reinterpret_cast<const uint4*>(X + pix * 64 + ch * 8);
    *reinterpret_cast<uint4*>(smem + OFF_A_NS + (ch * NPIX + pix) * 16) = v;
  }
  for (int i = tid; i < 256 * 8; i += 128) {
    int row = i >> 3, ch = i & 7;
    uint4 v = *reinterpret_cast<const uint4*>(W + (row & 63) * 64 + ch * 8);
    *reinterpret_cast<uint4*>(smem + OFF_B_NS + (ch * 256 + row) * 16) = v;
  }
  // rows 64..255 of the swizzled B buffer are only used by the timing runs: fill with zeros
  for (int i = tid; i < (32768 - 8192) / 16; i += 128)
    *reinterpret_cast<uint4*>(smem + OFF_B_SW + 8192 + i * 16) = make_uint4(0, 0, 0, 0);
  fence_proxy_async_smem();
  __syncthreads();
  mbar_wait(&tma_bar, 0);

  const uint32_t idesc64 = umma_idesc_bf16(128, 64);
  uint32_t phase = 0;
  for (int t = 0; t < NTEST; ++t) {
    const int mode = t / NSHIFT, j = c_shifts[t % NSHIFT];
    if (tid == 0) {
      tc_fence_after();
      for (int k = 0; k < 4; ++k) {
        uint64_t ad, bd;
        if (mode == 0 || mode == 1) {
          uint32_t a = smem_u32(smem + OFF_A_SW) + 512 + j * 128 + k * 32;
          uint32_t bo = (mode == 1) ? ((a >> 7) & 7) : 0;
          ad = umma_smem_desc(a, 16, 1024, UMMA_LAYOUT_SW128, bo);
          bd = umma_smem_desc(smem_u32(smem + OFF_B_SW) + k * 32, 16, 1024, UMMA_LAYOUT_SW128, 0);
        } else {
          uint32_t abase = smem_u32(smem + (mode == 2 ? OFF_A_NS : OFF_A_NT));
          ad = umma_smem_desc(abase + (2 * k) * NPIX * 16 + j * 16, NPIX * 16, 128, UMMA_LAYOUT_NONE, 0);
          bd = umma_smem_desc(smem_u32(smem + OFF_B_NS) + (2 * k) * 256 * 16, 256 * 16, 128,
                              UMMA_LAYOUT_NONE, 0);
        }
        umma_bf16_ss(tmem, ad, bd, idesc64, k > 0);
      }
      umma_commit(&mma_bar);
    }
    mbar_wait(&mma_bar, phase);
    phase ^= 1;
    tc_fence_after();
    for (int half = 0; half < 2; ++half) {
      uint32_t v[32];
      tmem_ld_32x32(tmem + (uint32_t(warp * 32) << 16) + half * 32, v);
      tmem_ld_wait();
      for (int c = 0; c < 32; ++c)
        out[(size_t(t) * 128 + tid) * 64 + half * 32 + c] = __uint_as_float(v[c]);
    }
    tc_fence_before();
    __syncthreads();
  }

  // ---- timing: 36 MMAs (9 taps x 4 k-steps) x REP, single issuing thread
  if (tid == 0) {
    const int REP = 16;
    for (int cfg = 0; cfg < NTIME; ++cfg) {
      const int N = (cfg & 3) == 0 ? 64 : (cfg & 3) == 1 ? 128 : (cfg & 3) == 2 ? 256 : 16;
      const bool sw = cfg < 4;
      const uint32_t idesc = umma_idesc_bf16(128, N);
      tc_fence_after();
      long long t0 = clock64();
      for (int r = 0; r < REP; ++r) {
        for (int tap = 0; tap < 9; ++tap) {
          const int j = (tap / 3) * 33 + (tap % 3);  // shifted views, stays inside 256 px
          for (int k = 0; k < 4; ++k) {
            uint64_t ad, bd;
            if (sw) {
              ad = umma_smem_desc(smem_u32(smem + OFF_A_SW) + j * 128 + k * 32, 16, 1024,
                                  UMMA_LAYOUT_SW128, 0);
              bd = umma_smem_desc(smem_u32(smem + OFF_B_SW) + k * 32, 16, 1024, UMMA_LAYOUT_SW128, 0);
            } else {
              ad = umma_smem_desc(smem_u32(smem + OFF_A_NS) + (2 * k) * NPIX * 16 + j * 16, NPIX * 16,
                                  128, UMMA_LAYOUT_NONE, 0);
              bd = umma_smem_desc(smem_u32(smem + OFF_B_NS) + (2 * k) * 256 * 16, 256 * 16, 128,
                                  UMMA_LAYOUT_NONE, 0);
            }
            umma_bf16_ss(tmem, ad, bd, idesc, (tap | k | r) > 0);
          }
        }
      }
      umma_commit(&mma_bar);
      mbar_wait(&mma_bar, phase);
      phase ^= 1;
      long long t1 = clock64();
      cycles[cfg] = (t1 - t0) / REP;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 256);
}

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                             const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                             CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                             CUtensorMapFloatOOBfill);

int main() {
  CK(cudaSetDevice(0));
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, 0));
  printf("device: %s sm_%d%d, %d SMs, clock %d kHz\n", prop.name, prop.major, prop.minor,
         prop.multiProcessorCount, prop.clockRate);
  EncodeFn encode = nullptr;
  cudaDriverEntryPointQueryResult qres;
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", (void**)&encode, cudaEnableDefault, &qres));
  if (!encode) { printf("no cuTensorMapEncodeTiled\n"); return 1; }

  std::vector<__nv_bfloat16> hX(NPIX * 64), hW(64 * 64);
  std::vector<float> fX(NPIX * 64), fW(64 * 64);
  srand(1);
  for (int i = 0; i < NPIX * 64; ++i) {
    float v = float((rand() % 255) - 127) / 64.f;
    hX[i] = __float2bfloat16(v); fX[i] = __bfloat162float(hX[i]);
  }
  for (int i = 0; i < 64 * 64; ++i) {
    float v = float((rand() % 255) - 127) / 128.f;
    hW[i] = __float2bfloat16(v); fW[i] = __bfloat162float(hW[i]);
  }
  __nv_bfloat16 *dX, *dW; float* dOut; long long* dCyc;
  CK(cudaMalloc(&dX, hX.size() * 2)); CK(cudaMalloc(&dW, hW.size() * 2));
  CK(cudaMalloc(&dOut, size_t(NTEST) * 128 * 64 * 4)); CK(cudaMalloc(&dCyc, NTIME * 8));
  CK(cudaMemcpy(dX, hX.data(), hX.size() * 2, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dW, hW.data(), hW.size() * 2, cudaMemcpyHostToDevice));
  CK(cudaMemset(dOut, 0, size_t(NTEST) * 128 * 64 * 4));

  CUtensorMap tmX, tmW, tmX3;
  {
    cuuint64_t dims[2] = {64, NPIX}; cuuint64_t strides[1] = {128};
    cuuint32_t box[2] = {64, 128}; cuuint32_t es[2] = {1, 1};
    CUresult r = encode(&tmX, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, dX, dims, strides, box, es,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                        CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("encode X: %d\n", int(r));
  }
  {
    cuuint64_t dims[2] = {64, 64}; cuuint64_t strides[1] = {128};
    cuuint32_t box[2] = {64, 64}; cuuint32_t es[2] = {1, 1};
    CUresult r = encode(&tmW, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, dW, dims, strides, box, es,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                        CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("encode W: %d\n", int(r));
  }
  {
    cuuint64_t dims[3] = {8, NPIX, 8}; cuuint64_t strides[2] = {128, 16};
    cuuint32_t box[3] = {8, NPIX, 8}; cuuint32_t es[3] = {1, 1, 1};
    CUresult r = encode(&tmX3, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, dX, dims, strides, box, es,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                        CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("encode X3 (chunk-major 3-D): %d\n", int(r));
  }
  CK(cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
  probe_kernel<<<1, 128, SMEM_BYTES>>>(tmX, tmW, tmX3, dX, dW, dOut, dCyc);
  CK(cudaGetLastError());
  CK(cudaDeviceSynchronize());

  std::vector<float> hOut(size_t(NTEST) * 128 * 64);
  long long hCyc[NTIME];
  CK(cudaMemcpy(hOut.data(), dOut, hOut.size() * 4, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(hCyc, dCyc, sizeof(hCyc), cudaMemcpyDeviceToHost));
  const char* names[4] = {"T1 SW128/TMA  base_offset=0   ", "T2 SW128/TMA  base_offset=addr",
                          "T3 noswz chunk-major generic  ", "T4 noswz chunk-major 3-D TMA  "};
  for (int t = 0; t < NTEST; ++t) {
    int mode = t / NSHIFT, j = h_shifts[t % NSHIFT];
    double maxerr = 0;
    for (int m = 0; m < 128; ++m)
      for (int n = 0; n < 64; ++n) {
        double ref = 0;
        for (int c = 0; c < 64; ++c) ref += double(fX[(j + m) * 64 + c]) * double(fW[n * 64 + c]);
        double e = fabs(ref - double(hOut[(size_t(t) * 128 + m) * 64 + n]));
        if (e > maxerr) maxerr = e;
      }
    printf("%s shift %3d px : max|err| = %.3e %s\n", names[mode], j, maxerr, maxerr < 1e-2 ? "OK" : "FAIL");
  }
  const char* tn[NTIME] = {"SW128 N=64", "SW128 N=128", "SW128 N=256", "SW128 N=16",
                           "NOSWZ N=64", "NOSWZ N=128", "NOSWZ N=256", "NOSWZ N=16"};
  for (int i = 0; i < NTIME; ++i)
    printf("T5 %-12s : %lld cycles per 36 MMAs (%.1f / MMA)\n", tn[i], hCyc[i], hCyc[i] / 36.0);
  return 0;
}
