// Hardware probe 2 (test infrastructure, not product): tcgen05.mma issue/throughput floor as a
// function of shape, operand source (SS vs TS) and cta_group.  Timing only, data is garbage.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o umma_probe2 umma_probe2.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_bf16.h>
#include "../face-super-resolution_b200/csrc/ptx_sm100.cuh"
using namespace fen;

#define CK(x)                                                                         \
  do {                                                                                \
    cudaError_t e_ = (x);                                                             \
    if (e_ != cudaSuccess) {                                                          \
      printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); \
      exit(1);                                                                        \
    }                                                                                 \
  } while (0)

__device__ __forceinline__ void umma_bf16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc,
                                             uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}\n" ::"r"(tmem_d),
      "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(acc)
      : "memory");
}
__device__ __forceinline__ void umma_bf16_ss2(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc,
                                              uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc)
      : "memory");
}

struct Cfg { int M, N, mode; };  // mode 0: SS shifted views, 1: SS fixed aligned view, 2: TS
constexpr int NCFG = 16;
__constant__ Cfg c_cfg[NCFG] = {{128, 64, 0},  {128, 64, 1}, {128, 32, 0}, {128, 96, 0},
                                {128, 128, 0}, {64, 64, 0},  {64, 128, 0}, {64, 256, 0},
                                {128, 64, 2},  {128, 128, 2}, {128, 256, 2}, {128, 16, 2},
                                {128, 64, 3}, {128, 64, 4}, {128, 32, 4}, {128, 64, 5}};

constexpr int SMEM1 = 32768 + 32768 + 1024;

__global__ void __launch_bounds__(128, 1) time1(long long* cycles) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  const int tid = threadIdx.x;
  for (int i = tid; i < 65536 / 16; i += 128) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
  if (tid < 32) tmem_alloc(&slot, 512);
  if (tid == 0) { mbar_init(&bar, 1); fence_mbar_init(); }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = slot;
  if (tid == 0) {
    uint32_t phase = 0;
    const uint32_t a0 = smem_u32(smem), b0 = smem_u32(smem + 32768);
    for (int c = 0; c < NCFG; ++c) {
      const Cfg cf = c_cfg[c];
      const uint32_t idesc = umma_idesc_bf16(cf.M, cf.N);
      uint64_t ad[36], bd[4];
#pragma unroll
      for (int i = 0; i < 36; ++i) {
        int tap = i >> 2, k = i & 3;
        int j = cf.mode != 1 ? (tap / 3) * 33 + (tap % 3) : 0;
        ad[i] = umma_smem_desc(a0 + j * 128 + k * 32, 16, 1024, UMMA_LAYOUT_SW128, 0);
      }
#pragma unroll
      for (int k = 0; k < 4; ++k) bd[k] = umma_smem_desc(b0 + k * 32, 16, 1024, UMMA_LAYOUT_SW128, 0);
      const int REP = 16;
      tc_fence_after();
      long long t0 = clock64();
      for (int r = 0; r < REP; ++r) {
        if (cf.mode == 2) {
#pragma unroll
          for (int i = 0; i < 36; ++i) umma_bf16_ts(tmem, tmem + 256 + (i & 3) * 8, bd[i & 3], idesc, 1);
        } else if (cf.mode == 3) {
#pragma unroll
          for (int i = 0; i < 36; ++i) umma_bf16_ss(tmem + (i & 1) * 64, ad[i], bd[i & 3], idesc, 1);
        } else if (cf.mode == 4) {
#pragma unroll
          for (int i = 0; i < 36; ++i) umma_bf16_ss(tmem + (i & 3) * 64, ad[i], bd[i & 3], idesc, 1);
        } else if (cf.mode == 5) {
#pragma unroll
          for (int i = 0; i < 36; ++i) umma_bf16_ts(tmem + (i & 3) * 64, tmem + 256 + (i & 3) * 8, bd[i & 3], idesc, 1);
        } else {
#pragma unroll
          for (int i = 0; i < 36; ++i) umma_bf16_ss(tmem, ad[i], bd[i & 3], idesc, 1);
        }
      }
      umma_commit(&bar);
      mbar_wait(&bar, phase);
      phase ^= 1;
      long long t1 = clock64();
      cycles[c] = (t1 - t0) / REP;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (tid < 32) tmem_dealloc(tmem, 512);
}

// ---- 2-CTA: cta_group::2, M = 256 (128 rows per CTA), N in {64,128,256} (N/2 rows of B per CTA)
constexpr int NCFG2 = 4;
__constant__ int c_n2[NCFG2] = {64, 128, 256, 32};

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128, 1) time2(long long* cycles) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  const int tid = threadIdx.x;
  uint32_t rank;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
  for (int i = tid; i < 65536 / 16; i += 128) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
  if (tid < 32) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&slot)), "r"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  if (tid == 0) { mbar_init(&bar, 1); fence_mbar_init(); }
  fence_proxy_async_smem();
  tc_fence_before();
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
  tc_fence_after();
  const uint32_t tmem = slot;
  if (rank == 0 && tid == 0) {
    uint32_t phase = 0;
    const uint32_t a0 = smem_u32(smem), b0 = smem_u32(smem + 32768);
    for (int c = 0; c < NCFG2; ++c) {
      const int N = c_n2[c];
      const uint32_t idesc = umma_idesc_bf16(256, N);
      uint64_t ad[36], bd[4];
#pragma unroll
      for (int i = 0; i < 36; ++i) {
        int tap = i >> 2, k = i & 3;
        int j = (tap / 3) * 33 + (tap % 3);
        ad[i] = umma_smem_desc(a0 + j * 128 + k * 32, 16, 1024, UMMA_LAYOUT_SW128, 0);
      }
#pragma unroll
      for (int k = 0; k < 4; ++k) bd[k] = umma_smem_desc(b0 + k * 32, 16, 1024, UMMA_LAYOUT_SW128, 0);
      const int REP = 16;
      tc_fence_after();
      long long t0 = clock64();
      for (int r = 0; r < REP; ++r) {
#pragma unroll
        for (int i = 0; i < 36; ++i) umma_bf16_ss2(tmem, ad[i], bd[i & 3], idesc, 1);
      }
      asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
      mbar_wait(&bar, phase);
      phase ^= 1;
      long long t1 = clock64();
      cycles[c] = (t1 - t0) / REP;
    }
  }
  tc_fence_before();
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
  if (tid < 32)
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512) : "memory");
}

int main() {
  CK(cudaSetDevice(0));
  long long* d;
  CK(cudaMalloc(&d, 64 * 8));
  CK(cudaMemset(d, 0, 64 * 8));
  CK(cudaFuncSetAttribute(time1, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM1));
  CK(cudaFuncSetAttribute(time2, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM1));
  time1<<<1, 128, SMEM1>>>(d);
  CK(cudaGetLastError());
  CK(cudaDeviceSynchronize());
  long long h[64];
  CK(cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost));
  static const Cfg cfg[NCFG] = {{128, 64, 0},  {128, 64, 1}, {128, 32, 0}, {128, 96, 0},
                                {128, 128, 0}, {64, 64, 0},  {64, 128, 0}, {64, 256, 0},
                                {128, 64, 2},  {128, 128, 2}, {128, 256, 2}, {128, 16, 2},
                                {128, 64, 3}, {128, 64, 4}, {128, 32, 4}, {128, 64, 5}};
  const char* mn[6] = {"SS shifted", "SS fixed  ", "TS (A tmem)", "SS 2 accumulators", "SS 4 accumulators", "TS 4 accumulators"};
  for (int i = 0; i < NCFG; ++i)
    printf("cta_group::1 M=%3d N=%3d %s : %6lld cyc / 36 MMA = %6.1f per MMA (ideal %5.1f)\n", cfg[i].M,
           cfg[i].N, mn[cfg[i].mode], h[i], h[i] / 36.0, cfg[i].M * cfg[i].N * 16 / 4096.0);
  time2<<<2, 128, SMEM1>>>(d + 32);
  CK(cudaGetLastError());
  CK(cudaDeviceSynchronize());
  CK(cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost));
  static const int n2[NCFG2] = {64, 128, 256, 32};
  for (int i = 0; i < NCFG2; ++i)
    printf("cta_group::2 M=256 N=%3d SS shifted : %6lld cyc / 36 MMA = %6.1f per MMA (ideal %5.1f)\n", n2[i],
           h[32 + i], h[32 + i] / 36.0, 256 * n2[i] * 16 / 8192.0);
  return 0;
}
