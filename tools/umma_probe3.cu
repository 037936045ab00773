// Hardware probe 3 (test infrastructure): does issuing tcgen05.mma from 2 or 4 warps concurrently
// (independent accumulators) beat the ~50-cycle per-instruction floor seen from a single thread?
#include <cstdio>
#include <cstdlib>
#include <cuda_bf16.h>
#include "../face-super-resolution_b200/csrc/ptx_sm100.cuh"
using namespace fen;
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1);} } while (0)
constexpr int SMEM1 = 65536 + 1024;
__global__ void __launch_bounds__(128, 1) time3(long long* cycles) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bar[4];
  __shared__ uint32_t slot;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  for (int i = tid; i < 65536 / 16; i += 128) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
  if (tid < 32) tmem_alloc(&slot, 512);
  if (tid == 0) { for (int i = 0; i < 4; ++i) mbar_init(&bar[i], 1); fence_mbar_init(); }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = slot;
  const uint32_t a0 = smem_u32(smem), b0 = smem_u32(smem + 32768);
  uint32_t phase = 0;  // per-thread: parity of the next completion of bar[warp]
  for (int nw = 1; nw <= 4; nw *= 2) {
    for (int ncfg = 0; ncfg < 2; ++ncfg) {
      const int N = ncfg == 0 ? 64 : 32;
      const uint32_t idesc = umma_idesc_bf16(128, N);
      __syncthreads();
      long long t0 = clock64();
      if (warp < nw && lane == 0) {
        uint64_t ad[36], bd[4];
#pragma unroll
        for (int i = 0; i < 36; ++i) {
          int tap = i >> 2, k = i & 3;
          int j = (tap / 3) * 33 + (tap % 3);
          ad[i] = umma_smem_desc(a0 + j * 128 + k * 32, 16, 1024, UMMA_LAYOUT_SW128, 0);
        }
#pragma unroll
        for (int k = 0; k < 4; ++k) bd[k] = umma_smem_desc(b0 + k * 32, 16, 1024, UMMA_LAYOUT_SW128, 0);
        tc_fence_after();
        for (int r = 0; r < 16; ++r) {
#pragma unroll
          for (int i = 0; i < 36; ++i) umma_bf16_ss(tmem + warp * 64, ad[i], bd[i & 3], idesc, 1);
        }
        umma_commit(&bar[warp]);
        mbar_wait(&bar[warp], phase);
        phase ^= 1;
      }
      __syncthreads();
      long long t1 = clock64();
      if (tid == 0) cycles[(nw == 1 ? 0 : nw == 2 ? 1 : 2) * 2 + ncfg] = (t1 - t0) / 16;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (tid < 32) tmem_dealloc(tmem, 512);
}
int main() {
  long long* d; CK(cudaMalloc(&d, 64 * 8)); CK(cudaMemset(d, 0, 64 * 8));
  CK(cudaFuncSetAttribute(time3, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM1));
  time3<<<1, 128, SMEM1>>>(d); CK(cudaGetLastError()); CK(cudaDeviceSynchronize());
  long long h[64]; CK(cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost));
  for (int w = 0; w < 3; ++w) for (int n = 0; n < 2; ++n)
    printf("%d issuing warps, M=128 N=%d: %lld cycles per (36 MMAs per warp) -> %.1f cyc per MMA overall\n",
           1 << w, n == 0 ? 64 : 32, h[w * 2 + n], h[w * 2 + n] / (36.0 * (1 << w)));
  return 0;
}
