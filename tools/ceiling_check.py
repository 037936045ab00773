"""External check of the N = 64 ceiling the body kernel is measured against: what do NVIDIA's own
libraries reach on the same shape, on the same box?
  * cuBLAS (torch.matmul, bf16):  [M = 262144, K = 576] x [K = 576, N = 64]  - one 64->64 3x3 conv at batch 64 as a GEMM
    (M = 64 images x 4096 pixels, K = 9 taps x 64 channels), the im2col matrix taken as given (its construction is free here)
  * cuDNN (F.conv2d, bf16, channels_last): the real 64->64 3x3 convolution on [64, 64, 64, 64]
Both are timed with CUDA events over back-to-back launches on rotating buffers (> L2 in total)."""
import json, sys
import torch
import torch.nn.functional as F
dev = torch.device("cuda:0")
ev = lambda: torch.cuda.Event(enable_timing=True)
def timeit(fn, n=50, warm=5):
    for i in range(warm): fn(i)
    torch.cuda.synchronize(); e0, e1 = ev(), ev(); e0.record()
    for i in range(n): fn(i)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
out = {}
for B in (64, 256):
    M, K, N = B * 4096, 576, 64
    flop = 2.0 * M * K * N
    a = [torch.randn(M, K, device=dev, dtype=torch.bfloat16) for _ in range(3)]
    w = torch.randn(K, N, device=dev, dtype=torch.bfloat16)
    ms = timeit(lambda i: torch.matmul(a[i % 3], w))
    out[f"cublas_gemm_M{M}_N64_K576"] = {"ms": round(ms, 4), "tflops": round(flop / ms / 1e9, 1)}
    wt = torch.randn(N, K, device=dev, dtype=torch.bfloat16)
    ms = timeit(lambda i: torch.matmul(a[i % 3], wt.t()))
    out[f"cublas_gemm_M{M}_N64_K576_Bt"] = {"ms": round(ms, 4), "tflops": round(flop / ms / 1e9, 1)}
    del a
    x = [torch.randn(B, 64, 64, 64, device=dev, dtype=torch.bfloat16).contiguous(memory_format=torch.channels_last) for _ in range(4)]
    cw = torch.randn(64, 64, 3, 3, device=dev, dtype=torch.bfloat16).contiguous(memory_format=torch.channels_last)
    cb = torch.randn(64, device=dev, dtype=torch.bfloat16)
    torch.backends.cudnn.benchmark = True
    ms = timeit(lambda i: F.conv2d(x[i % 4], cw, cb, padding=1))
    out[f"cudnn_conv3x3_c64_B{B}_bf16_nhwc"] = {"ms": round(ms, 4), "tflops": round(flop / ms / 1e9, 1)}
    del x
    torch.cuda.empty_cache()
# N = 128 and N = 256 for comparison (the same M, K): where the library stops being N-limited
for N in (128, 256):
    M, K = 64 * 4096, 576
    a = [torch.randn(M, K, device=dev, dtype=torch.bfloat16) for _ in range(3)]
    w = torch.randn(K, N, device=dev, dtype=torch.bfloat16)
    ms = timeit(lambda i: torch.matmul(a[i % 3], w))
    out[f"cublas_gemm_M{M}_N{N}_K576"] = {"ms": round(ms, 4), "tflops": round(2.0 * M * K * N / ms / 1e9, 1)}
    del a
print(json.dumps(out, indent=1))
