"""Developer script: times the 64->64 conv kernel alone (C-ABI call) on a B200."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from fsr_b200 import _lib
lib = _lib.load()
dev = torch.device("cuda:0")
B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
epi = int(sys.argv[2]) if len(sys.argv) > 2 else 5
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 40
act = [torch.randn(B, 64, 64, 64, device=dev).mul_(0.3).to(torch.bfloat16) for _ in range(2)]
w = torch.randn(64, 64, 3, 3, device=dev) * 0.05
wp = torch.empty(9 * 64 * 64, dtype=torch.bfloat16, device=dev)
bias = torch.zeros(64, device=dev); slope = torch.full((64,), 0.25, device=dev)
sums = torch.zeros(B, 64, device=dev)
st = torch.cuda.current_stream().cuda_stream
_lib.check(lib.fen_pack_conv3x3(w.data_ptr(), 64, 64, wp.data_ptr(), st), "pack")
def once(i):
    _lib.check(lib.fen_conv3x3_c64(act[i & 1].data_ptr(), wp.data_ptr(), bias.data_ptr(), slope.data_ptr(),
                                   act[i & 1].data_ptr(), sums.data_ptr(), act[(i + 1) & 1].data_ptr(), B, 64, 64, epi, st), "conv")
for i in range(6): once(i)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for i in range(reps): once(i)
e1.record(); torch.cuda.synchronize()
us = e0.elapsed_time(e1) / reps * 1e3
fl = 2.0 * 4096 * 64 * 64 * 9 * B
ntile = (B * 33 + 147) // 148
print(f"conv64 B={B} epi={epi}: {us:.1f} us/launch, {fl / us / 1e6:.0f} TFLOP/s, {us * 1.965e3 / ntile:.0f} cyc/tile (if 1.965 GHz, {ntile} tiles/CTA)")
if os.environ.get("FEN_DBG"):
    import ctypes
    dbg = torch.zeros(148, 8, dtype=torch.int64, device=dev)
    lib.fen_debug_set_counters(ctypes.c_void_p(dbg.data_ptr()))
    once(0); torch.cuda.synchronize()
    lib.fen_debug_set_counters(None)
    d = dbg.cpu()[:141].double()
    names = ["epi tmem ld+release", "MMA issue loops", "MMA wait acc_empty", "MMA wait TMA full", "MMA thread total", "tiles", "epi wait acc_full", "epi total"]
    for i, n in enumerate(names):
        print(f"  {n:22s} mean {d[:, i].mean().item():9.0f}  max {d[:, i].max().item():9.0f}")
