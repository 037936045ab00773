"""Developer script: cycle counters of the persistent body kernel (full model, batch from argv)."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import fsr_b200
from fsr_b200 import _lib
from oracle import weights
lib = _lib.load(); dev = torch.device("cuda:0")
B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
cfg = dict(num_groups=6, blocks_per_group=10)
m = fsr_b200.FaceEnhanceNet(**cfg); m.load_state_dict(weights.make_state_dict(0, "T1", **cfg)); m = m.to(dev).eval()
x = torch.rand(B, 3, 64, 64, device=dev)
with torch.no_grad():
    for _ in range(2): m(x)
    dbg = torch.zeros(148, 16, dtype=torch.int64, device=dev)
    lib.fen_debug_set_counters(ctypes.c_void_p(dbg.data_ptr()))
    m(x); torch.cuda.synchronize()
    lib.fen_debug_set_counters(None)
d = dbg.cpu().double(); d = d[d[:, 12] > 0]
names = ["epi wait flags", "(unused)", "epi SE compute", "epi conv2 tile loops", "epi conv2 acc waits", "aux total",
         "mma wait acc_empty", "mma wait full/w", "mma issue", "mma total", "epi wait acc_full", "epi wait done", "epi total",
         "epi other tile loops", "epi other acc waits", "mma conv2 layers"]
print(f"{len(d)} CTAs, 127 layers (60 conv2 = SE layers, 67 others); cycles per layer, mean over CTAs:")
for i, n in enumerate(names):
    div = 60 if i in (0, 2, 3, 4, 15) else 67 if i in (13, 14) else 127
    print(f"  {n:22s} {d[:, i].mean().item() / div:9.0f}   (per {'conv2 layer' if div == 60 else 'other layer' if div == 67 else 'layer'})")
