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
d = dbg.cpu().double(); d = d[d[:, 5] > 0]
names = ["prod wait wfree", "prod wait flags", "prod SE compute", "prod fused boxes", "prod plain/other", "prod total",
         "mma wait acc_empty", "mma wait full/w", "mma issue", "mma total", "epi wait acc_full", "epi wait done", "epi total", "mma fused layers", "mma plain layers", "mma fused wait full"]
print(f"{len(d)} CTAs, 127 layers; cycles per layer (mean over CTAs / max CTA):")
for i, n in enumerate(names):
    print(f"  {n:20s} {d[:, i].mean().item() / 127:9.0f} {d[:, i].max().item() / 127:9.0f}")
nf = 6 * 10  # fused layers: 9 conv1 + 1 group conv per group
print(f"per FUSED layer (60): mma {d[:, 13].mean().item() / nf:.0f} cycles (waiting for data {d[:, 15].mean().item() / nf:.0f}); "
      f"per PLAIN layer (67): mma {d[:, 14].mean().item() / 67:.0f} (waiting {(d[:, 7] - d[:, 15]).mean().item() / 67:.0f})")
