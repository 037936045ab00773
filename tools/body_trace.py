"""Developer script: timeline of CTA 70 of the persistent body kernel (needs a -DFEN_BODY_DEBUG=2 build)."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, fsr_b200
from fsr_b200 import _lib
from oracle import weights
lib = _lib.load(); dev = torch.device("cuda:0")
B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
cfg = dict(num_groups=6, blocks_per_group=10)
m = fsr_b200.FaceEnhanceNet(**cfg); m.load_state_dict(weights.make_state_dict(0, "T1", **cfg)); m = m.to(dev).eval()
x = torch.rand(B, 3, 64, 64, device=dev)
with torch.no_grad():
    for _ in range(2): m(x)
    dbg = torch.zeros(4096 + 128 * 16, dtype=torch.int64, device=dev)
    lib.fen_debug_set_counters(ctypes.c_void_p(dbg.data_ptr()))
    m(x); torch.cuda.synchronize()
    lib.fen_debug_set_counters(None)
t = dbg.cpu()[4096:].view(128, 16)[:127, :8].double()
names = ["flags ok", "1st data", "2nd-last MMA", "last MMA", "SE ready", "1st acc", "last store", "flag out"]
print("layer kind  " + "  ".join(f"{n:>12s}" for n in names) + "   (cycles relative to this layer's 'flags ok'; layer length = next flags ok)")
for L in range(20, 32):
    kind = "after" if L == 126 else ("group" if L % 21 == 20 else ("conv1" if (L % 21) % 2 == 0 else "conv2"))
    base = t[L, 0]
    row = "  ".join(f"{(t[L, e] - base).item():12.0f}" if t[L, e] > 0 else f"{'-':>12s}" for e in range(8))
    print(f"{L:4d} {kind:6s} {row}   next layer starts at {(t[L + 1, 0] - base).item():.0f}")
    if kind == "conv2":
        tt = dbg.cpu()[4096:].view(128, 16)[L].double()
        print("        SE phases (rel. flags ok): enter %d | P1 done %d | bar1 %d | P2+bar2 %d | P3+bar3 %d | P4+bar4 %d | P5+bar5 %d" %
              tuple((tt[e] - base).item() for e in range(8, 15)))
