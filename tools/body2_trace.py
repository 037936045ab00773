"""Developer script: pass-level timeline of CTA 70 of the second-generation body kernel (-DFEN_B2_TRACE=1 build)."""
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
    dbg = torch.zeros(6144 + 512 * 8, dtype=torch.int64, device=dev)
    lib.fen_debug_set_counters(ctypes.c_void_p(dbg.data_ptr()))
    m(x); torch.cuda.synchronize()
    lib.fen_debug_set_counters(None)
t = dbg.cpu()[:4096].view(512, 8); t2 = dbg.cpu()[4096:6144].view(4, 16, 32); ts = dbg.cpu()[6144:].view(512, 8)
names = ["tma start", "A 1st data", "B 1st data", "A last issued", "B last issued", "epi 1st acc", "-", "flag out"]
print("pass  " + "  ".join(f"{n:>13s}" for n in names) + "   (cycles relative to the pass's tma start)   next pass tma start")
for P in range(80, 104):
    base = t[P, 0].item()
    row = "  ".join(f"{t[P, e].item() - base:13d}" if t[P, e].item() > 0 else f"{'-':>13s}" for e in range(8))
    print(f"{P:4d}  {row}   {t[P + 1, 0].item() - base}")
print("SE warp (relative to the same pass's tma start): flags poll start | flags ok + sums loaded | S ready | MMA result seen | scale ready")
for P in range(80, 104):
    if ts[P, 0].item() == 0: continue
    base = t[P, 0].item()
    print(f"{P:4d}  " + "  ".join(f"{ts[P, e].item() - base:9d}" for e in range(5)))
print("pass end (relative to the pass's tma start): warp 0 tiles stored | warp 7 fenced + arrived | all 8 arrived | flag out")
for P in range(80, 104):
    base = t[P, 0].item()
    print(f"{P:4d}  " + "  ".join(f"{ts[P, e].item() - base:9d}" for e in (5, 7, 6)) + f"  {t[P, 7].item() - base:9d}")
tot = t[253, 7].item() - t[0, 0].item()
print("total cycles pass 0 -> flag of pass 253:", tot)

n2 = ["epi wait acc", "epi acc seen", "epi tmem read", "epi stored", "mma tile start", "mma acc ok", "mma data ok", "mma issued", "mma committed", "tma box wait", "tma box issued"]
for li in range(4):
    P = 80 + li; base = t[P, 0].item()
    print(f"--- pass {P} (tma start = 0)")
    for e in (9, 10, 4, 5, 6, 7, 8, 0, 1, 2, 3):
        row = [(t2[li, e, i].item() - base) if t2[li, e, i].item() > 0 else None for i in range(12)]
        print(f"  {n2[e]:15s} " + " ".join(f"{v:6d}" if v is not None else "     -" for v in row))
