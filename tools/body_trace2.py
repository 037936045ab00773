"""Developer script: per-tile timeline of layers 20..23 of CTA 70 of the body kernel (-DFEN_BODY_DEBUG=2 build)."""
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
    dbg = torch.zeros(8192 + 4 * 512, dtype=torch.int64, device=dev)
    lib.fen_debug_set_counters(ctypes.c_void_p(dbg.data_ptr()))
    m(x); torch.cuda.synchronize()
    lib.fen_debug_set_counters(None)
d = dbg.cpu()
lay = d[4096:4096 + 128 * 16].view(128, 16)
t2 = d[8192:].view(4, 16, 32)
names = ["mma tile start", "mma issued", "epi acc seen", "epi tile done", "tma box issued", "mma data ok", "mma acc ok", "", "", "released", "acc committed", "nm top", "nm prefence", "nm postfence", "nm after issue", ""]
for li in range(4):
    L = 20 + li
    base = lay[L, 0].item()
    print(f"--- layer {L}: flags ok = 0; flag out {lay[L,7].item()-base}, next layer flags ok {lay[L+1,0].item()-base}")
    for e in (4, 0, 6, 5, 1, 9, 10, 11, 12, 13, 14, 2, 3):
        row = [(t2[li, e, i].item() - base) if t2[li, e, i].item() > 0 else None for i in range(20)]
        print(f"  {names[e]:15s} " + " ".join(f"{v:6d}" if v is not None else "     -" for v in row))
