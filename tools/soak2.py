"""Developer script: like soak.py, but on a mismatch it says WHERE - which stage of the forward first differs from the
first run (conv_first / residual groups / body / upsample stages / output) and which images / rows / columns / channels.
    python tools/soak2.py [iterations] [batch]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, fsr_b200
from oracle import weights
dev = torch.device("cuda:0")
cfg = dict(num_groups=6, blocks_per_group=10)
sd = weights.make_state_dict(0, "T1", **cfg)
g = torch.Generator().manual_seed(7)
sd["conv_last.weight"] = torch.randn(sd["conv_last.weight"].shape, generator=g) * 1e-2   # expose the body
m = fsr_b200.FaceEnhanceNet(**cfg); m.load_state_dict(sd); m = m.to(dev).eval()
iters = int(sys.argv[1]) if len(sys.argv) > 1 else 300
B = int(sys.argv[2]) if len(sys.argv) > 2 else 64
x = torch.rand(B, 3, 64, 64, device=dev)
TAPS = [("conv_first", 0, 0)] + [(f"group{k}", 4, k) for k in range(6)] + [("body", 1, 0), ("up0", 2, 0), ("up1", 3, 0)]
def snapshot():
    y, se = m._run(x, want_se=True)
    return [m.feature_tap(x.shape, w, i).clone() for _, w, i in TAPS] + [y.clone()], se.clone()
bad = 0
with torch.no_grad():
    ref, ref_se = snapshot()
    for it in range(iters):
        cur, cur_se = snapshot()
        if not torch.equal(ref_se, cur_se):
            dd = (ref_se != cur_se)
            r_first = int(dd.any(dim=2).any(dim=0).nonzero()[0])
            imgs = dd[:, r_first].any(dim=1).nonzero().flatten().tolist()
            if os.environ.get("SOAK_SLOTS") and bad < 6:
                for im in imgs[:2]:
                    sl = dd[im, r_first].nonzero().flatten().tolist()
                    print(f"   image {im} RCAB {r_first}: differing slots {sl}; ref {[round(float(ref_se[im, r_first, k]), 3) for k in sl]} cur {[round(float(cur_se[im, r_first, k]), 3) for k in sl]}; "
                          f"nonzero slots {(ref_se[im, r_first] != 0).nonzero().flatten().tolist()}", flush=True)
            if bad < 6: print(f"iteration {it}: SE vectors differ first at RCAB {r_first} (group {r_first // 10}, block {r_first % 10}) for images {imgs}: "
                  f"{int(dd[:, r_first].sum())} channels, max |d| {float((ref_se[:, r_first] - cur_se[:, r_first]).abs().max()):.3e}", flush=True)
        for (name, _, _), a, b in zip(TAPS + [("output", 0, 0)], ref, cur):
            if not torch.equal(a, b):
                d = (a.float() - b.float()).abs()
                nz = d.nonzero()
                dims = [f"{int(nz[:, k].min())}..{int(nz[:, k].max())}" for k in range(nz.shape[1])]
                if bad < 6: print(f"iteration {it}: FIRST difference at {name}: {nz.shape[0]} elements, max |d| {float(d.max()):.3e}, index ranges {dims} of shape {tuple(a.shape)}", flush=True)
                bad += 1
                break
print(f"B={B}: {iters} forwards, {bad} differed from the first")
