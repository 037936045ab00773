"""Developer script: repeat the full forward many times and check that every run reproduces the first one BIT FOR BIT
(the only cross-CTA reduction of the forward, the SE pool, is accumulated in 64-bit fixed point: integer atomics are
associative).  A race (a tile computed from stale data) or a reappearing order dependence shows up as any difference."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, fsr_b200
from oracle import weights
dev = torch.device("cuda:0")
cfg = dict(num_groups=6, blocks_per_group=10)
sd = weights.make_state_dict(0, "T1", **cfg)
g = torch.Generator().manual_seed(7)
sd["conv_last.weight"] = torch.randn(sd["conv_last.weight"].shape, generator=g) * 1e-2   # expose the body
m = fsr_b200.FaceEnhanceNet(**cfg); m.load_state_dict(sd); m = m.to(dev).train()
worst = 0.0
for B in [int(a) for a in sys.argv[2:]] or [64]:
    x = torch.rand(B, 3, 64, 64, device=dev)
    with torch.no_grad():
        ref = m(x).clone()
        scale = (ref - torch.nn.functional.interpolate(x, scale_factor=4, mode="bicubic", align_corners=False)).abs().max().item()
        for it in range(int(sys.argv[1]) if len(sys.argv) > 1 else 50):
            y = m(x)
            d = (y - ref).abs().max().item() / scale      # relative to the body's contribution
            worst = max(worst, d)
            if not torch.equal(y, ref) or torch.isnan(y).any():
                print(f"B={B} iteration {it}: max deviation {d:.3e} - NOT bit-identical to the first run"); sys.exit(1)
    print(f"B={B}: {sys.argv[1] if len(sys.argv) > 1 else 50} forwards, max deviation from the first {worst:.3e}")
