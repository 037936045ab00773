"""Developer script: does the forward of an image depend on the batch it travels in?  One 256-image launch (29 tiles per
pass and CTA in the body kernel) against the same images in chunks of 64 / 10 / 1, and 100 repetitions of the big launch.
Expected: run-to-run bit-identical for a given batch (the cross-CTA SE sums are integer atomics), NOT identical across batch
sizes - the fp32 partial sums a warp keeps over its run of tiles are grouped by tile ownership, which follows the batch
geometry; the SE scale then differs in its last bits and a few bf16 roundings flip downstream (as with cuDNN, whose
algorithm choice depends on the batch).     python tools/batch_invariance.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, fsr_b200
from oracle import weights
dev = torch.device("cuda:0")
cfg = dict(num_groups=2, blocks_per_group=10)
sd = weights.make_state_dict(0, "T1", **cfg)
g = torch.Generator().manual_seed(7)
sd["conv_last.weight"] = torch.randn(sd["conv_last.weight"].shape, generator=g) * 1e-2   # expose the body
m = fsr_b200.FaceEnhanceNet(**cfg); m.load_state_dict(sd); m = m.to(dev).eval()
x = torch.rand(256, 3, 64, 64, device=dev)
with torch.no_grad():
    big = m(x).clone()
    for chunk in (64, 10):
        parts = torch.cat([m(x[i:i + chunk]).clone() for i in range(0, 256, chunk)])
        print(f"256 at once vs chunks of {chunk}: {'bit-identical' if torch.equal(big, parts) else 'DIFFERENT, max |d| %.3e' % float((big - parts).abs().max())}")
    one = torch.cat([m(x[i:i + 1]).clone() for i in range(0, 16)])
    print(f"first 16 images one by one: {'bit-identical' if torch.equal(big[:16], one) else 'DIFFERENT, max |d| %.3e' % float((big[:16] - one).abs().max())}")
    rep = sum(int(not torch.equal(m(x), big)) for _ in range(100))
    print(f"100 repetitions of the 256-image launch: {rep} differed")
