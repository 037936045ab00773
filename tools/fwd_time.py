"""Developer script: forward time vs batch size."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import fsr_b200
from oracle import weights
dev = torch.device("cuda:0")
cfg = dict(num_groups=6, blocks_per_group=10)
m = fsr_b200.FaceEnhanceNet(**cfg); m.load_state_dict(weights.make_state_dict(0, "T1", **cfg)); m = m.to(dev).eval()
for B in [int(a) for a in sys.argv[1:]] or [64]:
    x = torch.rand(B, 3, 64, 64, device=dev)
    with torch.no_grad():
        for _ in range(3): m(x)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10): m(x)
        e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    print(f"B={B:4d}: {ms:7.3f} ms/forward  {B / ms * 1e3:8.0f} img/s  ({ms / B * 1e3:6.1f} us/img)")
