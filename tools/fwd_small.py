"""Developer script: one small forward (for compute-sanitizer runs)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, fsr_b200
from oracle import weights, fen_oracle
B = int(sys.argv[1]) if len(sys.argv) > 1 else 2
cfg = dict(num_groups=1, blocks_per_group=2)
sd = weights.make_state_dict(0, "T1", **cfg)
m = fsr_b200.FaceEnhanceNet(**cfg); m.load_state_dict(sd); m = m.to("cuda").eval()
x = torch.rand(B, 3, 64, 64, generator=torch.Generator().manual_seed(1))
with torch.no_grad():
    y = m(x.cuda()).cpu()
torch.cuda.synchronize()
ref = fen_oracle.fen_forward(sd, x).clamp(0, 1)
print("B", B, "psnr", fen_oracle.psnr(y, ref), "maxabs", (y - ref).abs().max().item())
