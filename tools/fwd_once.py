import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, fsr_b200
from oracle import weights
dev = torch.device("cuda:0")
cfg = dict(num_groups=int(sys.argv[1]), blocks_per_group=int(sys.argv[2])); B = int(sys.argv[3]); reps = int(sys.argv[4])
m = fsr_b200.FaceEnhanceNet(**cfg); m.load_state_dict(weights.make_state_dict(0, "T1", **cfg)); m = m.to(dev).eval()
x = torch.rand(B, 3, 64, 64, device=dev)
with torch.no_grad():
    for i in range(reps):
        y = m(x)
    torch.cuda.synchronize()
print("ok", cfg, B, reps, float(y.mean()))
