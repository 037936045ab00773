"""Developer script: a few Stage-1 steps of a small model (for ncu launch lists of the training step).
    python tools/step_once.py [groups] [blocks] [batch] [steps]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, fsr_b200
from oracle import weights
a = [int(v) for v in sys.argv[1:]] + [1, 2, 32, 2][len(sys.argv) - 1:]
cfg = dict(num_groups=a[0], blocks_per_group=a[1])
dev = torch.device("cuda:0")
m = fsr_b200.FaceEnhanceNet(**cfg); m.load_state_dict(weights.make_state_dict(0, "T1", **cfg)); m = m.to(dev).train()
st = fsr_b200.Stage1Step(m)
hr = torch.rand(a[2], 3, 256, 256, device=dev)
for _ in range(a[3]):
    loss, norm = st.step(hr)
torch.cuda.synchronize()
print("loss", loss.item(), "grad norm", norm.item())
