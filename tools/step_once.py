"""Developer script: a few Stage-1 training steps of the full model at batch 32 (for `ncu` launch lists) and the time of
the fused body kernel in the training forward vs the inference forward at the same batch."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, fsr_b200
from fsr_b200 import _lib
from oracle import weights
dev = torch.device("cuda:0"); lib = _lib.load()
B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
cfg = dict(num_groups=6, blocks_per_group=10)
m = fsr_b200.FaceEnhanceNet(**cfg); m.load_state_dict(weights.make_state_dict(0, "T1", **cfg)); m = m.to(dev).train()
st = fsr_b200.Stage1Step(m)
hr = torch.rand(B, 3, 256, 256, device=dev)
for _ in range(steps):
    st.step(hr)
torch.cuda.synchronize()
if os.environ.get("FEN_BODY_MS"):
    lib.fen_profile_body(1)
    st.step(hr); t_train = lib.fen_last_body_ms()
    m.eval()
    x = torch.rand(B, 3, 64, 64, device=dev)
    with torch.no_grad():
        m(x); m(x); t_inf = lib.fen_last_body_ms()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(); e0.record()
        for _ in range(10): m(x)
        e1.record(); torch.cuda.synchronize()
    lib.fen_profile_body(0)
    print(f"batch {B}: body kernel in the training forward {t_train:.3f} ms, in the inference forward {t_inf:.3f} ms; "
          f"whole inference forward {e0.elapsed_time(e1) / 10:.3f} ms")
