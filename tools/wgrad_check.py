"""Developer script: the weight-gradient kernel generations against each other (FEN_WGRAD=1 mma.sync is the reference;
the env variable is read on every call) through one backward pass, plus their time in a training step."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, fsr_b200
from oracle import weights
dev = torch.device("cuda:0")
cfg = dict(num_groups=1, blocks_per_group=2)
m = fsr_b200.FaceEnhanceNet(**cfg); m.load_state_dict(weights.make_state_dict(0, "T1", **cfg)); m = m.to(dev).train()
for shape in [(3, 3, 64, 64), (1, 3, 64, 128)]:
    x = torch.rand(*shape, device=dev)
    dout = torch.rand(shape[0], 3, 4 * shape[2], 4 * shape[3], device=dev) / 1e5
    res = {}
    for v in sys.argv[1:] or ["1", "2"]:
        os.environ["FEN_WGRAD"] = v
        m.zero_grad()
        m(x).backward(dout)
        torch.cuda.synchronize()
        res[v] = {k: p.grad.clone() for k, p in m.named_parameters()}
    ref = res["1"]
    for v, g in res.items():
        if v == "1":
            continue
        worst = max(((g[k] - ref[k]).norm() / ref[k].norm().clamp_min(1e-30)).item() for k in ref)
        bad = [(k, ((g[k] - ref[k]).norm() / ref[k].norm().clamp_min(1e-30)).item()) for k in ref
               if (g[k] - ref[k]).norm() > 1e-3 * ref[k].norm()]
        print(f"shape {shape} FEN_WGRAD={v}: worst relative difference to mma.sync {worst:.2e}; tensors above 1e-3: {bad[:6]}")
