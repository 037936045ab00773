"""Developer script: times the batch-64 forward and its body kernel for the library selected by FEN_B200_LIB (a
variants/libfen_b200_<name>.so built by `python face-super-resolution_b200/build.py variant <name> -D...`), and checks
the output against the fp32 oracle on 2 images.   FEN_B200_LIB=... python tools/variant_time.py [label] [batch]"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, fsr_b200
from fsr_b200 import _lib
from oracle import fen_oracle, weights
label = sys.argv[1] if len(sys.argv) > 1 else os.path.basename(_lib.LIB_PATH)
B = int(sys.argv[2]) if len(sys.argv) > 2 else 64
lib = _lib.load(); dev = torch.device("cuda:0")
cfg = dict(num_groups=6, blocks_per_group=10)
sd = weights.make_state_dict(0, "T1", **cfg)
m = fsr_b200.FaceEnhanceNet(**cfg); m.load_state_dict(sd); m = m.to(dev).eval()
pool = [torch.rand(B, 3, 64, 64, device=dev, generator=torch.Generator(device=dev).manual_seed(i)) for i in range(48)]
with torch.no_grad():
    for i in range(5): y = m(pool[i])
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(40): m(pool[i % 48])
    e1.record(); torch.cuda.synchronize()
    fwd_ms = e0.elapsed_time(e1) / 40
    lib.fen_profile_body(1); body = []
    for i in range(30):
        m(pool[i % 48]); body.append(lib.fen_last_body_ms())
    lib.fen_profile_body(0)
    x = pool[0][:2]
    y = m(x).cpu()
ref = fen_oracle.fen_forward(sd, x.cpu())
body_ms = sum(body) / len(body)
flop = 127 * B * 2.0 * 4096 * 64 * 64 * 9
print(json.dumps({"variant": label, "batch": B, "forward_ms": round(fwd_ms, 4), "img_per_s": round(B / fwd_ms * 1e3),
                  "body_ms": round(body_ms, 4), "body_tflops": round(flop / body_ms / 1e9, 1),
                  "psnr_db": round(fen_oracle.psnr(y, ref), 2), "max_abs": float((y - ref).abs().max())}), flush=True)
