"""HBM roofline of the integer LR generator: algorithmic bytes = 196 608 read + 12 288 written per 256x256x3
image (uint8 out), + 49 152 when it also emits the fp32 NCHW model input.  The C ABI is called directly on preallocated
outputs (the timed region holds the kernel launches only).  python tools/lr_bandwidth.py [batch ...]"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, fsr_b200
from fsr_b200 import _lib
lib = _lib.load()
dev = torch.device("cuda:0")
batches = [int(a) for a in sys.argv[1:]] or [1024, 4096]
peak = 6544.7
try:
    peak = float(json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:
    pass
res = {}
for B in batches:
    npool = max(2, (3 * 1024 + B - 1) // B)                     # >= 600 MB of inputs: far beyond the 126 MB L2
    pool = [torch.randint(0, 256, (B, 256, 256, 3), dtype=torch.uint8, device=dev) for _ in range(npool)]
    u8 = torch.empty((B, 64, 64, 3), dtype=torch.uint8, device=dev)
    f32 = torch.empty((B, 3, 64, 64), dtype=torch.float32, device=dev)
    st = torch.cuda.current_stream().cuda_stream
    for name, pu, pf, bytes_img in (("u8_only", u8.data_ptr(), None, 196608 + 12288),
                                    ("u8_and_f32", u8.data_ptr(), f32.data_ptr(), 196608 + 12288 + 49152),
                                    ("f32_only", None, f32.data_ptr(), 196608 + 49152)):
        def once(i):
            _lib.check(lib.fen_lr_from_hr_u8(pool[i % npool].data_ptr(), pu, pf, B, 256, 256, 3, st), "fen_lr_from_hr_u8")
        for i in range(3): once(i)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n = 30
        e0.record()
        for i in range(n): once(i)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / n
        gbs = B * bytes_img / ms / 1e6
        res[f"{name}_b{B}"] = {"batch": B, "us_per_launch": round(ms * 1e3, 1), "img_per_s": round(B / ms * 1e3),
                               "achieved_gbs": round(gbs, 1), "peak_gbs": peak, "frac": round(gbs / peak, 3),
                               "algorithmic_bytes_per_image": bytes_img}
    del pool
print(json.dumps(res, indent=1))
