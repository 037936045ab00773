"""HBM roofline of the integer LR generator: algorithmic bytes = 196 608 read + 12 288 written per 256x256x3
image (uint8 out), + 49 152 when it also emits the fp32 NCHW model input.  python tools/lr_bandwidth.py [batch]"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, fsr_b200
dev = torch.device("cuda:0")
B = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
peak = 6544.7
try:
    peak = float(json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:
    pass
pool = [torch.randint(0, 256, (B, 256, 256, 3), dtype=torch.uint8, device=dev) for _ in range(3)]   # 3 x 201 MB > L2
res = {}
for name, kw, bytes_img in (("u8_only", dict(want_u8=True, want_f32=False), 196608 + 12288),
                            ("u8_and_f32", dict(want_u8=True, want_f32=True), 196608 + 12288 + 49152),
                            ("f32_only", dict(want_u8=False, want_f32=True), 196608 + 49152)):
    for i in range(3): fsr_b200.lr_from_hr(pool[i % 3], **kw)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = 30
    e0.record()
    for i in range(n): fsr_b200.lr_from_hr(pool[i % 3], **kw)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    gbs = B * bytes_img / ms / 1e6
    res[name] = {"batch": B, "us_per_launch": round(ms * 1e3, 1), "img_per_s": round(B / ms * 1e3), "achieved_gbs": round(gbs, 1),
                 "peak_gbs": peak, "frac": round(gbs / peak, 3), "algorithmic_bytes_per_image": bytes_img}
print(json.dumps(res, indent=1))
