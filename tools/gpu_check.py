"""Developer smoke script (not a test): runs the CUDA path against the CPU oracle on a B200 and
prints per-stage errors.  Usage on the GPU box:  python tools/gpu_check.py [stage ...]"""
import ctypes as C
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.nn.functional as F

import fsr_b200
from fsr_b200 import _lib
from oracle import fen_oracle, lr_oracle, weights

dev = torch.device("cuda:0")
lib = _lib.load()
stages = sys.argv[1:] or ["lr", "conv", "small", "full"]


def nhwc_bf16(t):  # NCHW fp32 -> NHWC bf16 cuda
    return t.permute(0, 2, 3, 1).contiguous().to(torch.bfloat16).to(dev)


def conv_case(B, H, W, epi, seed=0):
    g = torch.Generator().manual_seed(seed)
    x = (torch.randn(B, 64, H, W, generator=g) * 0.5).to(torch.bfloat16).float()
    w = (torch.randn(64, 64, 3, 3, generator=g) * 0.06).to(torch.bfloat16).float()
    bias = torch.randn(64, generator=g) * 0.1
    slope = torch.rand(64, generator=g) * 0.4 + 0.05
    res = (torch.randn(B, 64, H, W, generator=g) * 0.5).to(torch.bfloat16).float()
    ref = F.conv2d(x, w, bias, padding=1)
    sums_ref = ref.sum(dim=(2, 3))
    if epi == 0:
        ref = F.prelu(ref, slope)
    elif epi == 2:
        ref = ref + res
    xd, rd = nhwc_bf16(x), nhwc_bf16(res)
    wp = torch.empty(9 * 64 * 64, dtype=torch.bfloat16, device=dev)
    wd = w.to(dev).contiguous()
    st = torch.cuda.current_stream().cuda_stream
    _lib.check(lib.fen_pack_conv3x3(wd.data_ptr(), 64, 64, wp.data_ptr(), st), "pack")
    out = torch.full((B, H, W, 64), float("nan"), dtype=torch.bfloat16, device=dev)
    sums = torch.zeros(B, 64, dtype=torch.float32, device=dev)
    bd, sd = bias.to(dev), slope.to(dev)
    rc = lib.fen_conv3x3_c64(xd.data_ptr(), wp.data_ptr(), bd.data_ptr(), sd.data_ptr(), rd.data_ptr(),
                             sums.data_ptr(), out.data_ptr(), B, H, W, epi, st)
    _lib.check(rc, "conv")
    torch.cuda.synchronize()
    got = out.float().cpu().permute(0, 3, 1, 2)
    err = (got - ref).abs()
    print(f"conv epi={epi} B={B} {H}x{W}: max|err|={err.max().item():.4e} (ref max {ref.abs().max().item():.3f}) "
          f"nan={torch.isnan(got).sum().item()}", flush=True)
    if epi == 1:
        e2 = (sums.cpu() - sums_ref).abs().max().item()
        print(f"   channel sums: max|err|={e2:.4e} (ref max {sums_ref.abs().max().item():.2f})", flush=True)
    if err.max().item() > 0.1 or torch.isnan(got).any():
        bad = (err > 0.1) | torch.isnan(got)
        idx = bad.nonzero()
        print("   first bad idx (n,c,y,x):", idx[:5].tolist(), " count", bad.sum().item())
        per_row = bad.any(dim=1).any(dim=2)[0].nonzero().flatten().tolist()
        print("   bad rows of image 0:", per_row[:40])


if "lr" in stages:
    rng = np.random.default_rng(0)
    hr = rng.integers(0, 256, (5, 256, 256, 3), dtype=np.uint8)
    u8, f32 = fsr_b200.lr_from_hr(torch.from_numpy(hr).to(dev))
    ref = lr_oracle.lr_from_hr_u8(hr)
    print("lr u8 mismatches:", int((u8.cpu().numpy() != ref).sum()),
          " f32 mismatches:", int((f32.cpu().numpy() != lr_oracle.to_tensor_chw(ref)).sum()), flush=True)

if "conv" in stages:
    conv_case(1, 64, 64, 5)
    conv_case(2, 64, 64, 0)
    conv_case(3, 64, 64, 1)
    conv_case(2, 64, 64, 2)
    conv_case(1, 128, 128, 5)
    conv_case(64, 64, 64, 1)


def net_case(num_groups, blocks, B, tier="T1", seed=0):
    cfg = dict(num_groups=num_groups, blocks_per_group=blocks)
    sd = weights.make_state_dict(seed, tier, **cfg)
    m = fsr_b200.FaceEnhanceNet(**cfg)
    m.load_state_dict(sd, strict=True)
    m = m.to(dev).eval()
    x = torch.rand(B, 3, 64, 64, generator=torch.Generator().manual_seed(seed + 1))
    taps = {}
    nb = min(B, 2)
    t0 = time.time()
    ref = fen_oracle.fen_forward(sd, x[:nb], taps=taps)
    t_ref = time.time() - t0
    with torch.no_grad():
        out = m(x.to(dev))
        torch.cuda.synchronize()
        am = m.get_attention_maps(x.to(dev))
    got = out[:nb].cpu()
    print(f"net G={num_groups} Bk={blocks} B={B} {tier}: PSNR={fen_oracle.psnr(got, ref):.2f} dB "
          f"max|err|={(got - ref).abs().max().item():.4e} nan={torch.isnan(out).sum().item()} "
          f"(oracle {t_ref:.1f}s for {nb} img)", flush=True)
    def rel(a, b):
        return ((a - b).norm() / b.norm()).item()
    f0 = m.feature_tap(x.shape, 0)[:nb].float().cpu().permute(0, 3, 1, 2)
    print(f"   conv_first rel-L2 {rel(f0, taps['conv_first']):.3e}")
    for g in range(num_groups):
        t = m.feature_tap(x.shape, 4, g)[:nb].float().cpu().permute(0, 3, 1, 2)
        print(f"   group{g} rel-L2 {rel(t, taps[f'group{g}']):.3e}")
    for which, name in ((1, "body"), (2, "up0"), (3, "up1")):
        t = m.feature_tap(x.shape, which)[:nb].float().cpu().permute(0, 3, 1, 2)
        print(f"   {name} rel-L2 {rel(t, taps[name]):.3e}")
    se = torch.stack([am[f"group{g}_rcab{b}"] for g in range(num_groups) for b in range(blocks)], 1)[:nb].cpu()
    print(f"   SE scales max|err| {(se - taps['se']).abs().max().item():.3e}", flush=True)


if "small" in stages:
    net_case(1, 2, 2)
    net_case(2, 2, 3, tier="T0")
if "full" in stages:
    net_case(6, 10, 64)
if "time" in stages:
    cfg = dict(num_groups=6, blocks_per_group=10)
    m = fsr_b200.FaceEnhanceNet(**cfg)
    m.load_state_dict(weights.make_state_dict(0, "T1", **cfg))
    m = m.to(dev).eval()
    x = torch.rand(64, 3, 64, 64, device=dev)
    with torch.no_grad():
        for _ in range(3):
            m(x)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            m(x)
        e1.record()
        torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    print(f"forward B=64: {ms:.3f} ms -> {64 / ms * 1e3:.0f} img/s, launches {lib.fen_last_launch_count()}")
