"""Parity tests proper: the CUDA path, called through the C ABI (ctypes), against the CPU oracle, the
committed golden vectors (made from the real reference / cv2) and size-independent properties.
Run on a B200 with `pytest -m gpu`.

Tolerances: LR generator bit-exact.  SR output: PSNR >= 50 dB and max-abs <= 2e-2 on [0,1] against the
fp32 reference output (BASELINE.json north_star), weight tier T1."""
import ctypes as C
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

import cases
import fsr_b200
from fsr_b200 import _lib
from oracle import fen_oracle, lr_oracle, weights

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))
LR_GOLD = np.load(os.path.join(HERE, "golden", "lr_golden.npz"))
FEN_GOLD = np.load(os.path.join(HERE, "golden", "fen_golden.npz"))
PSNR_BAR, MAXABS_BAR = 50.0, 2e-2


@pytest.fixture(scope="module")
def dev(built_lib):
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    return torch.device("cuda:0")


# ------------------------------------------------------------------ LR generator (bit-exact)
@pytest.mark.parametrize("name", [c[0] for c in cases.LR_CASES])
def test_lr_kernel_matches_cv2_golden(name, dev):
    hr = cases.lr_input(name)
    u8, f32 = fsr_b200.lr_from_hr(torch.from_numpy(hr).to(dev).unsqueeze(0))
    assert np.array_equal(u8[0].cpu().numpy(), LR_GOLD[name])
    assert np.array_equal(f32[0].cpu().numpy(), lr_oracle.to_tensor_chw(LR_GOLD[name]))


def test_lr_kernel_large_batch_against_oracle(dev):
    rng = np.random.default_rng(11)
    hr = rng.integers(0, 256, (96, 256, 256, 3), dtype=np.uint8)
    hr[7] = (rng.integers(0, 8, (256, 256, 3)) * 32).astype(np.uint8)   # tie-heavy image
    hr[8] = (rng.integers(0, 2, (256, 256, 3)) * 255).astype(np.uint8)  # saturating image
    u8, f32 = fsr_b200.lr_from_hr(torch.from_numpy(hr).to(dev))
    ref = lr_oracle.lr_from_hr_u8(hr)
    assert np.array_equal(u8.cpu().numpy(), ref)
    assert np.array_equal(f32.cpu().numpy(), lr_oracle.to_tensor_chw(ref))


def test_lr_kernel_properties_and_edges(dev):
    # constant 4x4 blocks are fixed points; empty batch; single output; only-one-output requests
    blocks = torch.from_numpy(cases.lr_input("blocks_256")).to(dev).unsqueeze(0)
    u8, _ = fsr_b200.lr_from_hr(blocks, want_f32=False)
    assert torch.equal(u8[0], blocks[0, ::4, ::4])
    e_u8, e_f32 = fsr_b200.lr_from_hr(torch.zeros(0, 8, 8, 3, dtype=torch.uint8, device=dev))
    assert e_u8.shape == (0, 2, 2, 3) and e_f32.shape == (0, 3, 2, 2)
    one = torch.full((1, 4, 4, 1), 200, dtype=torch.uint8, device=dev)
    assert fsr_b200.lr_from_hr(one)[0].item() == 200
    with pytest.raises(ValueError):
        fsr_b200.lr_from_hr(torch.zeros(1, 6, 8, 3, dtype=torch.uint8, device=dev))
    with pytest.raises(ValueError):
        fsr_b200.lr_from_hr(torch.zeros(1, 8, 8, 5, dtype=torch.uint8, device=dev))


def test_create_lr_image_numpy_boundary(dev):
    hr = cases.lr_input("random_256")
    out = fsr_b200.create_lr_image(hr, lr_size=64, method="bicubic")
    assert out.dtype == np.uint8 and np.array_equal(out, LR_GOLD["random_256"])


def test_float_lr_generator_matches_trainer_call(dev):
    """F.interpolate(hr, 0.25, bicubic) in fp32 (tolerance 2.4e-7: summation order) and
    the scripts' uint8 variant (truncation: a 1-LSB difference only where v * 255 sits on an integer)."""
    g = torch.Generator().manual_seed(21)
    hr8 = torch.randint(0, 256, (5, 3, 256, 256), generator=g, dtype=torch.uint8)
    hr = hr8.float() / 255.0                                   # scripts: uint8 image / 255
    ref = lr_oracle.lr_from_hr_float(hr)
    f32, u8 = fsr_b200.lr_from_hr_float(hr.to(dev), want_f32=True, want_u8=True, bgr=True)
    assert f32.shape == (5, 3, 64, 64) and u8.shape == (5, 64, 64, 3)
    assert (f32.cpu() - ref).abs().max().item() <= 2.4e-7
    assert f32.min() < 0 and f32.max() > 1                     # neither rounded nor clamped
    q_ref = lr_oracle.quantize_u8_hwc(ref.numpy(), bgr=True)
    diff = np.abs(u8.cpu().numpy().astype(np.int32) - q_ref.astype(np.int32))
    assert diff.max() <= 1 and (diff != 0).mean() <= 1e-3
    # trainer input (continuous values), other shapes, error behaviour
    x = torch.rand(2, 1, 8, 12, generator=g)
    f32, _ = fsr_b200.lr_from_hr_float(x.to(dev))
    assert (f32.cpu() - lr_oracle.lr_from_hr_float(x)).abs().max().item() <= 2.4e-7
    with pytest.raises(ValueError):
        fsr_b200.lr_from_hr_float(torch.rand(1, 3, 6, 8, device=dev))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        fsr_b200.lr_from_hr_float(torch.rand(1, 3, 8, 8))


def test_sr_to_uint8_matches_scripts_to_numpy(dev):
    g = torch.Generator().manual_seed(22)
    sr = torch.rand(3, 3, 40, 56, generator=g) * 1.4 - 0.2     # values outside [0, 1] get clipped
    sr[0, 0, 0, :4] = torch.tensor([0.0, 1.0, 127.0 / 255.0, 254.999 / 255.0])
    for bgr in (False, True):
        got = fsr_b200.sr_to_uint8(sr.to(dev), bgr=bgr).cpu().numpy()
        assert np.array_equal(got, lr_oracle.quantize_u8_hwc(sr.numpy(), bgr=bgr))


# ------------------------------------------------------------------ Stage-1 step: loss / optimiser side
def test_l1_loss_and_gradient_match_torch(dev):
    """nn.L1Loss(mean) (combined.py:38-47) and what loss.backward() hands to the network output."""
    g = torch.Generator().manual_seed(31)
    sr = (torch.rand(2, 3, 256, 256, generator=g) * 1.3 - 0.15).requires_grad_(True)
    hr = torch.rand(2, 3, 256, 256, generator=g)
    hr.view(-1)[:100] = sr.detach().view(-1)[:100]                  # exact zeros: sign(0) = 0 as in torch
    ref = torch.nn.L1Loss(reduction="mean")(sr, hr)
    ref.backward()
    loss, dsr = fsr_b200.l1_loss(sr.detach().to(dev), hr.to(dev))
    assert abs(loss.item() - ref.item()) <= 1e-6 * abs(ref.item()) + 1e-9
    assert torch.equal(dsr.cpu(), sr.grad)
    loss2, none = fsr_b200.l1_loss(sr.detach().to(dev), hr.to(dev), want_grad=False)
    assert none is None and loss2.item() == loss.item()             # deterministic two-stage reduction
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        fsr_b200.l1_loss(sr.detach(), hr)


def test_psnr_matches_trainer_metric(dev):
    """Trainer._compute_psnr (trainer.py:621-628): 10 log10(1 / mse) over the whole batch."""
    g = torch.Generator().manual_seed(51)
    hr = torch.rand(3, 3, 128, 96, generator=g)
    sr = (hr + torch.randn(hr.shape, generator=g) * 0.03).clamp(0, 1)
    ref = 10 * torch.log10(1.0 / torch.mean((sr - hr) ** 2))
    got = fsr_b200.psnr(sr.to(dev), hr.to(dev))
    assert abs(got.item() - ref.item()) <= 1e-4
    assert abs(fsr_b200.psnr(sr.to(dev) * 255, hr.to(dev) * 255, data_range=255.0).item() - ref.item()) <= 1e-3
    assert fsr_b200.psnr(hr.to(dev), hr.to(dev)).item() == float("inf")


@pytest.mark.parametrize("max_norm,wd", [(0.5, 0.0), (0.5, 1e-2), (0.0, 0.0), (1e3, 0.0)])
def test_clip_adamw_matches_torch(max_norm, wd, dev):
    """clip_grad_norm_(0.5) + AdamW(lr 1e-4) of the trainer (trainer.py:217-221, 490-503), 5 steps on a flat
    vector of the model's size class, against torch.optim.AdamW on the CPU."""
    g = torch.Generator().manual_seed(41)
    n = 200_003
    p0 = torch.randn(n, generator=g) * 0.05
    ref_p = torch.nn.Parameter(p0.clone())
    opt_ref = torch.optim.AdamW([ref_p], lr=1e-4, weight_decay=wd, foreach=False)
    p_dev = p0.clone().to(dev)
    opt = fsr_b200.ClipAdamW(p_dev, lr=1e-4, weight_decay=wd, max_norm=max_norm)
    for step in range(5):
        grad = torch.randn(n, generator=g) * (10.0 if step % 2 == 0 else 1e-4)   # clipped and unclipped steps
        ref_p.grad = grad.clone()
        ref_norm = torch.nn.utils.clip_grad_norm_([ref_p], max_norm) if max_norm > 0 else grad.norm()
        opt_ref.step()
        norm = opt.step(grad.to(dev))
        assert abs(norm.item() - ref_norm.item()) <= 2e-6 * ref_norm.item()
        assert (p_dev.cpu() - ref_p.detach()).abs().max().item() <= 2e-7   # lr 1e-4: updates are ~1e-4
    m_ref = opt_ref.state[ref_p]["exp_avg"]
    assert (opt.exp_avg.cpu() - m_ref).abs().max().item() <= 5e-6 * m_ref.abs().max().item() + 1e-9
    rel = (opt.exp_avg_sq.cpu() - opt_ref.state[ref_p]["exp_avg_sq"]).abs().max() / opt_ref.state[ref_p]["exp_avg_sq"].abs().max()
    assert rel.item() <= 5e-5      # (clip coefficient from a norm that differs by ~2e-6 relative, squared, 5 steps)


# ------------------------------------------------------------------ single convolution (C ABI)
def _conv_call(dev, x, w, bias, slope, res, epi):
    lib = _lib.load()
    B, _, H, W = x.shape
    to_nhwc = lambda t: t.permute(0, 2, 3, 1).contiguous().to(torch.bfloat16).to(dev)
    xd, rd = to_nhwc(x), to_nhwc(res)
    wd, bd, sd = w.to(dev).contiguous(), bias.to(dev), slope.to(dev)
    wp = torch.empty(9 * 64 * 64, dtype=torch.bfloat16, device=dev)
    st = torch.cuda.current_stream().cuda_stream
    _lib.check(lib.fen_pack_conv3x3(wd.data_ptr(), 64, 64, wp.data_ptr(), st), "fen_pack_conv3x3")
    out = torch.full((B, H, W, 64), float("nan"), dtype=torch.bfloat16, device=dev)
    sums = torch.zeros(B, 64, dtype=torch.float32, device=dev)
    rc = lib.fen_conv3x3_c64(xd.data_ptr(), wp.data_ptr(), bd.data_ptr(), sd.data_ptr(), rd.data_ptr(),
                             sums.data_ptr(), out.data_ptr(), B, H, W, epi, st)
    _lib.check(rc, "fen_conv3x3_c64")
    torch.cuda.synchronize()
    return out.float().cpu().permute(0, 3, 1, 2), sums.cpu()


@pytest.mark.parametrize("B,H,W,epi", [(1, 64, 64, 5), (2, 64, 64, 0), (3, 64, 64, 1), (2, 64, 64, 2),
                                       (1, 128, 128, 0), (1, 64, 192, 2), (5, 128, 64, 1), (150, 64, 64, 1)])
def test_conv3x3_against_fp32_reference(B, H, W, epi, dev):
    g = torch.Generator().manual_seed(B * 1000 + H + epi)
    bf = lambda t: t.to(torch.bfloat16).float()      # operands are bf16 on the device
    x = bf(torch.randn(B, 64, H, W, generator=g) * 0.5)
    w = bf(torch.randn(64, 64, 3, 3, generator=g) * 0.06)
    bias = torch.randn(64, generator=g) * 0.1
    slope = torch.rand(64, generator=g) * 0.4 + 0.05
    res = bf(torch.randn(B, 64, H, W, generator=g) * 0.5)
    ref = F.conv2d(x, w, bias, padding=1)            # plain PyTorch fp32 reference of the same op
    sums_ref = ref.sum(dim=(2, 3))
    if epi == 0:
        ref = F.prelu(ref, slope)
    elif epi == 2:
        ref = ref + res
    got, sums = _conv_call(dev, x, w, bias, slope, res, epi)
    assert not torch.isnan(got).any()
    # fp32 accumulate, bf16 store: error <= half a bf16 ulp of the result (+ accumulation order)
    tol = ref.abs().clamp(min=1.0) * 2.0 ** -8 + 1e-3
    assert ((got - ref).abs() <= tol).all(), (got - ref).abs().max().item()
    if epi == 1:
        assert (sums - sums_ref).abs().max().item() <= 1e-3 * sums_ref.abs().max().item() + 1e-2


def test_conv3x3_zero_padding_and_linearity(dev):
    # all-ones input, centre-tap-only weights -> identity; all-ones weights -> border counts 4/6/9
    x = torch.ones(1, 64, 64, 64)
    z = torch.zeros(64)
    w_id = torch.zeros(64, 64, 3, 3)
    w_id[torch.arange(64), torch.arange(64), 1, 1] = 1.0
    got, _ = _conv_call(dev, x, w_id, z, z, x, 5)
    assert torch.equal(got, x)
    w_cnt = torch.zeros(64, 64, 3, 3)
    w_cnt[:, 0] = 1.0
    got, _ = _conv_call(dev, x, w_cnt, z, z, x, 5)
    assert got[0, 0, 0, 0] == 4 and got[0, 5, 0, 7] == 6 and got[0, 9, 31, 0] == 6 and got[0, 3, 20, 20] == 9
    assert got[0, 63, 63, 63] == 4 and got[0, 1, 63, 30] == 6


def test_conv3x3_rejects_bad_arguments(dev):
    lib = _lib.load()
    t = torch.zeros(64 * 64 * 64, dtype=torch.bfloat16, device=dev)
    st = torch.cuda.current_stream().cuda_stream
    rc = lib.fen_conv3x3_c64(t.data_ptr(), t.data_ptr(), None, None, None, None, t.data_ptr(), 1, 64, 0, 5, st)
    assert rc == _lib.FEN_EINVAL
    rc = lib.fen_conv3x3_c64(t.data_ptr(), t.data_ptr(), None, None, None, None, t.data_ptr(), 1, 64, 64, 2, st)
    assert rc == _lib.FEN_EINVAL and b"residual" in lib.fen_last_error()
    rc = lib.fen_conv3x3_c64(t.data_ptr(), t.data_ptr(), None, None, None, None, t.data_ptr(), 1, 64, 64, 3, st)
    assert rc == _lib.FEN_EINVAL


# ------------------------------------------------------------------ whole network
def _model(cfg, sd, dev, train=False):
    m = fsr_b200.FaceEnhanceNet(**cfg)
    m.load_state_dict(sd, strict=True)
    m = m.to(dev)
    return m.train() if train else m.eval()


@pytest.mark.parametrize("name", [c[0] for c in cases.FEN_CASES])
def test_forward_matches_reference_golden(name, dev):
    _, cfg, tier, seed, _ = [c for c in cases.FEN_CASES if c[0] == name][0]
    sd = weights.make_state_dict(seed, tier, **cfg)
    x = torch.from_numpy(cases.fen_input(name)).to(dev)
    g_train = torch.from_numpy(FEN_GOLD[name + "/train"])
    with torch.no_grad():
        y_eval = _model(cfg, sd, dev)(x).cpu()
        y_train = _model(cfg, sd, dev, train=True)(x).cpu()
    for got, ref in ((y_eval, g_train.clamp(0, 1)), (y_train, g_train)):
        assert got.shape == ref.shape and got.dtype == torch.float32
        assert fen_oracle.psnr(got, ref) >= PSNR_BAR
        assert (got - ref).abs().max().item() <= MAXABS_BAR
    assert y_eval.min() >= 0 and y_eval.max() <= 1
    assert y_train.min() < 0 and y_train.max() > 1        # train mode really is unclamped
    if tier == "T0":  # conv_last == 0: output is the bicubic skip alone, fp32 all the way
        assert (y_train - g_train).abs().max().item() <= 2e-6


def test_forward_batch64_full_model_against_oracle(dev):
    """BASELINE.json config 2: bf16 inference, batch 64, 6 x 10 x 64 model, tier T1 - ALL 64 images against the fp32
    oracle (a few seconds of CPU), plus the report-only stress tier T2 (conv_last sigma 1e-2)."""
    cfg = dict(num_groups=6, blocks_per_group=10)
    sd = weights.make_state_dict(0, "T1", **cfg)
    x = torch.rand(64, 3, 64, 64, generator=torch.Generator().manual_seed(5))
    m = _model(cfg, sd, dev)
    with torch.no_grad():
        y = m(x.to(dev)).cpu()
    taps = {}
    torch.set_num_threads(os.cpu_count() or 1)
    ref = fen_oracle.fen_forward(sd, x, taps=taps)
    per_img = [fen_oracle.psnr(y[i], ref[i]) for i in range(64)]
    psnr, max_abs = fen_oracle.psnr(y, ref), (y - ref).abs().max().item()
    print(f"\nbatch-64 parity (all 64 images): PSNR {psnr:.2f} dB (worst image {min(per_img):.2f} dB), max|err| {max_abs:.3e}")
    assert psnr >= PSNR_BAR and min(per_img) >= PSNR_BAR and max_abs <= MAXABS_BAR
    assert not torch.isnan(y).any()
    # intermediate feature maps: bf16 rounding grows slowly; > 5 % would be a bug
    body = m.feature_tap(x.shape, 1).float().cpu().permute(0, 3, 1, 2)
    rel = ((body - taps["body"]).norm() / taps["body"].norm()).item()
    assert rel < 0.05, rel
    # batch independence: image 31 alone gives the same answer up to the summation order of the SE
    # pool (different tile->CTA split), which bf16 re-rounding amplifies to the parity-noise level;
    # any cross-image leak would show up at the 0.3 level of the conv_last residual
    with torch.no_grad():
        y1 = m(x[31:32].to(dev)).cpu()
    assert (y1[0] - y[31]).abs().max().item() <= MAXABS_BAR
    assert fen_oracle.psnr(y1[0], y[31]) >= PSNR_BAR
    # tier T2 (stress, report only): the body's rounding noise amplified 10x by conv_last
    sd2 = weights.make_state_dict(0, "T2", **cfg)
    m2 = _model(cfg, sd2, dev)
    with torch.no_grad():
        y2 = m2(x[:8].to(dev)).cpu()
    ref2 = fen_oracle.fen_forward(sd2, x[:8])
    line = (f"T2 (conv_last sigma 1e-2, report only): PSNR {fen_oracle.psnr(y2, ref2):.2f} dB, "
            f"max|err| {(y2 - ref2).abs().max().item():.3e}; T1 batch 64: PSNR {psnr:.2f} dB "
            f"(worst image {min(per_img):.2f} dB), max|err| {max_abs:.3e}")
    print(line)
    out = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out")
    os.makedirs(out, exist_ok=True)
    with open(os.path.join(out, "parity_t1_t2.txt"), "w") as f:
        f.write(line + "\n")
    assert fen_oracle.psnr(y2, ref2) >= 30.0          # sanity only


@pytest.mark.parametrize("B", [1, 2, 5, 7, 9, 10, 12, 16, 24, 88, 140])
def test_body_kernel_batch_geometries(B, dev):
    """The persistent body kernel splits the batch into two interleaved image sets (even B) and gives
    every CTA a fixed run of tiles: 1 tile per CTA (B <= 8), odd / even tile counts, runs that straddle
    two images, runs of different length in the two sets (one issuer warp without tiles in one of them) ...
    Every geometry must agree with the oracle (bf16 noise
    only) on a model whose conv_last is LARGE enough to expose the body (sigma 3e-2, not the T1 1e-3)."""
    cfg = dict(num_groups=1, blocks_per_group=2)
    sd = weights.make_state_dict(3, "T1", **cfg)
    g = torch.Generator().manual_seed(77)
    sd["conv_last.weight"] = torch.randn(sd["conv_last.weight"].shape, generator=g) * 3e-2
    x = torch.rand(B, 3, 64, 64, generator=torch.Generator().manual_seed(B))
    m = _model(cfg, sd, dev, train=True)
    with torch.no_grad():
        y = m(x.to(dev)).cpu()
    idx = sorted(set([0, B // 2, B - 1]))
    taps = {}
    ref = fen_oracle.fen_forward(sd, x[idx], training=True, taps=taps)
    err = (y[idx] - ref).abs().max().item()
    scale = (ref - F.interpolate(x[idx], scale_factor=4, mode="bicubic", align_corners=False)).abs().max().item()
    assert not torch.isnan(y).any()
    assert scale > 0.05, scale                    # the body really contributes
    assert err <= 0.03 * scale + 1e-3, (err, scale)
    maps = m.eval().get_attention_maps(x.to(dev))
    got = torch.stack([maps["group0_rcab0"], maps["group0_rcab1"]], 1).cpu()[idx]
    assert (got - taps["se"]).abs().max().item() <= 1e-2


def test_attention_maps_match_reference(dev):
    name = "small_T1"
    _, cfg, tier, seed, _ = [c for c in cases.FEN_CASES if c[0] == name][0]
    sd = weights.make_state_dict(seed, tier, **cfg)
    x = torch.from_numpy(cases.fen_input(name)).to(dev)
    maps = _model(cfg, sd, dev).get_attention_maps(x)
    assert sorted(maps) == ["group0_rcab0", "group0_rcab1"]
    gold = torch.from_numpy(FEN_GOLD[name + "/se"])
    got = torch.stack([maps["group0_rcab0"], maps["group0_rcab1"]], 1).cpu()
    assert got.shape == gold.shape
    assert (got - gold).abs().max().item() <= 5e-3


def test_forward_other_shapes_and_repacking(dev):
    cfg = dict(num_groups=1, blocks_per_group=1)
    sd = weights.make_state_dict(4, "T1", **cfg)
    m = _model(cfg, sd, dev)
    x = torch.rand(2, 3, 128, 64, generator=torch.Generator().manual_seed(9))   # fully convolutional
    with torch.no_grad():
        y = m(x.to(dev)).cpu()
    ref = fen_oracle.fen_forward(sd, x)
    assert y.shape == (2, 3, 512, 256)
    assert fen_oracle.psnr(y, ref) >= PSNR_BAR and (y - ref).abs().max().item() <= MAXABS_BAR
    # in-place weight update must invalidate the packed copy
    with torch.no_grad():
        m.conv_last.weight.mul_(3.0)
        y2 = m(x.to(dev)).cpu()
    sd2 = {k: v.clone() for k, v in sd.items()}
    sd2["conv_last.weight"] = sd2["conv_last.weight"] * 3.0
    ref2 = fen_oracle.fen_forward(sd2, x)
    assert (y2 - ref2).abs().max().item() <= MAXABS_BAR and (y2 - y).abs().max().item() > 1e-4


def test_forward_error_behaviour(dev):
    m = _model(dict(num_groups=1, blocks_per_group=1), weights.make_state_dict(0, "T0", num_groups=1, blocks_per_group=1), dev)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        m(torch.rand(1, 3, 64, 64))
    with pytest.raises(RuntimeError, match="expected input"):
        m(torch.rand(1, 4, 64, 64, device=dev))
    m.train()                      # train mode with grad: the per-layer path with a grad_fn (tests/test_gpu_backward.py)
    assert m(torch.rand(1, 3, 64, 64, device=dev)).grad_fn is not None
    wide = fsr_b200.FaceEnhanceNet(num_channels=128, num_groups=1, blocks_per_group=1).to(dev).eval()
    with pytest.raises(ValueError, match="num_channels must be 64"):
        wide(torch.rand(1, 3, 64, 64, device=dev))


@pytest.mark.parametrize("name", [c[0] for c in cases.FEN2_CASES])
def test_forward_other_configs_match_reference_golden(name, dev):
    """the dataclass-default 3 x 4 config, FaceEnhanceNetLite (32 channels, embedded in the 64-channel
    kernels) and an input whose sides are not multiples of 64, against outputs of the unmodified reference."""
    gold = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "fen_golden2.npz"))
    _, ctor, cfg, tier, seed, shape = [c for c in cases.FEN2_CASES if c[0] == name][0]
    sd = weights.make_state_dict(seed, tier, **cfg)
    m = fsr_b200.FaceEnhanceNetLite() if ctor == "lite" else fsr_b200.FaceEnhanceNet(**cfg)
    m.load_state_dict(sd, strict=True)
    m = m.to(dev)
    x = torch.from_numpy(cases.fen2_input(name)).to(dev)
    ref = torch.from_numpy(gold[name + "/train"])
    with torch.no_grad():
        y_train = m.train()(x).cpu()
        y_eval = m.eval()(x).cpu()
    for got, r in ((y_train, ref), (y_eval, ref.clamp(0, 1))):
        assert got.shape == r.shape
        assert fen_oracle.psnr(got, r) >= PSNR_BAR and (got - r).abs().max().item() <= MAXABS_BAR
    if ctor == "lite":
        maps = m.get_attention_maps(x)
        assert maps["group0_rcab0"].shape == (1, 32)
        taps = {}
        fen_oracle.fen_forward(sd, x.cpu(), taps=taps)
        assert (maps["group2_rcab3"].cpu() - taps["se"][:, 11]).abs().max().item() <= 1e-2


@pytest.mark.parametrize("hw", [(17, 23), (64, 100), (100, 100), (128, 128), (33, 130)])
def test_forward_arbitrary_sizes_against_oracle(hw, dev):
    """The demo accepts any LR size up to 128 x 128 (app/demo.py:247-251); the network is fully convolutional."""
    cfg = dict(num_groups=1, blocks_per_group=2)
    sd = weights.make_state_dict(6, "T1", **cfg)
    g = torch.Generator().manual_seed(hw[0] * 1000 + hw[1])
    sd["conv_last.weight"] = torch.randn(sd["conv_last.weight"].shape, generator=g) * 1e-2
    x = torch.rand(2, 3, hw[0], hw[1], generator=g)
    m = _model(cfg, sd, dev, train=True)
    with torch.no_grad():
        y = m(x.to(dev)).cpu()
    ref = fen_oracle.fen_forward(sd, x, training=True)
    assert y.shape == ref.shape == (2, 3, 4 * hw[0], 4 * hw[1])
    assert not torch.isnan(y).any()
    assert fen_oracle.psnr(y, ref) >= 45.0 and (y - ref).abs().max().item() <= 3e-2   # (conv_last 10x the T1 scale)


def test_lite_and_ragged_backward_against_oracle(dev):
    """Gradients of a 32-channel model (gathered back from the 64-channel embedding) and of a ragged input size."""
    cfg = dict(num_groups=1, blocks_per_group=1, num_channels=32, reduction_ratio=2)
    sd = weights.make_state_dict(14, "T1", **cfg)
    m = fsr_b200.FaceEnhanceNet(**cfg)
    m.load_state_dict(sd, strict=True)
    m = m.to(dev).train()
    rng = np.random.default_rng(31)
    for shape in [(2, 3, 64, 64), (1, 3, 40, 72)]:
        x = torch.from_numpy(rng.random(shape, dtype=np.float32))
        yy, xx = np.mgrid[0:4 * shape[2], 0:4 * shape[3]].astype(np.float32)
        dout = (1.0 + 0.5 * np.sin(yy / 17.0) * np.cos(xx / 23.0))[None, None] * np.ones((shape[0], 3, 1, 1))
        dout = torch.from_numpy((dout / dout.size).astype(np.float32))
        m.zero_grad()
        m(x.to(dev)).backward(dout.to(dev))
        _, ref = fen_oracle.fen_backward(sd, x, dout)
        for k, p in m.named_parameters():
            rel = ((p.grad.cpu() - ref[k]).norm() / ref[k].norm().clamp_min(1e-30)).item()
            assert p.grad.shape == ref[k].shape and rel <= 2e-2, (shape, k, rel)


def test_forward_u8_is_the_scripts_output_path(dev):
    """scripts/test_model.py:176-190 (to_numpy): uint8 HWC = trunc(clip(sr * 255, 0, 255)), optionally BGR.  forward_u8
    does it in the conv_last epilogue; bit-identical to the separate conversion kernel and to numpy on the fp32 output."""
    cfg = dict(num_groups=1, blocks_per_group=2)
    sd = weights.make_state_dict(7, "T1", **cfg)
    g = torch.Generator().manual_seed(3)
    sd["conv_last.weight"] = torch.randn(sd["conv_last.weight"].shape, generator=g) * 2e-2   # some pixels clamp at both ends
    m = _model(cfg, sd, dev)
    for shape in [(3, 3, 64, 64), (1, 3, 40, 72)]:
        x = torch.rand(*shape, generator=g).to(dev)
        with torch.no_grad():
            y = m(x)
        for bgr in (False, True):
            u8 = m.forward_u8(x, bgr=bgr)
            assert u8.dtype == torch.uint8 and u8.shape == (shape[0], 4 * shape[2], 4 * shape[3], 3)
            assert torch.equal(u8, fsr_b200.sr_to_uint8(y, bgr=bgr))
            ref = np.clip(y.cpu().numpy() * 255.0, 0, 255).astype(np.uint8).transpose(0, 2, 3, 1)
            assert np.array_equal(u8.cpu().numpy(), ref[..., ::-1] if bgr else ref)
        assert int((u8 == 0).sum()) > 0 and int((u8 == 255).sum()) > 0


def test_two_models_on_two_streams_do_not_share_state(dev):
    """The per-layer bias / slope table of the body kernel travels with each launch (it used to live in one
    __constant__ bank that every fen_forward overwrote): two different models running concurrently on two streams give
    the results they give alone."""
    cfg = dict(num_groups=2, blocks_per_group=3)
    sds = [weights.make_state_dict(40 + i, "T1", **cfg) for i in range(2)]
    for sd in sds:
        sd["conv_last.weight"] = sd["conv_last.weight"] * 20.0
    ms = [_model(cfg, sd, dev) for sd in sds]
    x = torch.rand(16, 3, 64, 64, device=dev)
    with torch.no_grad():
        alone = [m(x).clone() for m in ms]
        assert (alone[0] - alone[1]).abs().max().item() > 1e-2
        streams = [torch.cuda.Stream(device=dev) for _ in range(2)]
        torch.cuda.synchronize()
        for _ in range(5):
            outs = []
            for m, st in zip(ms, streams):
                with torch.cuda.stream(st):
                    outs.append(m(x))
            torch.cuda.synchronize()
            for o, a in zip(outs, alone):
                assert (o - a).abs().max().item() <= 5e-3      # (SE-pool summation order only)


def test_lr_generator_feeds_the_model(dev):
    """Config 4's chain: uint8 HR -> integer LR kernel -> forward, vs the oracle chain."""
    hr = cases.lr_input("smooth_256")[None]
    cfg = dict(num_groups=1, blocks_per_group=2)
    sd = weights.make_state_dict(2, "T1", **cfg)
    _, lr = fsr_b200.lr_from_hr(torch.from_numpy(hr).to(dev))
    with torch.no_grad():
        y = _model(cfg, sd, dev)(lr).cpu()
    ref_lr = torch.from_numpy(lr_oracle.to_tensor_chw(lr_oracle.lr_from_hr_u8(hr)))
    assert torch.equal(lr.cpu(), ref_lr)
    ref = fen_oracle.fen_forward(sd, ref_lr)
    assert fen_oracle.psnr(y, ref) >= PSNR_BAR and (y - ref).abs().max().item() <= MAXABS_BAR


# ------------------------------------------------------------------ SSIM metric / loss
@pytest.mark.parametrize("name", [c[0] for c in cases.SSIM_CASES])
def test_ssim_matches_reference_golden(name, dev):
    """src/losses/ssim_loss.py:44-98 (ssim) and :166-226 (SSIMLoss): value within 2e-6 of the unmodified reference
    (fp32; separable evaluation of the same window), gradient within 1e-4 relative."""
    gold = np.load(os.path.join(HERE, "golden", "ssim_golden.npz"))
    _, shape, ws, sigma = [c for c in cases.SSIM_CASES if c[0] == name][0]
    pred, target = (torch.from_numpy(a).to(dev) for a in cases.ssim_inputs(name))
    m = fsr_b200.ssim(pred, target, window_size=ws, sigma=sigma)
    per = fsr_b200.ssim(pred, target, window_size=ws, sigma=sigma, size_average=False)
    assert m.shape == () and per.shape == (shape[0],)
    assert abs(m.item() - float(gold[name + "/mean"])) <= 2e-6
    assert np.abs(per.cpu().numpy() - gold[name + "/per_image"]).max() <= 2e-6
    if name + "/loss_grad" in gold.files:
        p = pred.clone().requires_grad_(True)
        loss = fsr_b200.SSIMLoss(window_size=ws, sigma=sigma, channel=shape[1]).to(dev)(p, target)
        assert abs(loss.item() - (1.0 - float(gold[name + "/mean"]))) <= 2e-6
        loss.backward()
        ref = torch.from_numpy(gold[name + "/loss_grad"])
        assert ((p.grad.cpu() - ref).norm() / ref.norm()).item() <= 1e-4
        # per-image means: the gradient of their sum is B times the gradient of the overall mean
        p2 = pred.clone().requires_grad_(True)
        fsr_b200.ssim(p2, target, window_size=ws, sigma=sigma, size_average=False).sum().backward()
        assert torch.allclose(p2.grad, -p.grad * shape[0], rtol=1e-4, atol=1e-9)
    with pytest.raises(ValueError):
        fsr_b200.ssim(pred, target, window_size=12)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        fsr_b200.ssim(pred.cpu(), target.cpu())


def test_forward_is_deterministic(dev):
    """The reference trains and evaluates with cudnn.deterministic = True (scripts/train.py:52-53).  The forward's only
    cross-CTA reduction (the SE pool) is accumulated in fixed point, so repeated runs agree bit for bit - in the fused
    body kernel (64-column input) and on the per-layer path (other widths)."""
    cfg = dict(num_groups=2, blocks_per_group=3)
    sd = weights.make_state_dict(50, "T1", **cfg)
    sd["conv_last.weight"] = sd["conv_last.weight"] * 20.0
    m = _model(cfg, sd, dev, train=True)
    for shape in [(24, 3, 64, 64), (3, 3, 64, 128)]:
        x = torch.rand(*shape, device=dev)
        with torch.no_grad():
            ref = m(x).clone()
            for _ in range(10):
                assert torch.equal(m(x), ref)


def test_forward_batch64_is_bit_repeatable(dev):
    """BASELINE config 2 itself (6 x 10 x 64, batch 64: 7 or 8 tiles per pass and CTA in the persistent body kernel, the
    geometry in which an issuer could run one mbarrier phase ahead of a ring slot - DESIGN.md 4.2): 400 forwards of one
    batch agree bit for bit with the first.  (tools/soak2.py is the long version: 12 000 forwards, every stage compared.)"""
    cfg = dict(num_groups=6, blocks_per_group=10)
    sd = weights.make_state_dict(3, "T1", **cfg)
    g = torch.Generator().manual_seed(11)
    sd["conv_last.weight"] = torch.randn(sd["conv_last.weight"].shape, generator=g) * 1e-2   # expose the body
    m = _model(cfg, sd, dev)
    x = torch.rand(64, 3, 64, 64, device=dev)
    with torch.no_grad():
        ref = m(x).clone()
        bad = sum(int(not torch.equal(m(x), ref)) for _ in range(400))
    assert bad == 0, f"{bad} of 400 batch-64 forwards differ from the first"
