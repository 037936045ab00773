"""Parity of the Stage-1 step's network side (fen_forward_train + fen_backward through the C ABI, reached the way
the reference's trainer reaches it: `sr = model(lr); loss.backward()`, src/training/trainer.py:462-488) against
.grad of the unmodified reference module (tests/golden/fen_grad_golden.npz) and the fp32 autograd oracle.

Tolerance (floating point; BASELINE.json states none for gradients, so it is stated here): activations and data
gradients are bf16 on the GPU, parameter gradients are accumulated in fp32.  Two regimes:

* COHERENT output gradient (one sign, smooth magnitude): every sum in the backward adds up coherently, so the
  error is the bf16 rounding of the operands: per parameter tensor relative L2 error <= 2e-2 (measured <= 0.85 %).
* SIGN-PATTERN output gradient (what nn.L1Loss produces, +-1/numel): every parameter gradient is a sqrt(N)-sized
  random-sign sum, and the ~0.4 % of PReLU inputs whose sign differs between the bf16 and the fp32 FORWARD each
  change one term by (1 - slope): that alone is sqrt(0.004) ~ 4 % per PReLU layer crossed, added in quadrature
  (measured 0.5 % at conv_last, 3.7 % after one PReLU, 7.5 % after four).  Bar: per tensor <= 0.2.
Both: the whole flat gradient has cosine >= 0.999 and a norm within 2 % of the fp32 reference."""
import os

import numpy as np
import pytest
import torch

import cases
import fsr_b200
from oracle import fen_oracle, weights

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))
REL_COHERENT, REL_SIGN, COS_BAR, NORM_BAR = 2e-2, 0.2, 0.999, 2e-2


@pytest.fixture(scope="module")
def dev(built_lib):
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    return torch.device("cuda:0")


def _compare(named_grads, ref_grads, rel_bar, whole=True, report_to=None):
    rows, bad = [], []
    num = den_a = den_b = 0.0
    for k, g in named_grads:
        r = ref_grads[k].double()
        g = g.detach().double().cpu()
        assert g.shape == r.shape, k
        rel = (g - r).norm().item() / max(r.norm().item(), 1e-30)
        num += float((g * r).sum()); den_a += float((g * g).sum()); den_b += float((r * r).sum())
        rows.append(f"{k:60s} ref {r.norm().item():.3e} got {g.norm().item():.3e} rel {rel:.3e}")
        if rel_bar is not None and not rel <= rel_bar:
            bad.append(rows[-1])
    cos = num / max((den_a * den_b) ** 0.5, 1e-300)
    ratio = (den_a / max(den_b, 1e-300)) ** 0.5
    report = "\n".join(rows) + f"\ncosine {cos:.6f} norm ratio {ratio:.4f}"
    print(report)
    if report_to:
        os.makedirs(os.path.dirname(report_to), exist_ok=True)
        with open(report_to, "w") as f:
            f.write(report + "\n")
    if rel_bar is not None:
        assert not bad, "per-tensor relative error above the bar:\n" + "\n".join(bad) + "\n\nall:\n" + report
    if whole:
        assert cos >= COS_BAR and abs(ratio - 1.0) <= NORM_BAR, report
    return cos, ratio


def _model(cfg, sd, dev):
    m = fsr_b200.FaceEnhanceNet(**cfg)
    m.load_state_dict(sd, strict=True)
    return m.to(dev).train()


@pytest.mark.parametrize("name", cases.GRAD_CASES)
def test_backward_matches_reference_golden(name, dev):
    gold = np.load(os.path.join(HERE, "golden", "fen_grad_golden.npz"))
    fwd_gold = np.load(os.path.join(HERE, "golden", "fen_golden.npz"))
    _, cfg, tier, seed, _ = [c for c in cases.FEN_CASES if c[0] == name][0]
    m = _model(cfg, weights.make_state_dict(seed, tier, **cfg), dev)
    x = torch.from_numpy(cases.fen_input(name)).to(dev)
    sr = m(x)
    assert sr.requires_grad and sr.grad_fn is not None
    ref_sr = torch.from_numpy(fwd_gold[name + "/train"])
    assert fen_oracle.psnr(sr.detach().cpu(), ref_sr) >= 50.0      # the per-layer train-mode forward, unclamped
    sr.backward(torch.from_numpy(cases.grad_dout(name)).to(dev))
    ref = {k: torch.from_numpy(gold[name + "/" + k]) for k, _ in m.named_parameters()}
    _compare([(k, p.grad) for k, p in m.named_parameters()], ref, REL_SIGN)


@pytest.mark.parametrize("cfg,batch,hw", [(dict(num_groups=1, blocks_per_group=2), 2, (64, 64)),
                                          (dict(num_groups=2, blocks_per_group=1), 1, (64, 64)),
                                          (dict(num_groups=1, blocks_per_group=1), 1, (64, 128)),    # two strips per row
                                          (dict(num_groups=1, blocks_per_group=1), 5, (128, 64))])   # odd batch, tall
def test_backward_coherent_gradient_against_oracle(cfg, batch, hw, dev):
    sd = weights.make_state_dict(9, "T1", **cfg)
    rng = np.random.default_rng(123)
    x = torch.from_numpy(rng.random((batch, 3, hw[0], hw[1]), dtype=np.float32))
    yy, xx = np.mgrid[0:4 * hw[0], 0:4 * hw[1]].astype(np.float32)
    dout = (1.0 + 0.5 * np.sin(yy / 17.0)[None, None] * np.cos(xx / 23.0)[None, None]) * np.ones((batch, 3, 1, 1))
    dout = torch.from_numpy((dout / dout.size).astype(np.float32))
    m = _model(cfg, sd, dev)
    m(x.to(dev)).backward(dout.to(dev))
    _, ref = fen_oracle.fen_backward(sd, x, dout)
    _compare([(k, p.grad) for k, p in m.named_parameters()], ref, REL_COHERENT)


def test_backward_two_groups_against_oracle_with_l1_loss(dev):
    # group skip, long skip and RCAB chaining across groups; loss = nn.L1Loss as in the trainer
    cfg = dict(num_groups=2, blocks_per_group=2)
    sd = weights.make_state_dict(5, "T1", **cfg)
    rng = np.random.default_rng(77)
    x = torch.from_numpy(rng.random((3, 3, 64, 64), dtype=np.float32))
    hr = torch.from_numpy(rng.random((3, 3, 256, 256), dtype=np.float32))
    m = _model(cfg, sd, dev)
    sr = m(x.to(dev))
    loss = torch.nn.L1Loss()(sr, hr.to(dev))
    loss.backward()
    # same d loss / d sr for both sides: the sign pattern of the GPU output (the forward has its own parity bar)
    dout = fen_oracle.l1_grad(sr.detach().cpu(), hr)
    sr_ref, ref = fen_oracle.fen_backward(sd, x, dout)
    assert fen_oracle.psnr(sr.detach().cpu(), sr_ref) >= 50.0
    assert abs(loss.item() - (sr_ref - hr).abs().mean().item()) <= 1e-4
    _compare([(k, p.grad) for k, p in m.named_parameters()], ref, REL_SIGN)


def test_train_mode_semantics(dev):
    cfg = dict(num_groups=1, blocks_per_group=1)
    sd = weights.make_state_dict(2, "T1", **cfg)
    m = _model(cfg, sd, dev)
    x = torch.rand(2, 3, 64, 64, device=dev)
    with torch.no_grad():
        y_ng = m(x)                       # train mode, no grad: fused inference kernels, unclamped
    assert not y_ng.requires_grad
    y = m(x)                              # train mode with grad: per-layer path
    assert y.requires_grad
    assert (y.detach() - y_ng).abs().max().item() <= 2e-2
    m.eval()
    y_eval = m(x)
    assert not y_eval.requires_grad and y_eval.min() >= 0 and y_eval.max() <= 1
    # gradients accumulate over two backward calls like torch's (.grad += )
    m.train()
    m.zero_grad()
    m(x).sum().backward()
    g1 = [p.grad.clone() for p in m.parameters()]
    m(x).sum().backward()
    for a, p in zip(g1, m.parameters()):
        assert torch.allclose(p.grad, 2 * a, rtol=1e-3, atol=1e-4 * float(a.abs().max()) + 1e-12)
    # two outstanding train-mode forwards (each owns its saved activations), as nn.Module allows
    x2 = torch.rand(2, 3, 64, 64, device=dev)
    m.zero_grad(); m(x).sum().backward(); ga = [p.grad.clone() for p in m.parameters()]
    m.zero_grad(); m(x2).sum().backward(); gb = [p.grad.clone() for p in m.parameters()]
    m.zero_grad()
    y1, y2 = m(x), m(x2)
    (y1.sum() + y2.sum()).backward()
    for a, b, p in zip(ga, gb, m.parameters()):
        assert torch.allclose(p.grad, a + b, rtol=1e-3, atol=1e-4 * float((a + b).abs().max()) + 1e-12)


def test_backward_with_negative_prelu_slopes(dev):
    """nn.PReLU allows slopes of any sign (blocks.py:127,146,216): the backward takes the sign of the PRE-activation
    from the bit masks the forward saves.  Reference gradients: tests/golden/fen_golden2.npz (unmodified module)."""
    gold = np.load(os.path.join(HERE, "golden", "fen_golden2.npz"))
    sd = cases.negative_slopes(weights.make_state_dict(cases.NEG_GRAD_SEED, "T1", **cases.NEG_GRAD_CFG), cases.NEG_GRAD_SEED)
    assert sum(int((v < 0).sum()) for k, v in sd.items() if k.endswith("prelu.weight")) > 50
    x, dout = cases.neg_grad_inputs()
    m = _model(cases.NEG_GRAD_CFG, sd, dev)
    sr = m(torch.from_numpy(x).to(dev))
    assert fen_oracle.psnr(sr.detach().cpu(), torch.from_numpy(gold["neg/train"])) >= 50.0
    sr.backward(torch.from_numpy(dout).to(dev))
    named = dict(m.named_parameters())
    keys = [k[len("neg/grad/"):] for k in gold.files if k.startswith("neg/grad/")]
    ref = {k: torch.from_numpy(gold["neg/grad/" + k]) for k in keys}
    _compare([(k, named[k].grad) for k in keys], ref, REL_SIGN, whole=False)
    # the eval-mode forward with negative slopes too (fused body kernel)
    m.eval()
    with torch.no_grad():
        y = m(torch.from_numpy(x).to(dev)).cpu()
    assert fen_oracle.psnr(y, torch.from_numpy(gold["neg/train"]).clamp(0, 1)) >= 50.0


def test_stage1_step_is_the_trainers_iteration(dev):
    # Trainer._train_epoch (trainer.py:410-505): float LR -> forward -> L1 -> backward -> clip 0.5 -> AdamW(1e-4)
    cfg = dict(num_groups=1, blocks_per_group=2)
    sd = weights.make_state_dict(4, "T1", **cfg)
    m = _model(cfg, sd, dev)
    step = fsr_b200.Stage1Step(m, lr=1e-4, max_norm=0.5)
    assert all(p.data_ptr() >= step.flat.data_ptr() for p in m.parameters())      # parameters are views of `flat`
    hr = torch.rand(2, 3, 256, 256, device=dev, generator=torch.Generator(device=dev).manual_seed(3))
    before = step.flat.clone()
    loss, norm = step.step(hr)
    g = step.last_grad
    # the same step with torch: oracle forward for the loss, torch AdamW on the gradient the kernels produced
    lr_img = torch.nn.functional.interpolate(hr.cpu(), scale_factor=0.25, mode="bicubic", align_corners=False)
    sr_ref = fen_oracle.fen_forward(sd, lr_img, training=True)
    assert abs(loss.item() - (sr_ref - hr.cpu()).abs().mean().item()) <= 1e-4
    assert abs(norm.item() - g.norm().item()) <= 1e-4 * g.norm().item()
    p = before.clone().requires_grad_(True)
    p.grad = g * min(1.0, 0.5 / (g.norm().item() + 1e-6))
    torch.optim.AdamW([p], lr=1e-4, weight_decay=0.0).step()
    assert torch.allclose(step.flat, p.detach(), rtol=0, atol=2e-7)
    assert (step.flat - before).abs().max().item() > 5e-5          # first AdamW step moves every weight by ~lr
    # next forward uses the updated weights (packed copies rebuilt), state_dict keeps the reference schema
    l2, _ = step.step(hr)
    assert torch.isfinite(l2).all() and l2.item() < loss.item()
    assert list(m.state_dict().keys()) == list(sd.keys())


def test_stage1_step_uses_fresh_weights_in_every_pass(dev):
    """Steps 2 and 3 must differentiate the UPDATED weights in forward, data-gradient and weight-gradient kernels
    alike (the transposed copies of the data-gradient convolutions are a separate cache).  A coherent (positive) d
    loss / d sr keeps the comparison at bf16-rounding level; lr is large so that stale weights would be far off."""
    cfg = dict(num_groups=1, blocks_per_group=2)
    sd = weights.make_state_dict(8, "T1", **cfg)
    m = _model(cfg, sd, dev)
    step = fsr_b200.Stage1Step(m, lr=3e-3, max_norm=0.5)
    hr = torch.rand(2, 3, 256, 256, device=dev, generator=torch.Generator(device=dev).manual_seed(11))
    for _ in range(2):
        step.step(hr)
    # third pass by hand on the step's own (twice updated) weights, with a coherent output gradient
    sd_now = {k: v.detach().cpu().clone() for k, v in m.state_dict().items()}
    assert max((sd_now[k] - sd[k]).abs().max().item() for k in sd) > 1e-3
    x = torch.rand(2, 3, 64, 64, device=dev, generator=torch.Generator(device=dev).manual_seed(12))
    yy, xx = np.mgrid[0:256, 0:256].astype(np.float32)
    dout = torch.from_numpy(((1.0 + 0.5 * np.sin(yy / 17.0) * np.cos(xx / 23.0)) / (2 * 3 * 256 * 256)).astype(np.float32))
    dout = dout.expand(2, 3, 256, 256).contiguous()
    sr, lease = m._forward_train(x)
    g = m._backward(x, dout.to(dev), lease)
    _, ref = fen_oracle.fen_backward(sd_now, x.cpu(), dout)
    named, off = [], 0
    for k, p in m.named_parameters():
        named.append((k, g[off:off + p.numel()].view(p.shape))); off += p.numel()
    _compare(named, ref, REL_COHERENT)


def test_backward_full_model_config5_against_oracle(dev):
    """BASELINE config 5's network (6 groups x 10 RCAB) with nn.L1Loss, batch 2: whole-gradient cosine >= 0.999 and
    norm within 2 % of the fp32 autograd oracle.  Per-tensor errors are printed and written to
    gpurun_out/grad_parity_6x10.txt (committed under profiles/): with a sign-pattern d loss / d sr they grow with
    the number of PReLU layers between the tensor and the output, so only the whole gradient has a bar here."""
    cfg = dict(num_groups=6, blocks_per_group=10)
    sd = weights.make_state_dict(0, "T1", **cfg)
    rng = np.random.default_rng(2025)
    x = torch.from_numpy(rng.random((2, 3, 64, 64), dtype=np.float32))
    hr = torch.from_numpy(rng.random((2, 3, 256, 256), dtype=np.float32))
    m = _model(cfg, sd, dev)
    sr = m(x.to(dev))
    loss = torch.nn.L1Loss()(sr, hr.to(dev))
    loss.backward()
    dout = fen_oracle.l1_grad(sr.detach().cpu(), hr)
    sr_ref, ref = fen_oracle.fen_backward(sd, x, dout)
    assert fen_oracle.psnr(sr.detach().cpu(), sr_ref) >= 50.0
    assert abs(loss.item() - (sr_ref - hr).abs().mean().item()) <= 1e-4
    out = os.path.join(os.path.dirname(HERE), "gpurun_out", "grad_parity_6x10.txt")
    cos, ratio = _compare([(k, p.grad) for k, p in m.named_parameters()], ref, None, whole=True, report_to=out)
    # early layers: how far is the worst tensor (informational bar, 60 PReLU layers deep)
    worst = max(((p.grad.detach().double().cpu() - ref[k].double()).norm() / ref[k].double().norm().clamp_min(1e-30)).item()
                for k, p in m.named_parameters())
    print(f"6x10 config-5 gradient: cosine {cos:.6f}, norm ratio {ratio:.4f}, worst per-tensor relative error {worst:.3f}")
    assert worst <= 0.6


def test_backward_rejects_bad_arguments(dev, built_lib):
    import ctypes as C
    from fsr_b200 import _lib
    cfg = _lib.FenConfig(64, 1, 1, 4, 4, 0.2)
    buf = torch.zeros(1024, dtype=torch.uint8, device=dev)
    p = buf.data_ptr()
    assert built_lib.fen_forward_train(C.byref(cfg), p, p, p, 1, 64, 64, p, 1024, None) == _lib.FEN_ENOMEM
    assert built_lib.fen_backward(C.byref(cfg), p, p, p, p, p, 1, 64, 400, p, 1 << 40, None) == _lib.FEN_EINVAL   # > 336 columns
    assert built_lib.fen_backward(C.byref(cfg), p, None, p, p, p, 1, 64, 64, p, 1 << 40, None) == _lib.FEN_EINVAL
    assert built_lib.fen_step_workspace_bytes(C.byref(cfg), 2, 64, 64) > built_lib.fen_forward_workspace_bytes(
        C.byref(cfg), 2, 64, 64)


@pytest.mark.parametrize("cfg,batch", [(dict(num_groups=2, blocks_per_group=2), 5), (dict(num_groups=6, blocks_per_group=10), 8)])
def test_backward_is_bit_deterministic(cfg, batch, dev):
    """scripts/train.py:52-53 trains with cudnn.deterministic = True.  Every cross-CTA reduction of the backward is
    order-independent here (weight gradients: per-CTA partials summed in CTA order; slope / SE-matrix / SE dot-product
    sums: 64-bit fixed-point integer atomics): repeated forward + backward passes give bit-identical gradients."""
    sd = weights.make_state_dict(3, "T1", **cfg)
    m = _model(cfg, sd, dev)
    g = torch.Generator().manual_seed(11)
    x = torch.rand(batch, 3, 64, 64, generator=g).to(dev)
    dout = ((torch.rand(batch, 3, 256, 256, generator=g) - 0.5) * 1e-4).to(dev)
    ref = None
    for it in range(4):
        m.zero_grad(set_to_none=True)
        m(x).backward(dout)
        flat = torch.cat([p.grad.reshape(-1) for p in m.parameters()])
        assert torch.isfinite(flat).all()
        if ref is None:
            ref = flat.clone()
            ref_named = {k: p.grad.clone() for k, p in m.named_parameters()}
            assert float(ref.abs().max()) > 0.0
        else:
            if not torch.equal(flat, ref):
                names = [k for k, p in m.named_parameters() if not torch.equal(p.grad, ref_named[k])]
                raise AssertionError(f"run {it}: {int((flat != ref).sum())} of {flat.numel()} gradient elements differ, in {names[:12]}")
