"""Pins the fp32 network oracle (oracle/fen_oracle.py) against outputs of the UNMODIFIED reference
module (tests/golden/fen_golden.npz, made by tests/golden/make_golden.py from /root/reference)."""
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

import cases
from oracle import fen_oracle, weights

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = np.load(os.path.join(HERE, "golden", "fen_golden.npz"))
TOL = 2e-5  # fp32 reassociation only (the oracle calls the same ATen ops)


@pytest.mark.parametrize("name", [c[0] for c in cases.FEN_CASES])
def test_oracle_matches_reference_golden(name):
    _, cfg, tier, seed, _ = [c for c in cases.FEN_CASES if c[0] == name][0]
    sd = weights.make_state_dict(seed, tier, **cfg)
    x = torch.from_numpy(cases.fen_input(name))
    taps = {}
    y_train = fen_oracle.fen_forward(sd, x, training=True, taps=taps)
    y_eval = fen_oracle.fen_forward(sd, x, training=False)
    g_train = torch.from_numpy(GOLD[name + "/train"])
    assert (y_train - g_train).abs().max().item() <= TOL
    assert (y_eval - g_train.clamp(0, 1)).abs().max().item() <= TOL
    assert y_eval.min() >= 0 and y_eval.max() <= 1
    assert (taps["se"] - torch.from_numpy(GOLD[name + "/se"])).abs().max().item() <= 1e-5


@pytest.mark.parametrize("name", cases.GRAD_CASES)
def test_oracle_backward_matches_reference_golden(name):
    # the gradient oracle (autograd over the restatement) against .grad of the unmodified reference module
    gold = np.load(os.path.join(HERE, "golden", "fen_grad_golden.npz"))
    _, cfg, tier, seed, _ = [c for c in cases.FEN_CASES if c[0] == name][0]
    sd = weights.make_state_dict(seed, tier, **cfg)
    sr, grads = fen_oracle.fen_backward(sd, torch.from_numpy(cases.fen_input(name)),
                                        torch.from_numpy(cases.grad_dout(name)))
    assert (sr - torch.from_numpy(GOLD[name + "/train"])).abs().max().item() <= TOL
    assert set(grads) == set(sd)
    for k, g in grads.items():
        ref = torch.from_numpy(gold[name + "/" + k])
        assert g.shape == ref.shape
        assert (g - ref).norm().item() <= 1e-4 * ref.norm().item() + 1e-12, k
    # nn.L1Loss gradient helper
    a, b = torch.tensor([[0.2, 0.9]]), torch.tensor([[0.5, 0.1]])
    assert torch.equal(fen_oracle.l1_grad(a, b), torch.tensor([[-0.5, 0.5]]))


def test_literal_init_is_vacuous_without_the_weight_recipe():
    # trap 1 of SURVEY.md: conv_last == 0  =>  forward == clamp(bicubic_up(x))
    cfg = dict(num_groups=1, blocks_per_group=1)
    sd = weights.make_state_dict(3, "T0", **cfg)
    x = torch.rand(1, 3, 64, 64, generator=torch.Generator().manual_seed(0))
    y = fen_oracle.fen_forward(sd, x)
    bic = F.interpolate(x, scale_factor=4, mode="bicubic", align_corners=False).clamp(0, 1)
    assert torch.equal(y, bic)
    sd1 = weights.make_state_dict(3, "T1", **cfg)
    assert (fen_oracle.fen_forward(sd1, x) - bic).abs().max() > 1e-3


def test_bicubic_phase_filters_match_interpolate():
    # the integer/2048 phase filters the CUDA epilogue uses (custom.py:158-161)
    w, off = fen_oracle.bicubic_x4_weights()
    x = torch.rand(1, 1, 9, 11, dtype=torch.float64, generator=torch.Generator().manual_seed(1))
    ref = F.interpolate(x, scale_factor=4, mode="bicubic", align_corners=False)[0, 0]
    h, wd = x.shape[2:]
    got = torch.zeros_like(ref)
    for Y in range(4 * h):
        for X in range(4 * wd):
            qy, ry, qx, rx = Y // 4, Y % 4, X // 4, X % 4
            acc = 0.0
            for i in range(4):
                yy = min(max(qy + off[ry] - 1 + i, 0), h - 1)
                for j in range(4):
                    xx = min(max(qx + off[rx] - 1 + j, 0), wd - 1)
                    acc += w[ry][i] * w[rx][j] * x[0, 0, yy, xx].item()
            got[Y, X] = acc / (2048.0 * 2048.0)
    assert (got - ref).abs().max().item() < 1e-12


def test_schema_matches_survey_counts():
    s = weights.state_dict_schema()
    assert len(s) == 444
    assert sum(int(np.prod(v)) for v in s.values()) == 5_115_651
    assert s["residual_groups.5.blocks.9.channel_attention.fc.0.weight"] == (16, 64)
    assert s["upsample.stages.1.conv.weight"] == (256, 64, 3, 3)
    assert fen_oracle.count_groups_blocks(weights.make_state_dict(0, "T0", num_groups=2, blocks_per_group=3)) == (2, 3)


def test_weight_recipe_is_deterministic_and_tiered():
    a = weights.make_state_dict(7, "T1", num_groups=1, blocks_per_group=1)
    b = weights.make_state_dict(7, "T1", num_groups=1, blocks_per_group=1)
    assert all(torch.equal(a[k], b[k]) for k in a)
    t0 = weights.make_state_dict(7, "T0", num_groups=1, blocks_per_group=1)
    assert t0["conv_last.weight"].abs().max() == 0 and t0["conv_first.bias"].abs().max() == 0
    assert torch.all(t0["residual_groups.0.blocks.0.prelu.weight"] == 0.25)
    assert a["conv_last.weight"].std() < 2e-3 and a["conv_last.weight"].abs().max() > 0
    with pytest.raises(ValueError):
        weights.make_state_dict(0, "T9")


def test_psnr_formula():
    a = torch.zeros(4, 4)
    assert fen_oracle.psnr(a, a) == float("inf")
    assert abs(fen_oracle.psnr(a, a + 0.1) - 20.0) < 1e-4


# ---- round 2 goldens: default 3 x 4 config, FaceEnhanceNetLite (32 channels), a ragged input size, negative slopes
GOLD2 = np.load(os.path.join(HERE, "golden", "fen_golden2.npz"))


@pytest.mark.parametrize("name", [c[0] for c in cases.FEN2_CASES])
def test_oracle_matches_reference_golden_round2(name):
    _, _, cfg, tier, seed, shape = [c for c in cases.FEN2_CASES if c[0] == name][0]
    sd = weights.make_state_dict(seed, tier, **cfg)
    y = fen_oracle.fen_forward(sd, torch.from_numpy(cases.fen2_input(name)), training=True)
    ref = torch.from_numpy(GOLD2[name + "/train"])
    assert y.shape == ref.shape == (shape[0], 3, 4 * shape[2], 4 * shape[3])
    assert (y - ref).abs().max().item() <= TOL


def test_oracle_backward_with_negative_slopes_matches_reference_golden():
    sd = cases.negative_slopes(weights.make_state_dict(cases.NEG_GRAD_SEED, "T1", **cases.NEG_GRAD_CFG), cases.NEG_GRAD_SEED)
    x, dout = cases.neg_grad_inputs()
    sr, grads = fen_oracle.fen_backward(sd, torch.from_numpy(x), torch.from_numpy(dout))
    assert (sr - torch.from_numpy(GOLD2["neg/train"])).abs().max().item() <= TOL
    stored = [k for k in GOLD2.files if k.startswith("neg/grad/")]
    assert len(stored) >= 10
    for key in stored:
        ref = torch.from_numpy(GOLD2[key])
        g = grads[key[len("neg/grad/"):]]
        assert (g - ref).norm().item() <= 1e-4 * ref.norm().item() + 1e-12, key
