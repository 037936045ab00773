"""Pins the LR-generator oracle (numpy restatement + C restatement) against the golden vectors made
with cv2 4.13.0 (tests/golden/make_golden.py), i.e. against the call the reference makes at
src/data/dataset.py:296 and src/data/prepare_data.py:38."""
import ctypes
import os

import numpy as np
import pytest

import cases
from oracle import lr_oracle

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = np.load(os.path.join(HERE, "golden", "lr_golden.npz"))


def _c_oracle(built_lib):
    so = os.path.join(os.path.dirname(HERE), "oracle", "liblr_oracle.so")
    lib = ctypes.CDLL(so)
    lib.lr_oracle_u8.restype = ctypes.c_int
    lib.lr_oracle_u8.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_long, ctypes.c_int, ctypes.c_int,
                                 ctypes.c_int]
    return lib


@pytest.mark.parametrize("name", [c[0] for c in cases.LR_CASES])
def test_numpy_oracle_matches_cv2_golden(name):
    hr = cases.lr_input(name)
    assert np.array_equal(lr_oracle.lr_from_hr_u8(hr), GOLD[name])


@pytest.mark.parametrize("name", [c[0] for c in cases.LR_CASES])
def test_c_oracle_matches_cv2_golden(name, built_lib):
    lib = _c_oracle(built_lib)
    hr = cases.lr_input(name)
    H, W, C = hr.shape
    out = np.empty((H // 4, W // 4, C), np.uint8)
    assert lib.lr_oracle_u8(hr.ctypes.data, out.ctypes.data, 1, H, W, C) == 0
    assert np.array_equal(out, GOLD[name])


def test_c_oracle_rejects_bad_sizes(built_lib):
    lib = _c_oracle(built_lib)
    buf = np.zeros(64, np.uint8)
    assert lib.lr_oracle_u8(buf.ctypes.data, buf.ctypes.data, 1, 6, 8, 1) == -1
    assert lib.lr_oracle_u8(None, buf.ctypes.data, 1, 8, 8, 1) == -1


def test_loop_statement_agrees_on_small_input():
    hr = cases.lr_input("ties_256")[:32, :48]
    assert np.array_equal(lr_oracle.lr_from_hr_u8_loops(hr), lr_oracle.lr_from_hr_u8(hr))


def test_batched_and_ragged_shapes():
    rng = np.random.default_rng(5)
    hr = rng.integers(0, 256, (3, 2, 8, 12, 2), dtype=np.uint8)
    out = lr_oracle.lr_from_hr_u8(hr)
    assert out.shape == (3, 2, 2, 3, 2)
    assert np.array_equal(out[1, 1], lr_oracle.lr_from_hr_u8(hr[1, 1]))
    empty = lr_oracle.lr_from_hr_u8(np.zeros((0, 8, 8, 3), np.uint8))
    assert empty.shape == (0, 2, 2, 3)
    with pytest.raises(ValueError):
        lr_oracle.lr_from_hr_u8(np.zeros((6, 8, 3), np.uint8))
    with pytest.raises(TypeError):
        lr_oracle.lr_from_hr_u8(np.zeros((8, 8, 3), np.float32))


def test_constant_blocks_are_fixed_points():
    # weights sum to 32*32 = 1024, so a constant 4x4 block maps to itself
    hr = cases.lr_input("blocks_256")
    assert np.array_equal(lr_oracle.lr_from_hr_u8(hr), hr[::4, ::4])


def test_half_even_rounding_differs_from_textbook_opencv_formula():
    # trap 4 of SURVEY.md: (u * 4096 + 2^21) >> 22 == floor(u/1024 + 0.5) is NOT what cv2 does on ties
    hr = cases.lr_input("ties_256")
    x = hr.reshape(64, 4, 64, 4, 3).astype(np.int64)
    a = np.array([-3, 19, 19, -3])
    u = np.einsum("yixjc,i,j->yxc", x, a, a)
    naive = np.clip((u + 512) >> 10, 0, 255)
    assert (naive != GOLD["ties_256"]).sum() > 0
    assert np.array_equal(lr_oracle.lr_from_hr_u8(hr), GOLD["ties_256"])


def test_to_tensor_matches_reference_formula():
    lr = GOLD["random_256"]
    t = lr_oracle.to_tensor_chw(lr)
    assert t.shape == (3, 64, 64) and t.dtype == np.float32
    assert np.array_equal(t, np.transpose(lr, (2, 0, 1)).astype(np.float32) / np.float32(255.0))


def test_live_cv2_if_available():
    cv2 = pytest.importorskip("cv2")
    rng = np.random.default_rng(77)
    for _ in range(4):
        hr = rng.integers(0, 256, (256, 256, 3), dtype=np.uint8)
        ref = cv2.resize(hr, (64, 64), interpolation=cv2.INTER_CUBIC)
        assert np.array_equal(lr_oracle.lr_from_hr_u8(hr), ref)


def test_float_lr_taps_restate_the_trainer_call():
    """trainer.py:416-421 is F.interpolate(scale_factor=0.25, bicubic): 16 fixed taps at the /4 ratio."""
    import torch
    g = torch.Generator().manual_seed(3)
    hr = torch.rand(2, 3, 64, 96, generator=g)
    ref = lr_oracle.lr_from_hr_float(hr).numpy()
    taps = lr_oracle.lr_from_hr_float_taps(hr.numpy())
    assert ref.shape == (2, 3, 16, 24)
    assert np.abs(ref - taps).max() <= 2.4e-7
    # no rounding, no clamp: a checkerboard of 0 / 1 leaves [0, 1]
    cb = torch.zeros(1, 1, 8, 8)
    cb[..., 1:3, 1:3] = 1.0
    out = lr_oracle.lr_from_hr_float(cb)
    assert out.max() > 1.0 and lr_oracle.lr_from_hr_float(1 - cb).min() < 0.0


def test_quantize_u8_truncates_and_reorders():
    x = np.array([[[0.9999, -0.2]], [[0.5, 1.2]], [[1.0 / 255 * 7.9, 0.0]]], dtype=np.float32)   # [C=3, H=1, W=2]
    q = lr_oracle.quantize_u8_hwc(x)
    assert q.shape == (1, 2, 3) and q.tolist() == [[[254, 127, 7], [0, 255, 0]]]
    assert lr_oracle.quantize_u8_hwc(x, bgr=True).tolist() == [[[7, 127, 254], [0, 255, 0]]]
