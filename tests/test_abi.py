"""The C-ABI library loads on a CPU-only box, exports every symbol include/fen_b200.h declares, and
its compute entry points fail loudly (FEN_ENODEV) instead of falling back when there is no GPU."""
import ctypes as C
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_functions():
    text = open(os.path.join(ROOT, "include", "fen_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(fen_[a-z0-9_]+)\s*\(", text)))


def test_header_declares_the_expected_entry_points():
    names = _declared_functions()
    for must in ("fen_forward", "fen_pack_weights", "fen_lr_from_hr_u8", "fen_conv3x3_c64",
                 "fen_forward_workspace_bytes", "fen_last_error"):
        assert must in names


def test_library_exports_every_declared_symbol(built_lib):
    from fsr_b200 import _lib
    raw = C.CDLL(_lib.LIB_PATH)
    for name in _declared_functions():
        assert hasattr(raw, name), f"{name} declared in fen_b200.h but not exported"
    assert sorted(_lib.SIGNATURES) == _declared_functions()  # ctypes binding covers the whole header
    assert built_lib.fen_abi_version() == 1


def test_layout_queries_need_no_gpu(built_lib):
    from fsr_b200 import _lib
    cfg = _lib.FenConfig(64, 6, 10, 4, 4, 0.2)
    assert built_lib.fen_param_count(C.byref(cfg)) == 5_115_651
    assert built_lib.fen_packed_bytes(C.byref(cfg)) > 5_115_651 * 2
    ws64 = built_lib.fen_forward_workspace_bytes(C.byref(cfg), 64, 64, 64)
    ws1 = built_lib.fen_forward_workspace_bytes(C.byref(cfg), 1, 64, 64)
    assert ws64 > 64 * 256 * 256 * 64 * 2 and ws1 < ws64
    cfg3 = _lib.FenConfig(64, 3, 4, 4, 4, 0.2)
    assert built_lib.fen_param_count(C.byref(cfg3)) < 5_115_651


def test_unsupported_configs_are_rejected(built_lib):
    from fsr_b200 import _lib
    for bad in (_lib.FenConfig(32, 3, 4, 2, 4, 0.2), _lib.FenConfig(64, 3, 4, 4, 2, 0.2),
                _lib.FenConfig(64, 0, 4, 4, 4, 0.2)):
        assert built_lib.fen_param_count(C.byref(bad)) == _lib.FEN_EINVAL
        assert built_lib.fen_last_error()


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_compute_entry_points_fail_loudly_without_gpu(built_lib):
    from fsr_b200 import _lib
    cfg = _lib.FenConfig(64, 1, 1, 4, 4, 0.2)
    buf = (C.c_uint8 * 64)()
    rc = built_lib.fen_forward(C.byref(cfg), buf, buf, buf, 1, 64, 64, 0, buf, 64, None, None)
    assert rc == _lib.FEN_ENODEV
    assert b"no CPU fallback" in built_lib.fen_last_error()
    assert built_lib.fen_lr_from_hr_u8(buf, buf, None, 1, 4, 4, 1, None) == _lib.FEN_ENODEV
    # the training-step entry points have no CPU path either
    assert built_lib.fen_forward_train(C.byref(cfg), buf, buf, buf, 1, 64, 64, buf, 1 << 40, None) == _lib.FEN_ENODEV
    assert built_lib.fen_backward(C.byref(cfg), buf, buf, buf, buf, buf, 1, 64, 64, buf, 1 << 40, None) == _lib.FEN_ENODEV
    assert built_lib.fen_pack_weights_bwd(C.byref(cfg), buf, buf, None) == _lib.FEN_ENODEV
    assert built_lib.fen_packed_bwd_bytes(C.byref(cfg)) > 2 * 9 * 64 * 64 * 2      # layout queries still answer
    assert built_lib.fen_step_workspace_bytes(C.byref(cfg), 1, 64, 64) > 0
    with pytest.raises(RuntimeError):
        _lib.check(rc, "fen_forward")
