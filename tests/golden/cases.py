"""Deterministic input generators shared by make_golden.py (run once, here, against the real
reference / cv2) and by the tests (which regenerate the same inputs anywhere).  numpy PCG64 only."""
import numpy as np

LR_CASES = [
    # name, (H, W, C), kind
    ("random_256", (256, 256, 3), "random"),
    ("binary_256", (256, 256, 3), "binary"),
    ("ties_256", (256, 256, 3), "ties"),
    ("smooth_256", (256, 256, 3), "smooth"),
    ("blocks_256", (256, 256, 3), "blocks"),
    ("random_128", (128, 128, 3), "random"),
    ("random_512_gray", (512, 512, 1), "random"),
    ("random_64x192_rgba", (64, 192, 4), "random"),
    ("extremes_256", (256, 256, 3), "extremes"),
]


def lr_input(name: str) -> np.ndarray:
    idx = [c[0] for c in LR_CASES].index(name)
    _, (H, W, C), kind = LR_CASES[idx]
    rng = np.random.default_rng(1000 + idx)
    if kind == "random":
        a = rng.integers(0, 256, (H, W, C))
    elif kind == "binary":
        a = rng.integers(0, 2, (H, W, C)) * 255
    elif kind == "ties":       # multiples of 32 make u/1024 land on exact .5 ties often
        a = rng.integers(0, 8, (H, W, C)) * 32
    elif kind == "smooth":
        yy, xx = np.mgrid[0:H, 0:W]
        a = np.stack([(yy * 3 + xx * 2 + 40 * c) % 256 for c in range(C)], -1)
    elif kind == "blocks":     # constant 4x4 blocks: output must equal the block value
        a = np.repeat(np.repeat(rng.integers(0, 256, (H // 4, W // 4, C)), 4, 0), 4, 1)
    elif kind == "extremes":   # saturating patterns (0/255 checkerboards of varying period)
        yy, xx = np.mgrid[0:H, 0:W]
        a = np.stack([(((yy // (c + 1)) + (xx // (c + 2))) % 2) * 255 for c in range(C)], -1)
    else:
        raise ValueError(kind)
    return np.ascontiguousarray(a.astype(np.uint8))


FEN_CASES = [
    # name, config, tier, seed, batch
    ("small_T1", dict(num_groups=1, blocks_per_group=2), "T1", 0, 2),
    ("small_T0", dict(num_groups=2, blocks_per_group=1), "T0", 1, 1),
    ("full_T1", dict(num_groups=6, blocks_per_group=10), "T1", 0, 1),
]


def fen_input(name: str) -> np.ndarray:
    idx = [c[0] for c in FEN_CASES].index(name)
    batch = FEN_CASES[idx][4]
    rng = np.random.default_rng(2000 + idx)
    return rng.random((batch, 3, 64, 64), dtype=np.float32)


GRAD_CASES = ["small_T1"]   # FEN_CASES entries that also have a gradient golden (fen_grad_golden.npz)


def grad_dout(name: str) -> np.ndarray:
    """d loss / d sr handed to the network in the gradient cases: the structure nn.L1Loss gives it
    (+-1 / numel, src/losses/combined.py:38-47) with a fixed random sign pattern."""
    idx = [c[0] for c in FEN_CASES].index(name)
    batch = FEN_CASES[idx][4]
    rng = np.random.default_rng(3000 + idx)
    shape = (batch, 3, 256, 256)
    sign = rng.integers(0, 2, shape).astype(np.float32) * 2.0 - 1.0
    return (sign / float(np.prod(shape))).astype(np.float32)


# ---- round 2: configurations beyond the 64-channel square benchmark, fen_golden2.npz
FEN2_CASES = [
    # name, reference constructor ("net" = FaceEnhanceNet(**cfg), "lite" = FaceEnhanceNetLite()), weight cfg, tier, seed, input shape
    ("default_T1", "net", dict(num_groups=3, blocks_per_group=4), "T1", 11, (1, 3, 64, 64)),        # dataclass defaults (custom.py:28-29)
    ("lite_T1", "lite", dict(num_groups=3, blocks_per_group=4, num_channels=32, reduction_ratio=2), "T1", 12, (1, 3, 64, 64)),
    ("ragged_T1", "net", dict(num_groups=1, blocks_per_group=2), "T1", 13, (1, 3, 48, 80)),          # neither side a multiple of 64
]


def fen2_input(name: str) -> np.ndarray:
    idx = [c[0] for c in FEN2_CASES].index(name)
    rng = np.random.default_rng(4000 + idx)
    return rng.random(FEN2_CASES[idx][5], dtype=np.float32)


def negative_slopes(sd, seed: int = 0):
    """The same state_dict with every PReLU slope redrawn from U(-0.4, 0.45): about half of them negative, which
    nn.PReLU allows (blocks.py:127,146,216) and a trained checkpoint may contain."""
    import torch
    rng = np.random.default_rng(5000 + seed)
    out = {k: v.clone() for k, v in sd.items()}
    for k in out:
        if k.endswith("prelu.weight"):
            out[k] = torch.from_numpy(rng.uniform(-0.4, 0.45, tuple(out[k].shape)).astype(np.float32))
    return out


# gradient golden with negative slopes: 1 x 1 model, batch 1, dout = grad_dout-like sign pattern; only the tensors
# named here are stored (every PReLU slope gradient, the SE matrices, biases, and conv_first.weight, which sees
# every layer above it)
NEG_GRAD_CFG = dict(num_groups=1, blocks_per_group=1)
NEG_GRAD_SEED = 21


def neg_grad_inputs():
    rng = np.random.default_rng(6000)
    x = rng.random((1, 3, 64, 64), dtype=np.float32)
    sign = rng.integers(0, 2, (1, 3, 256, 256)).astype(np.float32) * 2.0 - 1.0
    return x, (sign / float(sign.size)).astype(np.float32)


def neg_grad_stored(key: str, numel: int) -> bool:
    return numel <= 5000 or key == "conv_first.weight"


# ---- SSIM (src/losses/ssim_loss.py): (name, shape, window_size, sigma); pred = target + noise, clipped to [0, 1]
SSIM_CASES = [("ragged", (2, 3, 40, 52), 11, 1.5), ("sr_size", (2, 3, 256, 256), 11, 1.5), ("win7", (1, 1, 33, 64), 7, 1.0)]


def ssim_inputs(name: str):
    idx = [c[0] for c in SSIM_CASES].index(name)
    shape = SSIM_CASES[idx][1]
    rng = np.random.default_rng(7000 + idx)
    yy, xx = np.mgrid[0:shape[2], 0:shape[3]].astype(np.float32)
    base = 0.5 + 0.3 * np.sin(yy / 5.0)[None, None] * np.cos(xx / 7.0)[None, None]
    target = np.clip(base + 0.15 * rng.standard_normal(shape), 0, 1).astype(np.float32)
    pred = np.clip(target + 0.08 * rng.standard_normal(shape), 0, 1).astype(np.float32)
    return pred, target
