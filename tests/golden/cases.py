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
