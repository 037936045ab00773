"""Generates the committed golden fixtures.  Run HERE (the authoring container), where
/root/reference and cv2 exist:   python tests/golden/make_golden.py

  lr_golden.npz   outputs of cv2.resize(hr, (W/4, H/4), interpolation=cv2.INTER_CUBIC) - the call the
                  reference makes at src/data/dataset.py:296 / prepare_data.py:38 - for the inputs
                  of cases.LR_CASES (cv2 version recorded in the file).
  fen_golden.npz  outputs of the UNMODIFIED reference module `src.models.FaceEnhanceNet` (imported
                  from /root/reference) with weights from oracle/weights.py loaded by
                  load_state_dict(strict=True) in train() mode (unclamped; the eval() output is exactly its clamp to
                  [0,1], asserted at generation time), plus its get_attention_maps.
  fen_grad_golden.npz  .grad of every parameter of the same unmodified module after
                  `m.train(); m(x).backward(dout)` for cases.GRAD_CASES (dout = cases.grad_dout): what
                  loss.backward() leaves behind in Trainer._train_epoch (src/training/trainer.py:462-488).
Nothing at test time reads /root/reference."""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, HERE)
import cases  # noqa: E402
from oracle import weights  # noqa: E402


def make_lr():
    import cv2
    out = {"cv2_version": np.array(cv2.__version__)}
    for name, (H, W, C), _ in cases.LR_CASES:
        hr = cases.lr_input(name)
        src = hr[:, :, 0] if C == 1 else hr
        lr = cv2.resize(src, (W // 4, H // 4), interpolation=cv2.INTER_CUBIC)
        out[name] = lr.reshape(H // 4, W // 4, C)
    np.savez_compressed(os.path.join(HERE, "lr_golden.npz"), **out)
    print("lr_golden.npz:", {k: v.shape for k, v in out.items() if k != "cv2_version"})


def make_fen():
    sys.path.insert(0, "/root/reference")
    from src.models import FaceEnhanceNet
    torch.set_num_threads(8)
    out = {"torch_version": np.array(torch.__version__)}
    for name, cfg, tier, seed, batch in cases.FEN_CASES:
        sd = weights.make_state_dict(seed, tier, **cfg)
        m = FaceEnhanceNet(num_channels=64, scale_factor=4, **cfg)
        m.load_state_dict(sd, strict=True)
        x = torch.from_numpy(cases.fen_input(name))
        with torch.no_grad():
            m.eval()
            y_eval = m(x)
            att = m.get_attention_maps(x)
            m.train()
            y_train = m(x)
        assert torch.equal(y_eval, y_train.clamp(0.0, 1.0))  # eval output is exactly clamp(train output)
        out[name + "/train"] = y_train.numpy().astype(np.float32)
        out[name + "/se"] = np.stack([att[f"group{g}_rcab{b}"].numpy() for g in range(cfg["num_groups"])
                                      for b in range(cfg["blocks_per_group"])], 1)
        print(name, y_eval.shape, float(y_eval.min()), float(y_eval.max()), float(y_train.min()), float(y_train.max()))
    np.savez_compressed(os.path.join(HERE, "fen_golden.npz"), **out)


def make_fen_grad():
    sys.path.insert(0, "/root/reference")
    from src.models import FaceEnhanceNet
    torch.set_num_threads(8)
    out = {"torch_version": np.array(torch.__version__)}
    for name in cases.GRAD_CASES:
        _, cfg, tier, seed, _ = [c for c in cases.FEN_CASES if c[0] == name][0]
        m = FaceEnhanceNet(num_channels=64, scale_factor=4, **cfg)
        m.load_state_dict(weights.make_state_dict(seed, tier, **cfg), strict=True)
        m.train()
        m(torch.from_numpy(cases.fen_input(name))).backward(torch.from_numpy(cases.grad_dout(name)))
        for k, p in m.named_parameters():
            out[name + "/" + k] = p.grad.numpy().astype(np.float32)
        print(name, "grad norm", float(torch.sqrt(sum((p.grad.double() ** 2).sum() for p in m.parameters()))))
    np.savez_compressed(os.path.join(HERE, "fen_grad_golden.npz"), **out)


def make_fen2():
    """fen_golden2.npz: the unmodified reference on configurations beyond the benchmark's (default 3 x 4 config,
    FaceEnhanceNetLite, a ragged input size) and its gradients with NEGATIVE PReLU slopes."""
    sys.path.insert(0, "/root/reference")
    from src.models.custom import FaceEnhanceNet, FaceEnhanceNetLite
    torch.set_num_threads(8)
    out = {"torch_version": np.array(torch.__version__)}
    for name, ctor, cfg, tier, seed, shape in cases.FEN2_CASES:
        sd = weights.make_state_dict(seed, tier, **cfg)
        m = FaceEnhanceNetLite() if ctor == "lite" else FaceEnhanceNet(**cfg)
        m.load_state_dict(sd, strict=True)
        x = torch.from_numpy(cases.fen2_input(name))
        with torch.no_grad():
            m.train()
            y = m(x)
        out[name + "/train"] = y.numpy().astype(np.float32)
        print(name, tuple(y.shape), float(y.min()), float(y.max()))
    sd = cases.negative_slopes(weights.make_state_dict(cases.NEG_GRAD_SEED, "T1", **cases.NEG_GRAD_CFG), cases.NEG_GRAD_SEED)
    m = FaceEnhanceNet(**cases.NEG_GRAD_CFG)
    m.load_state_dict(sd, strict=True)
    m.train()
    x, dout = cases.neg_grad_inputs()
    y = m(torch.from_numpy(x))
    y.backward(torch.from_numpy(dout))
    out["neg/train"] = y.detach().numpy().astype(np.float32)
    for k, p in m.named_parameters():
        if cases.neg_grad_stored(k, p.numel()):
            out["neg/grad/" + k] = p.grad.numpy().astype(np.float32)
    n_neg = sum(int((v < 0).sum()) for k, v in sd.items() if k.endswith("prelu.weight"))
    print("negative-slope gradient case:", n_neg, "negative slopes")
    np.savez_compressed(os.path.join(HERE, "fen_golden2.npz"), **out)


def make_ssim():
    """ssim_golden.npz: src.losses.ssim_loss.ssim of the unmodified reference (mean, per-image means) and the gradient
    of SSIMLoss w.r.t. the prediction (small cases only)."""
    sys.path.insert(0, "/root/reference")
    from src.losses.ssim_loss import SSIMLoss, ssim
    out = {}
    for name, shape, ws, sigma in cases.SSIM_CASES:
        pred, target = (torch.from_numpy(a) for a in cases.ssim_inputs(name))
        out[name + "/mean"] = ssim(pred, target, window_size=ws, sigma=sigma).numpy()
        out[name + "/per_image"] = ssim(pred, target, window_size=ws, sigma=sigma, size_average=False).numpy()
        if pred.numel() < 50000:
            p = pred.clone().requires_grad_(True)
            SSIMLoss(window_size=ws, sigma=sigma, channel=shape[1])(p, target).backward()
            out[name + "/loss_grad"] = p.grad.numpy()
        print("ssim", name, float(out[name + "/mean"]))
    np.savez_compressed(os.path.join(HERE, "ssim_golden.npz"), **out)


if __name__ == "__main__":
    if "--ssim-only" in sys.argv:
        make_ssim()
        sys.exit(0)
    if "--round2-only" in sys.argv:
        make_fen2()
        print("fen_golden2.npz", os.path.getsize(os.path.join(HERE, "fen_golden2.npz")) // 1024, "KiB")
        sys.exit(0)
    if "--grad-only" in sys.argv:
        make_fen_grad()
        sys.exit(0)
    make_lr()
    make_fen()
    make_fen_grad()
    make_fen2()
    make_ssim()
    for f in ("lr_golden.npz", "fen_golden.npz", "fen_grad_golden.npz", "fen_golden2.npz"):
        print(f, os.path.getsize(os.path.join(HERE, f)) // 1024, "KiB")
