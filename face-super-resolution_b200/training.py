"""The reference's Stage-1 training step on the GPU (BASELINE config 5): nn.L1Loss
(src/losses/combined.py:38-47), clip_grad_norm_ and AdamW (src/training/trainer.py:217-221, 490-503), the
data-parallel gradient exchange, and `Stage1Step`, which strings them together with the float LR
generator and the network's forward / backward kernels (FaceEnhanceNet._forward_train / _backward) into one
iteration of Trainer._train_epoch (trainer.py:410-505).  Everything operates on flat fp32 tensors."""
from __future__ import annotations

from typing import Optional, Tuple

import torch
import torch.distributed as dist

from . import _lib


def _workspace(device: torch.device, n: int) -> torch.Tensor:
    lib = _lib.load()
    return torch.empty(int(lib.fen_train_workspace_bytes(n)), dtype=torch.uint8, device=device)


def _check_cuda_f32(*tensors: torch.Tensor) -> None:
    for t in tensors:
        if t.dtype != torch.float32:
            raise TypeError("float32 tensors expected")
        if not t.is_cuda:
            raise RuntimeError("the training kernels need CUDA tensors: there is no CPU fallback")
        if not t.is_contiguous():
            raise ValueError("contiguous tensors expected")


def l1_loss(sr: torch.Tensor, hr: torch.Tensor, want_grad: bool = True
            ) -> Tuple[torch.Tensor, Optional[torch.Tensor]]:
    """nn.L1Loss(reduction='mean')(sr, hr) and d loss / d sr = sign(sr - hr) / numel.  Returns (loss [1], dsr)."""
    _check_cuda_f32(sr, hr)
    if sr.shape != hr.shape:
        raise ValueError("sr and hr must have the same shape")
    lib = _lib.load()
    with torch.cuda.device(sr.device):
        loss = torch.empty(1, dtype=torch.float32, device=sr.device)
        dsr = torch.empty_like(sr) if want_grad else None
        ws = _workspace(sr.device, sr.numel())
        rc = lib.fen_l1_loss(sr.data_ptr(), hr.data_ptr(), sr.numel(), loss.data_ptr(),
                             dsr.data_ptr() if want_grad else None, ws.data_ptr(), ws.numel(),
                             torch.cuda.current_stream().cuda_stream)
        _lib.check(rc, "fen_l1_loss")
    return loss, dsr


def psnr(pred: torch.Tensor, target: torch.Tensor, data_range: float = 1.0) -> torch.Tensor:
    """Trainer._compute_psnr (trainer.py:621-628) / metrics.psnr (metrics.py:17-34) on the GPU:
    10 log10(data_range^2 / mean((pred - target)^2)) over the whole batch; returns a device scalar [1]."""
    _check_cuda_f32(pred, target)
    if pred.shape != target.shape:
        raise ValueError("pred and target must have the same shape")
    lib = _lib.load()
    with torch.cuda.device(pred.device):
        out = torch.empty(1, dtype=torch.float32, device=pred.device)
        ws = _workspace(pred.device, pred.numel())
        rc = lib.fen_psnr(pred.data_ptr(), target.data_ptr(), pred.numel(), float(data_range), out.data_ptr(),
                          ws.data_ptr(), ws.numel(), torch.cuda.current_stream().cuda_stream)
        _lib.check(rc, "fen_psnr")
    return out


EXCHANGE_DESCRIPTION = ("bucketed NCCL all-reduce (average) of the 20.5 MB fp32 gradient, 4 buckets in the order the "
                        "backward completes them, issued on a side stream while the next stages compute")


class BucketedAllReduce:
    """The exchange step of data-parallel Stage-1 training (the reference has no multi-GPU code).  The
    backward finishes the flat gradient slice by slice, last layers first (fen_backward_stages); slices are merged into
    `n_buckets` contiguous buckets and each bucket is averaged over the ranks as soon as it is complete - NCCL on a side
    stream that waits for an event recorded behind the bucket's last kernel, so that only the final bucket's
    all-reduce is exposed.  The gradient norm and AdamW then run on the averaged gradient, redundantly per rank."""

    def __init__(self, total: int, n_buckets: int = 4):
        self.total, self.n_buckets = int(total), max(1, int(n_buckets))
        self.stream = None
        self._lo = self._hi = None

    @staticmethod
    def active() -> bool:
        return dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1

    def _flush(self, grads: torch.Tensor) -> None:
        if self._lo is None:
            return
        piece = grads[self._lo:self._hi]
        self._lo = self._hi = None
        world = dist.get_world_size()
        if not piece.is_cuda:                      # gloo (CPU tests): no streams, no AVG
            dist.all_reduce(piece, op=dist.ReduceOp.SUM)
            piece.div_(world)
            return
        if self.stream is None:
            self.stream = torch.cuda.Stream(device=piece.device)
        done = torch.cuda.Event()
        done.record()                              # behind the last kernel that wrote this bucket
        with torch.cuda.stream(self.stream):
            self.stream.wait_event(done)
            dist.all_reduce(piece, op=dist.ReduceOp.AVG)

    def on_stage(self, grads: torch.Tensor, begin: int, count: int) -> None:
        """A slice [begin, begin + count) of the flat gradient is final (slices arrive in descending address order)."""
        if not self.active() or count == 0:       # (a stage may only feed a later batched weight-gradient launch)
            return
        if self._lo is not None and begin + count != self._lo:      # not adjacent to the pending bucket: send that first
            self._flush(grads)
        self._hi = begin + count if self._lo is None else self._hi
        self._lo = begin
        if self._hi - self._lo >= self.total // self.n_buckets or begin == 0 or self._small_tail(begin):
            self._flush(grads)

    def _small_tail(self, begin: int) -> bool:
        # what is still to come is small: send the pending bucket now so that only that small rest is exposed
        return begin <= self.total // (4 * self.n_buckets)

    def finish(self, grads: torch.Tensor) -> None:
        if not self.active():
            return
        self._flush(grads)
        if self.stream is not None:
            torch.cuda.current_stream().wait_stream(self.stream)


def allreduce_mean_(flat_grad: torch.Tensor) -> torch.Tensor:
    """Data-parallel gradient exchange: one all-reduce (sum) of the flat gradient, then / world_size, in place.
    NCCL over NVLink on GPUs; identity when torch.distributed is not initialised."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(flat_grad, op=dist.ReduceOp.SUM)
        flat_grad.div_(dist.get_world_size())
    return flat_grad


class ClipAdamW:
    """clip_grad_norm_(params, max_norm) + torch.optim.AdamW.step() on ONE flat fp32 parameter vector
    (trainer.py:217-221: AdamW(lr, weight_decay), default betas (0.9, 0.999), eps 1e-8; :490-496 clip 0.5)."""

    def __init__(self, flat_params: torch.Tensor, lr: float = 1e-4, betas=(0.9, 0.999), eps: float = 1e-8,
                 weight_decay: float = 0.0, max_norm: float = 0.5):
        _check_cuda_f32(flat_params)
        self.params = flat_params
        self.exp_avg = torch.zeros_like(flat_params)
        self.exp_avg_sq = torch.zeros_like(flat_params)
        self.lr, self.betas, self.eps, self.weight_decay, self.max_norm = lr, betas, eps, weight_decay, max_norm
        self.step_count = 0
        self.total_norm = torch.zeros(1, dtype=torch.float32, device=flat_params.device)
        self._ws = _workspace(flat_params.device, flat_params.numel())

    def step(self, flat_grad: torch.Tensor) -> torch.Tensor:
        """One optimiser step; returns the (device) total gradient norm before clipping."""
        _check_cuda_f32(flat_grad)
        if flat_grad.numel() != self.params.numel():
            raise ValueError("gradient size mismatch")
        lib = _lib.load()
        self.step_count += 1
        with torch.cuda.device(self.params.device):
            st = torch.cuda.current_stream().cuda_stream
            n = self.params.numel()
            _lib.check(lib.fen_grad_norm(flat_grad.data_ptr(), n, self.total_norm.data_ptr(), self._ws.data_ptr(),
                                         self._ws.numel(), st), "fen_grad_norm")
            _lib.check(lib.fen_clip_adamw_step(self.params.data_ptr(), flat_grad.data_ptr(), self.exp_avg.data_ptr(),
                                               self.exp_avg_sq.data_ptr(), n, self.total_norm.data_ptr(),
                                               self.max_norm, self.lr, self.betas[0], self.betas[1], self.eps,
                                               self.weight_decay, self.step_count, st), "fen_clip_adamw_step")
        return self.total_norm


class Stage1Step:
    """One iteration of Trainer._train_epoch for Stage 1 (trainer.py:410-505; L1 only, fp32 master weights,
    `mixed_precision: false` in configs/stage1_psnr_config.yaml:78):

        lr = F.interpolate(hr, scale_factor=0.25, mode='bicubic')      -> fen_lr_from_hr_f32
        sr = model(lr)                       (train mode, unclamped)    -> fen_forward_train
        loss = L1(sr, hr); loss.backward()                              -> fen_l1_loss, fen_backward
        [data parallel: all-reduce(mean) of the flat gradient, NCCL]    -> allreduce_mean_
        clip_grad_norm_(0.5); AdamW.step()                              -> fen_grad_norm, fen_clip_adamw_step

    The model's parameters are re-pointed at views of ONE flat fp32 vector (registration order = the order of the
    C ABI), so the optimiser updates them in place and `model.state_dict()` stays the reference's schema."""

    def __init__(self, model, lr: float = 1e-4, betas=(0.9, 0.999), eps: float = 1e-8, weight_decay: float = 0.0,
                 max_norm: float = 0.5):
        params = list(model.parameters())
        if not params or not params[0].is_cuda:
            raise RuntimeError("Stage1Step needs the model on a CUDA device: there is no CPU fallback")
        flat = torch.cat([p.detach().reshape(-1).to(torch.float32) for p in params]).contiguous()
        off = 0
        for p in params:
            n = p.numel()
            p.data = flat[off:off + n].view(p.shape)
            off += n
        self.model, self.flat = model, flat
        model._flat_master = flat      # fen_pack_weights reads it directly (no torch.cat per step)
        self.opt = ClipAdamW(flat, lr=lr, betas=betas, eps=eps, weight_decay=weight_decay, max_norm=max_norm)
        self.last_grad: Optional[torch.Tensor] = None
        self.exchange = True          # False: skip the gradient all-reduce (bench.py measures its exposed time that way)
        self.buckets = BucketedAllReduce(flat.numel())
        self.last_launches = 0        # kernels launched by the last step (counted by the library)

    def step(self, hr: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
        """hr: [B,3,4H,4W] fp32 CUDA in [0,1].  Returns (loss [1], total gradient norm before clipping [1]),
        both device scalars: nothing in the step synchronises with the host."""
        from .data import lr_from_hr_float
        _check_cuda_f32(hr)
        model = self.model
        model.train()
        lib = _lib.load()
        lr_img, _ = lr_from_hr_float(hr); n = lib.fen_last_launch_count()
        model._check_input(lr_img)
        sr, lease = model._forward_train(lr_img); n += lib.fen_last_launch_count()
        loss, dsr = l1_loss(sr, hr); n += lib.fen_last_launch_count()
        overlap = self.exchange and BucketedAllReduce.active()
        grads = model._backward(lr_img, dsr, lease, on_stage=self.buckets.on_stage if overlap else None)
        n += model._last_backward_launches if overlap else lib.fen_last_launch_count()
        del lease                      # the saved activations go back to the model's pool
        if overlap:
            self.buckets.finish(grads)
        norm = self.opt.step(grads); n += 3          # grad norm (2 kernels) + clip/AdamW
        self.last_launches = n
        model.mark_parameters_updated()
        self.last_grad = grads
        return loss, norm


# ===================================================================== SSIM (validation metric, Stage-2 loss term)
def _gaussian_1d(window_size: int, sigma: float) -> torch.Tensor:
    """The 1-D factor of create_gaussian_window (src/losses/ssim_loss.py:14-41), computed the same way in fp32."""
    coords = torch.arange(window_size, dtype=torch.float32)
    coords -= window_size // 2
    g = torch.exp(-(coords ** 2) / (2 * sigma ** 2))
    g /= g.sum()
    return g


def _ssim_call(pred, target, window_size, sigma, data_range, K, want_grad):
    import ctypes as C
    _check_cuda_f32(pred, target)
    if pred.shape != target.shape or pred.dim() != 4:
        raise ValueError("pred and target must be [B,C,H,W] tensors of the same shape")
    if window_size % 2 == 0 or window_size > 11:
        raise ValueError("window_size must be odd and at most 11 on the B200 path")
    lib = _lib.load()
    B, Cn, H, W = pred.shape
    g = _gaussian_1d(window_size, sigma)
    garr = (C.c_float * window_size)(*[float(v) for v in g])
    with torch.cuda.device(pred.device):
        per_image = torch.empty(B, dtype=torch.float32, device=pred.device)
        mean = torch.empty(1, dtype=torch.float32, device=pred.device)
        grad = torch.empty_like(pred) if want_grad else None
        nbytes = lib.fen_ssim_workspace_bytes(B, Cn, H, W, int(want_grad))
        _lib.check(nbytes, "fen_ssim_workspace_bytes")
        ws = torch.empty(int(nbytes), dtype=torch.uint8, device=pred.device)
        rc = lib.fen_ssim(pred.data_ptr(), target.data_ptr(), B, Cn, H, W, garr, window_size,
                          float((K[0] * data_range) ** 2), float((K[1] * data_range) ** 2), per_image.data_ptr(),
                          mean.data_ptr(), grad.data_ptr() if want_grad else None, ws.data_ptr(), ws.numel(),
                          torch.cuda.current_stream().cuda_stream)
        _lib.check(rc, "fen_ssim")
    return per_image, mean, grad


class _SsimFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, pred, target, window_size, sigma, data_range, size_average, K):
        want_grad = pred.requires_grad
        per_image, mean, grad = _ssim_call(pred.detach().contiguous(), target.detach().contiguous(), window_size, sigma,
                                           data_range, K, want_grad)
        ctx.grad_mean, ctx.size_average, ctx.B = grad, size_average, pred.shape[0]
        return mean[0] if size_average else per_image

    @staticmethod
    def backward(ctx, gout):
        if ctx.grad_mean is None:
            return (None,) * 7
        if ctx.size_average:
            g = ctx.grad_mean * gout
        else:                      # d mean_b / d pred = B * d mean / d pred on the pixels of image b
            g = ctx.grad_mean * (gout.view(-1, 1, 1, 1) * float(ctx.B))
        return (g, None, None, None, None, None, None)


def ssim(pred: torch.Tensor, target: torch.Tensor, window_size: int = 11, sigma: float = 1.5, data_range: float = 1.0,
         size_average: bool = True, K=(0.01, 0.03)) -> torch.Tensor:
    """src/losses/ssim_loss.py:44-98 on the GPU (same signature): the mean of the SSIM map (size_average) or its mean
    per image.  Differentiable w.r.t. `pred` (the target gets no gradient: it is the ground truth in every caller)."""
    return _SsimFunction.apply(pred, target, window_size, sigma, data_range, size_average, tuple(K))


class SSIMLoss(torch.nn.Module):
    """src/losses/ssim_loss.py:166-226: 1 - ssim(pred, target)."""

    def __init__(self, window_size: int = 11, sigma: float = 1.5, data_range: float = 1.0, size_average: bool = True,
                 channel: int = 3):
        super().__init__()
        self.window_size, self.sigma, self.data_range, self.size_average, self.channel = \
            window_size, sigma, data_range, size_average, channel
        self.register_buffer("window", (_gaussian_1d(window_size, sigma).unsqueeze(1) @ _gaussian_1d(window_size, sigma)
                                        .unsqueeze(0)).expand(channel, 1, window_size, window_size).contiguous())

    def forward(self, pred: torch.Tensor, target: torch.Tensor) -> torch.Tensor:
        return 1 - ssim(pred, target, window_size=self.window_size, sigma=self.sigma, data_range=self.data_range,
                        size_average=self.size_average)
