"""FaceEnhanceNet: drop-in for the reference's model API (src/models/custom.py), with the forward
pass executed by hand-written sm_100a kernels through the C ABI of include/fen_b200.h.

Same constructor (`FaceEnhanceNet(config=None, **kwargs)`), dataclass fields, attributes, sub-module
tree and therefore the same `state_dict()` keys / shapes (load_state_dict(strict=True) round-trips
with the reference).  forward(x: [B,3,H,W] fp32 CUDA in [0,1]) -> [B,3,4H,4W] fp32, clamped to
[0,1] iff not self.training (custom.py:187-188).

There is no CPU path and no cuDNN dispatch: CPU tensors, non-64-channel configs or a missing CUDA
library raise.  In train() mode with grad enabled, forward runs fen_forward_train (activations kept) and
returns a tensor whose grad_fn calls fen_backward: `loss.backward()` fills `.grad` of every parameter as
in the reference's Trainer._train_epoch (src/training/trainer.py:458-505).  The input gets no gradient.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import Any, Dict, Optional

import torch
import torch.nn as nn

from . import _lib
from .blocks import ResidualGroup, UpsampleModule


@dataclass
class FaceEnhanceNetConfig:
    """Same fields and defaults as the reference dataclass (custom.py:22-43)."""
    num_channels: int = 64
    num_groups: int = 3
    blocks_per_group: int = 4
    kernel_size: int = 3
    reduction_ratio: int = 4
    scale_factor: int = 4
    res_scale: float = 0.2
    in_channels: int = 3
    out_channels: int = 3
    init_scale: float = 0.1
    num_rcab_blocks: int = 8


class _StepLease:
    """One step workspace (the saved activations of ONE train-mode forward), checked out of the model's pool.  It
    goes back when the last reference dies - after the backward, or when the autograd graph that holds it is
    freed - so every outstanding forward owns its own activations (`y1 = m(x1); y2 = m(x2); (l1 + l2).backward()`
    works as with the reference's nn.Module)."""

    def __init__(self, pool, key, ws):
        self.pool, self.key, self.ws = pool, key, ws

    def __del__(self):
        try:
            if self.pool.get("key") == self.key:
                self.pool["free"].append(self.ws)
        except Exception:       # interpreter shutdown
            pass


class _FenTrainFunction(torch.autograd.Function):
    """sr = model(lr) in train() mode; backward = fen_backward (the network side of loss.backward())."""

    @staticmethod
    def forward(ctx, model, x, *params):
        out, lease = model._forward_train(x)
        ctx.model, ctx.x, ctx.lease = model, x, lease
        ctx.shapes = [p.shape for p in params]
        return out

    @staticmethod
    def backward(ctx, dout):
        flat = ctx.model._backward(ctx.x, dout, ctx.lease)
        grads, off = [], 0
        for shp in ctx.shapes:
            n = 1
            for d in shp:
                n *= d
            grads.append(flat[off:off + n].view(shp))
            off += n
        return (None, None, *grads)


class FaceEnhanceNet(nn.Module):
    def __init__(self, config: Optional[FaceEnhanceNetConfig] = None, **kwargs):
        super().__init__()
        if config is None:
            config = FaceEnhanceNetConfig()
        for key, value in kwargs.items():  # kwargs override dataclass fields (custom.py:78-80)
            if hasattr(config, key):
                setattr(config, key, value)
        self.config = config
        self.scale_factor = config.scale_factor
        self.num_channels = config.num_channels
        k, pad = config.kernel_size, config.kernel_size // 2
        self.conv_first = nn.Conv2d(config.in_channels, config.num_channels, k, padding=pad)
        self.residual_groups = nn.ModuleList([
            ResidualGroup(config.num_channels, config.blocks_per_group, k, config.reduction_ratio,
                          config.res_scale) for _ in range(config.num_groups)])
        self.conv_after_body = nn.Conv2d(config.num_channels, config.num_channels, k, padding=pad)
        self.upsample = UpsampleModule(config.num_channels, config.scale_factor)
        self.conv_last = nn.Conv2d(config.num_channels, config.out_channels, k, padding=pad)
        self._initialize_weights()
        # derived kernel-side state (never part of state_dict)
        self._packed: Optional[torch.Tensor] = None
        self._packed_key = None
        self._packed_bwd: Optional[torch.Tensor] = None
        self._packed_bwd_gen = -1
        self._pack_gen = 0            # bumped whenever the packed forward weights are rebuilt
        self._raw_update_gen = 0      # bumped by mark_parameters_updated (raw-pointer optimiser updates)
        self._workspaces: Dict[Any, torch.Tensor] = {}
        self._step_pool: Dict[str, Any] = {"key": None, "free": []}

    # ------------------------------------------------------------------ init (custom.py:128-145)
    def _initialize_weights(self) -> None:
        for m in self.modules():
            if isinstance(m, (nn.Conv2d, nn.Linear)):
                nn.init.kaiming_normal_(m.weight, mode="fan_out", nonlinearity="relu")
                if m.bias is not None:
                    nn.init.zeros_(m.bias)
        nn.init.zeros_(self.conv_last.weight)
        nn.init.zeros_(self.conv_last.bias)

    # ------------------------------------------------------------------ kernel-side plumbing
    def _c_config(self) -> _lib.FenConfig:
        """The kernel-side configuration.  The kernels are built for 64 feature channels; a narrower model
        (FaceEnhanceNetLite: 32) runs on them EMBEDDED in 64 channels - its weights sit in the top-left corner of
        zero-padded 64-channel tensors (_embedding), which computes exactly the same network (the padded channels
        stay identically zero through every layer) at the price of the unused lanes."""
        c = self.config
        if (c.kernel_size != 3 or c.in_channels != 3 or c.out_channels != 3):
            raise ValueError("unsupported config for the B200 path: kernel_size must be 3 and "
                             "in_channels == out_channels == 3 (no fallback path)")
        if c.num_channels == 64:
            return _lib.FenConfig(64, c.num_groups, c.blocks_per_group, c.reduction_ratio, c.scale_factor,
                                  float(c.res_scale))
        hidden = max(c.num_channels // c.reduction_ratio, 8)           # blocks.py:62
        if not (1 <= c.num_channels < 64) or 64 % hidden:
            raise ValueError("unsupported config for the B200 path: num_channels must be 64, or less than 64 with an "
                             "SE hidden width that divides 64 (no fallback path)")
        return _lib.FenConfig(64, c.num_groups, c.blocks_per_group, 64 // hidden, c.scale_factor, float(c.res_scale))

    def _embedding(self, device: torch.device) -> Optional[torch.Tensor]:
        """None for a 64-channel model; otherwise, for every element of the model's flat parameter vector, its index
        in the flat vector of the 64-channel model the kernels run (int64 tensor on `device`)."""
        if self.config.num_channels == 64:
            return None
        emb = getattr(self, "_embed_index", None)
        if emb is not None and emb.device == device:
            return emb
        c, cc = self.config, self._c_config()
        with torch.device("meta"):
            wide = FaceEnhanceNet(FaceEnhanceNetConfig(
                num_channels=64, num_groups=c.num_groups, blocks_per_group=c.blocks_per_group,
                reduction_ratio=cc.reduction_ratio, scale_factor=c.scale_factor, res_scale=c.res_scale))
        parts, off = [], 0
        for (name, p), (wname, wp) in zip(self.named_parameters(), wide.named_parameters()):
            assert name == wname and p.dim() == wp.dim(), (name, wname)
            idx = torch.arange(wp.numel(), dtype=torch.int64).view(wp.shape)
            idx = idx[tuple(slice(0, d) for d in p.shape)]      # top-left corner (PixelShuffle rows 4c + sub stay in place)
            parts.append(idx.reshape(-1) + off)
            off += wp.numel()
        self._embed_index = torch.cat(parts).to(device)
        self._embed_total = off
        return self._embed_index

    def _params_key(self):
        return tuple((p.data_ptr(), p._version) for p in self.parameters())

    def _ensure_packed(self, device: torch.device, stream_ptr: int) -> torch.Tensor:
        # torch's version counters see optimiser.step() and load_state_dict(); the fused optimiser kernel writes
        # through raw pointers, which only mark_parameters_updated() reports: both are part of the key
        key = (str(device), self._raw_update_gen, self._params_key())
        if self._packed is not None and self._packed_key == key:
            return self._packed
        lib, cfg = _lib.load(), self._c_config()
        nbytes = lib.fen_packed_bytes(C.byref(cfg))
        _lib.check(nbytes, "fen_packed_bytes")
        flat = getattr(self, "_flat_master", None)       # training.Stage1Step: the parameters are views of this vector
        if flat is None or flat.device != device:
            flat = torch.cat([p.detach().reshape(-1).to(torch.float32) for p in self.parameters()])
        emb = self._embedding(device)
        if emb is not None:                               # narrow model embedded in 64 channels
            wide = torch.zeros(self._embed_total, dtype=torch.float32, device=device)
            wide[emb] = flat
            flat = wide
        n_expected = lib.fen_param_count(C.byref(cfg))
        if flat.numel() != n_expected:
            raise RuntimeError(f"parameter count {flat.numel()} != kernel layout {n_expected}")
        packed = torch.empty(nbytes, dtype=torch.uint8, device=device)
        _lib.check(lib.fen_pack_weights(C.byref(cfg), flat.data_ptr(), packed.data_ptr(), stream_ptr),
                   "fen_pack_weights")
        self._packed, self._packed_key = packed, key
        self._pack_gen += 1           # every rebuild of the forward weights invalidates the transposed copies
        self._flat = flat
        return packed

    def mark_parameters_updated(self) -> None:
        """Tell the module that its parameters were updated in place through raw pointers (the fused optimiser
        kernel of training.Stage1Step does not bump torch's version counters): both packed copies (forward and
        transposed) are rebuilt before their next use."""
        self._raw_update_gen += 1
        self._packed_key = None

    def _ensure_packed_bwd(self, device: torch.device, stream_ptr: int) -> torch.Tensor:
        """Transposed / tap-flipped weights for the data-gradient convolutions (fen_pack_weights_bwd).  Keyed on the
        generation counter of the forward copy (a monotonically increasing integer, never a value-comparable tuple)."""
        self._ensure_packed(device, stream_ptr)
        if self._packed_bwd is not None and self._packed_bwd_gen == self._pack_gen \
                and self._packed_bwd.device == device:
            return self._packed_bwd
        lib, cfg = _lib.load(), self._c_config()
        nbytes = lib.fen_packed_bwd_bytes(C.byref(cfg))
        _lib.check(nbytes, "fen_packed_bwd_bytes")
        packed = self._packed_bwd
        if packed is None or packed.numel() != nbytes or packed.device != device:
            packed = torch.empty(nbytes, dtype=torch.uint8, device=device)
        _lib.check(lib.fen_pack_weights_bwd(C.byref(cfg), self._flat.data_ptr(), packed.data_ptr(), stream_ptr),
                   "fen_pack_weights_bwd")
        self._packed_bwd, self._packed_bwd_gen = packed, self._pack_gen
        return packed

    def _check_input(self, x: torch.Tensor) -> None:
        if not isinstance(x, torch.Tensor) or x.dim() != 4 or x.shape[1] != 3:
            raise RuntimeError(f"expected input [B,3,H,W], got {tuple(getattr(x, 'shape', ()))}")
        if not x.is_cuda:
            raise RuntimeError("FaceEnhanceNet (B200 path) needs a CUDA tensor: there is no CPU fallback")
        if next(self.parameters()).device != x.device:
            raise RuntimeError("input and parameters are on different devices")

    def _forward_train(self, x: torch.Tensor):
        """fen_forward_train: unclamped output + a lease on the step workspace holding the saved activations."""
        lib, cfg = _lib.load(), self._c_config()
        B, _, H, W = x.shape
        with torch.cuda.device(x.device):
            stream = torch.cuda.current_stream().cuda_stream
            packed = self._ensure_packed(x.device, stream)
            skey = (str(x.device), B, H, W)
            pool = self._step_pool
            if pool["key"] != skey:              # another shape: workspaces of the old one are dropped as they come back
                pool["key"], pool["free"] = skey, []
            if pool["free"]:
                ws = pool["free"].pop()
            else:
                nbytes = lib.fen_step_workspace_bytes(C.byref(cfg), B, H, W)
                _lib.check(nbytes, "fen_step_workspace_bytes")
                ws = torch.empty(nbytes, dtype=torch.uint8, device=x.device)
            lease = _StepLease(pool, skey, ws)
            s = self.scale_factor
            out = torch.empty((B, 3, H * s, W * s), dtype=torch.float32, device=x.device)
            _lib.check(lib.fen_forward_train(C.byref(cfg), packed.data_ptr(), x.data_ptr(), out.data_ptr(), B, H, W,
                                             ws.data_ptr(), ws.numel(), stream), "fen_forward_train")
        return out, lease

    def _backward(self, x: torch.Tensor, dout: torch.Tensor, lease: "_StepLease", on_stage=None) -> torch.Tensor:
        """fen_backward: flat fp32 gradient in parameter order for the _forward_train that produced `lease`.
        on_stage(grads, begin, count), if given, is called after every stage of the backward with the slice of the flat
        gradient that stage has just completed (fen_backward_stages; the data-parallel exchange hooks in here)."""
        lib, cfg = _lib.load(), self._c_config()
        ws = lease.ws
        B, _, H, W = x.shape
        dout = dout.detach().to(torch.float32).contiguous()
        with torch.cuda.device(x.device):
            stream = torch.cuda.current_stream().cuda_stream
            packed = self._ensure_packed(x.device, stream)
            packed_bwd = self._ensure_packed_bwd(x.device, stream)
            grads = torch.empty(self._flat.numel(), dtype=torch.float32, device=x.device)
            emb = self._embedding(x.device)
            if on_stage is None or emb is not None:
                _lib.check(lib.fen_backward(C.byref(cfg), packed.data_ptr(), packed_bwd.data_ptr(), x.data_ptr(),
                                            dout.data_ptr(), grads.data_ptr(), B, H, W, ws.data_ptr(), ws.numel(),
                                            stream), "fen_backward")
                if emb is not None:
                    grads = grads[emb]
                if on_stage is not None:
                    on_stage(grads, 0, grads.numel())
            else:
                launches = 0
                begin, count = C.c_int64(), C.c_int64()
                for stage in range(lib.fen_backward_num_stages(C.byref(cfg))):
                    _lib.check(lib.fen_backward_stages(C.byref(cfg), packed.data_ptr(), packed_bwd.data_ptr(),
                                                       x.data_ptr(), dout.data_ptr(), grads.data_ptr(), B, H, W,
                                                       ws.data_ptr(), ws.numel(), stage, stage + 1, stream),
                               "fen_backward_stages")
                    launches += lib.fen_last_launch_count()
                    _lib.check(lib.fen_backward_stage_range(C.byref(cfg), stage, C.byref(begin), C.byref(count)),
                               "fen_backward_stage_range")
                    on_stage(grads, begin.value, count.value)
                self._last_backward_launches = launches
        return grads

    def _run(self, x: torch.Tensor, want_se: bool, u8: Optional[bool] = None):
        self._check_input(x)
        lib, cfg = _lib.load(), self._c_config()
        x = x.detach().to(torch.float32).contiguous()
        if (not want_se and u8 is None and self.training and torch.is_grad_enabled()
                and any(p.requires_grad for p in self.parameters())):
            self._c_config()
            return _FenTrainFunction.apply(self, x, *self.parameters()), None
        B, _, H, W = x.shape
        with torch.cuda.device(x.device):
            stream = torch.cuda.current_stream().cuda_stream
            packed = self._ensure_packed(x.device, stream)
            wkey = (str(x.device), B, H, W)
            ws = self._workspaces.get(wkey)
            if ws is None:
                nbytes = lib.fen_forward_workspace_bytes(C.byref(cfg), B, H, W)
                _lib.check(nbytes, "fen_forward_workspace_bytes")
                self._workspaces.clear()
                ws = torch.empty(nbytes, dtype=torch.uint8, device=x.device)
                self._workspaces[wkey] = ws
            s = self.scale_factor
            if u8 is not None:          # forward_u8: uint8 HWC straight from the conv_last epilogue
                out = torch.empty((B, H * s, W * s, 3), dtype=torch.uint8, device=x.device)
                rc = lib.fen_forward_u8(C.byref(cfg), packed.data_ptr(), x.data_ptr(), out.data_ptr(), int(u8), B, H, W,
                                        ws.data_ptr(), ws.numel(), stream)
                _lib.check(rc, "fen_forward_u8")
                return out, None
            out = torch.empty((B, 3, H * s, W * s), dtype=torch.float32, device=x.device)
            n_rcab = self.config.num_groups * self.config.blocks_per_group
            se = torch.empty((B, n_rcab, 64), dtype=torch.float32, device=x.device) if want_se else None
            rc = lib.fen_forward(C.byref(cfg), packed.data_ptr(), x.data_ptr(), out.data_ptr(), B, H, W,
                                 1 if self.training else 0, ws.data_ptr(), ws.numel(),
                                 se.data_ptr() if se is not None else None, stream)
            _lib.check(rc, "fen_forward")
        if se is not None and self.config.num_channels != 64:
            se = se[:, :, :self.config.num_channels]
        return out, se

    # ------------------------------------------------------------------ reference API
    def forward(self, x: torch.Tensor) -> torch.Tensor:
        """custom.py:147-190."""
        return self._run(x, want_se=False)[0]

    def forward_u8(self, x: torch.Tensor, bgr: bool = False) -> torch.Tensor:
        """What the reference's evaluation scripts do with the output (scripts/test_model.py:176-190:
        `np.clip(sr * 255, 0, 255).astype(np.uint8)`, CHW -> HWC, optionally RGB -> BGR for cv2), fused into the last
        kernel: [B,3,H,W] fp32 -> [B,4H,4W,3] uint8, eval semantics (clamped), no autograd.  Bit-identical to
        data.sr_to_uint8(model.eval()(x))."""
        with torch.no_grad():
            return self._run(x, want_se=False, u8=bool(bgr))[0]

    def get_attention_maps(self, x: torch.Tensor) -> Dict[str, torch.Tensor]:
        """custom.py:192-230: {'group{g}_rcab{b}': [B, C] sigmoid channel weights}.  Served natively
        from the fused path (sub-module hooks do not fire there)."""
        with torch.no_grad():
            _, se = self._run(x, want_se=True)
        maps = {}
        for g in range(self.config.num_groups):
            for b in range(self.config.blocks_per_group):
                maps[f"group{g}_rcab{b}"] = se[:, g * self.config.blocks_per_group + b]
        return maps

    def feature_tap(self, x_shape, which: int, index: int = 0) -> torch.Tensor:
        """Parity/debug helper: NHWC bf16 intermediate of the last forward with input shape x_shape."""
        lib, cfg = _lib.load(), self._c_config()
        B, _, H, W = x_shape
        ws = next(iter(self._workspaces.values()))
        ptr = C.c_void_p()
        n = lib.fen_forward_tap(C.byref(cfg), ws.data_ptr(), B, H, W, which, index, C.byref(ptr))
        _lib.check(n, "fen_forward_tap")
        off = ptr.value - ws.data_ptr()
        mult = {0: 1, 1: 1, 2: 2, 3: 4, 4: 1}[which]
        return ws[off:off + n].view(torch.bfloat16).view(B, H * mult, W * mult, 64)[..., :self.config.num_channels]

    def get_model_info(self) -> Dict[str, Any]:
        """custom.py:232-256 (same keys)."""
        total = sum(p.numel() for p in self.parameters())
        trainable = sum(p.numel() for p in self.parameters() if p.requires_grad)
        c = self.config
        return {
            "name": "FaceEnhanceNet", "total_params": total, "trainable_params": trainable,
            "size_mb": total * 4 / (1024 ** 2), "num_groups": c.num_groups,
            "blocks_per_group": c.blocks_per_group, "total_rcab_blocks": c.num_groups * c.blocks_per_group,
            "num_channels": c.num_channels, "scale_factor": self.scale_factor,
            "input_size": "64x64", "output_size": f"{64 * self.scale_factor}x{64 * self.scale_factor}",
        }

    @classmethod
    def from_pretrained(cls, checkpoint_path: str, device: Optional[str] = None) -> "FaceEnhanceNet":
        """custom.py:258-292: accepts {'config', 'model_state_dict' | 'state_dict'} or a bare state_dict."""
        ckpt = torch.load(checkpoint_path, map_location="cpu")
        config = FaceEnhanceNetConfig(**ckpt["config"]) if isinstance(ckpt, dict) and "config" in ckpt \
            else FaceEnhanceNetConfig()
        model = cls(config)
        if "model_state_dict" in ckpt:
            model.load_state_dict(ckpt["model_state_dict"])
        elif "state_dict" in ckpt:
            model.load_state_dict(ckpt["state_dict"])
        else:
            model.load_state_dict(ckpt)
        return model.to(device) if device else model


def create_face_enhance_net(num_rcab_blocks: int = 8, num_channels: int = 64, scale_factor: int = 4,
                            **kwargs) -> FaceEnhanceNet:
    """custom.py:295-319."""
    return FaceEnhanceNet(FaceEnhanceNetConfig(num_rcab_blocks=num_rcab_blocks, num_channels=num_channels,
                                               scale_factor=scale_factor, **kwargs))


class FaceEnhanceNetLite(FaceEnhanceNet):
    """custom.py:323-333 (32 channels, 4 RCABs... as the reference builds it).  Runs on the 64-channel kernels with
    its weights embedded in zero-padded 64-channel tensors (FaceEnhanceNet._c_config): same results as a 32-channel
    implementation, half of the tensor-core lanes idle."""

    def __init__(self, **kwargs):
        super().__init__(FaceEnhanceNetConfig(num_channels=32, num_rcab_blocks=4, reduction_ratio=2, **kwargs))
