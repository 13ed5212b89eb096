"""Device-side input path of the step (csrc/msp_input.cu).

The reference prepares every batch on the CPU — `np.load(f) / 255` (classification/datasets.py:47), the float32 cast of
`ConvertToType` (transform/transforms.py:63-103), `RepeatChannels` (transform/transforms.py:134-142) — and moves the
fp32 result with a synchronous pageable `.to(device)` (train_model.py:60).  Here the raw uint8 bytes cross PCIe (4x to
12x fewer than the prepared fp32 batch) and one kernel produces the model's fp32 NCHW input on the device.
`ColorJitter` is the augmentation of the robustness evaluation (robustness/eval.py:61-66), applied to the whole batch
on the device with torchvision's own parameter draw.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Sequence, Tuple

import numpy as np
import torch

from . import _lib


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def u8_to_float(x: torch.Tensor, repeats: int = 1, divisor: float = 255.0, out: Optional[torch.Tensor] = None):
    """uint8 CUDA (N, C, *spatial) -> fp32 (N, C * repeats, *spatial): `x / divisor` in float64 rounded to float32
    (exactly numpy's `np.load(f) / 255` followed by the float32 cast), every channel repeated like np.repeat(axis=0)
    on a CHW image."""
    if not x.is_cuda or x.dtype != torch.uint8:
        raise RuntimeError("u8_to_float: expected a CUDA uint8 tensor (the B200 path has no CPU fallback)")
    x = x.contiguous()
    n, c = x.shape[0], x.shape[1]
    hw = int(np.prod(x.shape[2:])) if x.dim() > 2 else 1
    shape = (n, c * repeats, *x.shape[2:])
    if out is None:
        out = torch.empty(shape, dtype=torch.float32, device=x.device)
    elif tuple(out.shape) != shape or out.dtype != torch.float32 or not out.is_contiguous():
        raise ValueError(f"u8_to_float: `out` must be a contiguous fp32 tensor of shape {shape}")
    _lib.call("msp_u8_to_f32_nchw", x.data_ptr(), n, c, hw, int(repeats), float(divisor), out.data_ptr(), _stream())
    return out


class DeviceInput:
    """`RepeatChannels(repeats)` + `/255` + float32 cast as one device kernel behind a callable: feed it the uint8
    batch that `host.BatchPrefetcher` copied, get the model's fp32 NCHW input."""

    def __init__(self, repeats: int = 1, divisor: float = 255.0):
        self.repeats, self.divisor = int(repeats), float(divisor)

    def __call__(self, x_u8: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
        return u8_to_float(x_u8, self.repeats, self.divisor, out=out)


def _f32_scalar(v: float) -> float:
    return float(np.float32(v))


def color_jitter_apply(imgs: torch.Tensor, order: Sequence[int], brightness: Optional[float],
                       contrast: Optional[float], saturation: Optional[float], hue: Optional[float]) -> torch.Tensor:
    """The ColorJitter chain with GIVEN parameters (torchvision.transforms.ColorJitter.forward after get_params):
    `order` = permutation of (0 brightness, 1 contrast, 2 saturation, 3 hue); a None factor skips that adjustment."""
    if not imgs.is_cuda:
        raise RuntimeError("color_jitter: CUDA tensors only (the B200 path has no CPU fallback)")
    x = imgs.contiguous()
    if x.dtype != torch.float32:
        x = x.float()
    if x.dim() != 4 or x.shape[1] not in (1, 3):
        raise ValueError(f"color_jitter: expected (N, 1|3, H, W) images, got {tuple(x.shape)}")
    n, c, h, w = x.shape
    factors = {0: brightness, 1: contrast, 2: saturation, 3: hue}
    ops = [int(o) if factors[int(o)] is not None else -1 for o in order]
    order_arr = (C.c_int * 4)(*ops)
    b, ct, s = (1.0 if f is None else float(f) for f in (brightness, contrast, saturation))
    one_minus = (C.c_float * 3)(_f32_scalar(1.0 - b), _f32_scalar(1.0 - ct), _f32_scalar(1.0 - s))
    ws = torch.empty((n,), dtype=torch.float64, device=x.device) if 1 in ops else None
    y = torch.empty_like(x)
    _lib.call("msp_color_jitter", x.data_ptr(), n, c, h * w, order_arr, _f32_scalar(b), _f32_scalar(ct),
              _f32_scalar(s), float(0.0 if hue is None else hue), one_minus,
              None if ws is None else ws.data_ptr(), y.data_ptr(), _stream())
    return y


class ColorJitter:
    """torchvision.transforms.ColorJitter(brightness, contrast, saturation, hue) on the device.  The random draw is
    torchvision's own `ColorJitter.get_params` on the CPU generator (same order, same number of draws as the
    reference's call at robustness/eval.py:61-66), only the pixel arithmetic runs on the GPU."""

    def __init__(self, brightness=0, contrast=0, saturation=0, hue=0):
        import torchvision
        self._tv = torchvision.transforms.ColorJitter(brightness=brightness, contrast=contrast,
                                                      saturation=saturation, hue=hue)

    def get_params(self) -> Tuple:
        t = self._tv
        return t.get_params(t.brightness, t.contrast, t.saturation, t.hue)

    def __call__(self, imgs: torch.Tensor) -> torch.Tensor:
        fn_idx, b, c, s, h = self.get_params()
        return color_jitter_apply(imgs, [int(i) for i in fn_idx], b, c, s, h)
