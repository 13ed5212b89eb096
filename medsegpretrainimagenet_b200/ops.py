"""Thin, autograd-free Python wrappers over the C ABI (include/msp_b200.h).

Activations are bf16 tensors of logical shape (N, H, W, C) whose last dimension is contiguous and whose
pixel stride (`t.stride(2)`) may exceed C: a channel slice `buf[..., a:b]` of a wider buffer is a valid
operand, which is how torch.cat(dim=1) of the reference (blocks.py:628,635) becomes zero-copy.
PyTorch is used only for device memory and the current stream.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional, Tuple

import torch

from . import _lib
from ._lib import BnActDesc, ConvDesc, call

ACT_NONE, ACT_RELU, ACT_SIGMOID = 0, 1, 2
_BF16 = torch.bfloat16


def ceil8(c: int) -> int:
    return (c + 7) // 8 * 8


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _p(t: Optional[torch.Tensor]):
    return None if t is None else t.data_ptr()


def _chk_nhwc(t: torch.Tensor, name: str) -> Tuple[int, int, int, int, int]:
    if t.dtype != _BF16 or t.dim() != 4 or not t.is_cuda:
        raise ValueError(f"{name}: expected a CUDA bf16 (N,H,W,C) tensor, got {t.dtype} {tuple(t.shape)}")
    n, h, w, c = t.shape
    cs = t.stride(2)
    if t.stride(3) != 1 or t.stride(1) != w * cs or (n > 1 and t.stride(0) != h * w * cs):
        raise ValueError(f"{name}: not a pixel-major NHWC view (strides {t.stride()})")
    return n, h, w, c, cs


def deterministic() -> bool:
    """torch.use_deterministic_algorithms (run_experiment.py:65; every downstream YAML of the reference sets it): the
    floating-point reductions that normally finish with atomics (BatchNorm statistics in the conv epilogue, BatchNorm
    backward sums, bias gradients) then store per-CTA partial rows and add them in a fixed order."""
    return torch.are_deterministic_algorithms_enabled()


_SMS = {}


def sm_count(device) -> int:
    idx = torch.device(device).index
    idx = torch.cuda.current_device() if idx is None else idx
    if idx not in _SMS:
        _SMS[idx] = torch.cuda.get_device_properties(idx).multi_processor_count
    return _SMS[idx]


def new_stats(k: int, device) -> torch.Tensor:
    """Zeroed BatchNorm statistic accumulators of a conv epilogue: [2, K], or the per-CTA row workspace
    [SMs, 2, K] of the deterministic mode."""
    if deterministic():
        return torch.zeros((sm_count(device), 2, k), dtype=torch.float32, device=device)
    return torch.zeros((2, k), dtype=torch.float32, device=device)


def stats_total(stats: torch.Tensor) -> torch.Tensor:
    """[2, K] totals of either layout (tests / inspection; the step itself uses bn_finalize / reduce_rows)."""
    return stats if stats.dim() == 2 else stats.sum(0)


def _stat_args(stats):
    if stats is None:
        return None, None, 0
    if stats.dim() == 3:
        return stats[0, 0].data_ptr(), stats[0, 1].data_ptr(), stats.shape[0]
    return stats[0].data_ptr(), stats[1].data_ptr(), 0


def reduce_rows(ws: torch.Tensor, reset: bool = False) -> torch.Tensor:
    """[rows, *shape] fp32 -> [*shape]: the rows added in fixed order (second stage of the deterministic reductions)."""
    out = torch.empty(ws.shape[1:], dtype=torch.float32, device=ws.device)
    call("msp_reduce_rows", _p(ws), ws.shape[0], out.numel(), _p(out), int(reset), _stream())
    return out


def new_act(n: int, h: int, w: int, c: int, device, zero: bool = False) -> torch.Tensor:
    f = torch.zeros if zero else torch.empty
    return f((n, h, w, c), dtype=_BF16, device=device)


# ------------------------------------------------------------------------------------------------
# layout conversion
# ------------------------------------------------------------------------------------------------
def nchw_to_nhwc(x: torch.Tensor, cpad: Optional[int] = None) -> torch.Tensor:
    """fp32 NCHW -> bf16 NHWC with the channel dimension zero-padded to `cpad` (default ceil8(C))."""
    x = x.contiguous()
    if x.dtype != torch.float32:
        x = x.float()
    n, c, h, w = x.shape
    cpad = ceil8(c) if cpad is None else cpad
    y = new_act(n, h, w, cpad, x.device)
    call("msp_nchw_f32_to_nhwc_bf16", _p(x), n, c, h, w, cpad, _p(y), _stream())
    return y


def nhwc_to_nchw(x: torch.Tensor, c: Optional[int] = None) -> torch.Tensor:
    n, h, w, cc, cs = _chk_nhwc(x, "nhwc_to_nchw")
    c = cc if c is None else c
    y = torch.empty((n, c, h, w), dtype=torch.float32, device=x.device)
    call("msp_nhwc_bf16_to_nchw_f32", _p(x), n, c, h, w, cs, _p(y), _stream())
    return y


# ------------------------------------------------------------------------------------------------
# convolution
# ------------------------------------------------------------------------------------------------
def pack_weights(w: torch.Tensor, need_dgrad: bool = True):
    """OIHW fp32 -> (bf16 [K8][taps][C8] for fprop, bf16 [C8][taps][K8] for dgrad)."""
    w = w.detach()
    if w.dtype != torch.float32 or not w.is_contiguous():
        w = w.float().contiguous()
    k, c, kh, kw = w.shape
    c8, k8 = ceil8(c), ceil8(k)
    if k8 != k:
        raise ValueError("conv: output channels must be a multiple of 8")
    wf = torch.empty((k, kh * kw, c8), dtype=_BF16, device=w.device)
    wd = torch.empty((c8, kh * kw, k8), dtype=_BF16, device=w.device) if need_dgrad else None
    call("msp_pack_weights", _p(w), k, c, kh, kw, c8, k8, _p(wf), _p(wd), _stream())
    return wf, wd


class WeightPackCache:
    """bf16 operand copies of a model's convolution weights, repacked by ONE kernel per forward.

    The weights stay the reference's fp32 OIHW Parameters; the tensor-core kernels read bf16 [K][tap][C] / [C][tap][K]
    copies.  Packing them per layer inside every forward is ~50-80 launches of ~6.7 us for microseconds of work.  The
    cache keeps one persistent pair of buffers per weight (keyed by its storage address); `begin_step()` — the top of
    every forward of the converted model — repacks ALL of them with msp_pack_weights_batched, `lookup()` hands the
    buffers to the convolutions.  The repack is unconditional: tensor version counters are not a safe "unchanged"
    signal (fused optimizers and CUDA-graph replays update the parameters without moving them).  A weight seen for the
    first time (or while the table is being rebuilt) is packed on the spot, as before."""

    def __init__(self):
        self.entries = {}       # data_ptr -> [weight alias (OIHW fp32 source of the pack), wf, wd]
        self.folds = {}         # data_ptr of an up-conv weight -> (weight alias, folded fp32 [K, C, 1, 9] buffer)
        self.table = None
        self.total_blocks = 0
        self.dirty = False
        self.packed = False     # begin_step() of the running forward repacked every entry of the table
        self.in_table = set()

    def lookup(self, w: torch.Tensor, need_dgrad: bool):
        if w.dtype != torch.float32 or not w.is_contiguous():
            return pack_weights(w, need_dgrad)          # a temporary fp32 copy would be packed: not cacheable
        key = w.data_ptr()
        e = self.entries.get(key)
        if e is not None and e[0].shape == w.shape and (e[2] is not None or not need_dgrad):
            if not (self.packed and key in self.in_table):
                k, c, kh, kw = w.shape
                call("msp_pack_weights", _p(w.detach()), k, c, kh, kw, ceil8(c), ceil8(k), _p(e[1]), _p(e[2]), _stream())
            return e[1], (e[2] if need_dgrad else None)
        wf, wd = pack_weights(w, need_dgrad)
        self.entries[key] = [w.detach(), wf, wd]
        self.in_table.discard(key)                      # (a replaced entry's old buffers are what the table points to)
        self.dirty = True
        return wf, wd

    def build_table(self) -> None:
        """(Re)build the device table of the batched kernel at the end of a forward; never inside a graph capture."""
        self.packed = False
        if not self.dirty or not self.entries:
            return
        rows, first = [], 0
        for w, wf, wd in self.entries.values():
            k, c, kh, kw = w.shape
            if kh * kw <= 9:      # tiled transposing mode of the kernel: one block per 32 x 32 channel tile (all taps)
                nblk = ((ceil8(k) + 31) // 32) * ((ceil8(c) + 31) // 32)
            else:
                elems = k * kh * kw * ceil8(c) * (2 if wd is not None else 1)
                nblk = max(1, min(1024, (elems + 2047) // 2048))       # ~8 elements per thread
            rows.append([w.data_ptr(), wf.data_ptr(), wd.data_ptr() if wd is not None else 0, k, c, kh * kw,
                         ceil8(c), ceil8(k), first, nblk])
            first += nblk
        self.total_blocks = first
        dev = next(iter(self.entries.values()))[1].device
        self.table = torch.tensor(rows, dtype=torch.int64).to(dev)
        self.in_table = set(self.entries.keys())
        self.dirty = False

    def lookup_folded(self, w: torch.Tensor):
        """Operand copies of the FOLDED weights of an up-convolution (fold_upconv_weight): the fold kernel refreshes a
        persistent fp32 [K, C, 1, 9] buffer at the top of every forward, which then takes part in the batched pack."""
        if w.dtype != torch.float32 or not w.is_contiguous():
            return pack_weights(fold_upconv_weight(w), True)
        key = w.data_ptr()
        f = self.folds.get(key)
        if f is None or f[0].shape != w.shape:
            f = (w.detach(), torch.empty((w.shape[0], w.shape[1], 1, 9), dtype=torch.float32, device=w.device))
            self.folds[key] = f
            fold_upconv_weight(w, out=f[1])
        elif not self.packed:
            fold_upconv_weight(w, out=f[1])             # begin_step() has not refreshed it (table being rebuilt)
        return self.lookup(f[1], True)

    def begin_step(self) -> None:
        self.packed = False
        for w, buf in self.folds.values():              # up-conv weights: fold first, the batched pack reads the result
            fold_upconv_weight(w, out=buf)
        if self.table is None:
            return                                      # first forward: lookup() packs layer by layer
        call("msp_pack_weights_batched", self.table.data_ptr(), self.table.shape[0], self.total_blocks, _stream())
        self.packed = True

    def end_step(self) -> None:
        self.packed = False


_PACK: Optional[WeightPackCache] = None     # the cache of the model whose forward is running (converter sets it)
_PACK_ENABLED = os.environ.get("MSP_PACK_CACHE", "1") != "0"     # 0: pack layer by layer inside every forward


def set_active_pack_cache(cache: Optional[WeightPackCache]) -> None:
    global _PACK
    _PACK = cache


def fold_upconv_weight(w: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """(K, C, 2, 2) weights of nearest-x2 -> Conv2d(k=2, 'same') -> the (K, C, 1, 9) pre-summed taps of the four
    output-parity classes on the low-res input (msp_fold_upconv_weights; see include/msp_b200.h)."""
    w = w.detach()
    if w.dtype != torch.float32 or not w.is_contiguous():
        w = w.float().contiguous()
    k, c, kh, kw = w.shape
    assert (kh, kw) == (2, 2)
    if out is None:
        out = torch.empty((k, c, 1, 9), dtype=torch.float32, device=w.device)
    call("msp_fold_upconv_weights", _p(w), k, c, _p(out), _stream())
    return out


def packed_folded_weights(w: torch.Tensor):
    if _PACK is not None and _PACK_ENABLED:
        return _PACK.lookup_folded(w)
    return pack_weights(fold_upconv_weight(w), True)


def packed_weights(w: torch.Tensor, need_dgrad: bool = True):
    """Operand copies of `w` for conv_fprop / conv_dgrad: from the running model's cache, else packed now."""
    if _PACK is not None and _PACK_ENABLED:
        return _PACK.lookup(w, need_dgrad)
    return pack_weights(w, need_dgrad)


def set_conv_policy(pair: int = -1, halo: int = -1) -> None:
    """Kernel-variant policy of conv fprop / dgrad (see msp_conv_set_policy in include/msp_b200.h)."""
    call("msp_conv_set_policy", int(pair), int(halo))


def conv_out_size(h: int, w: int, kh: int, kw: int, stride: int, padding) -> Tuple[int, int, int, int]:
    """-> (Ho, Wo, pad_t, pad_l); padding is an int, (ph, pw) or 'same' (torch semantics: for even
    kernels the extra padding goes to the bottom/right, blocks.py:518 probe in SURVEY App. B)."""
    if padding == "same":
        if stride != 1:
            raise ValueError("padding='same' needs stride 1")
        return h, w, (kh - 1) // 2, (kw - 1) // 2
    if isinstance(padding, int):
        ph = pw = padding
    else:
        ph, pw = padding
    return (h + 2 * ph - kh) // stride + 1, (w + 2 * pw - kw) // stride + 1, ph, pw


# Optional per-call device timing of the convolution kernels (bench.py roofline): when a list is
# installed, every conv call appends (kind, algorithmic_flops, start_event, end_event, algorithmic_bytes,
# kernel variant name), the events being recorded on the launching stream.
_CONV_TIMELINE = None


def conv_timeline_active() -> bool:
    return _CONV_TIMELINE is not None


def set_conv_timeline(lst):
    global _CONV_TIMELINE
    _CONV_TIMELINE = lst


class _timed:
    def __init__(self, kind, flops, nbytes=0.0):
        self.kind, self.flops, self.nbytes = kind, flops, nbytes

    def __enter__(self):
        if _CONV_TIMELINE is not None:
            self.e0 = torch.cuda.Event(enable_timing=True)
            self.e1 = torch.cuda.Event(enable_timing=True)
            self.l0 = _lib.launch_count()
            self.e0.record()
        return self

    def __exit__(self, *exc):
        if _CONV_TIMELINE is not None:
            self.e1.record()
            name = (_lib.lib.msp_conv_last_kernel() or b"").decode()
            # last field: number of kernels this call launched (bench.py matches them, in stream order, with the
            # profiler's per-kernel device times: event pairs around a 10 us kernel of an eagerly enqueued step mostly
            # measure the host's launch gaps)
            _CONV_TIMELINE.append((self.kind, self.flops, self.e0, self.e1, self.nbytes, name,
                                   _lib.launch_count() - self.l0))
        return False


def _conv_desc(n, h, w, c, x_cs, ho, wo, k, y_cs, kh, kw, stride, pad_t, pad_l, relu=0, win_px=0,
               wp=0, stat_rows=0) -> ConvDesc:
    return ConvDesc(n, h, w, c, x_cs, ho, wo, k, y_cs, kh, kw, stride, pad_t, pad_l, relu, win_px, wp, stat_rows)


def conv_fprop(x, wf, bias, k, kh, kw, stride, pad_t, pad_l, ho, wo, relu=False, out=None, stats=None,
               c_true=None):
    """y = conv(x, w) [+ bias] [ReLU]; if `stats` (fp32 [2, K], zeroed by the caller) is given the
    per-channel sum / sum of squares of y are accumulated into it (BatchNorm batch statistics)."""
    n, h, w, c, x_cs = _chk_nhwc(x, "conv_fprop(x)")
    if out is None:
        out = new_act(n, ho, wo, k, x.device)
    _, _, _, ko, y_cs = _chk_nhwc(out, "conv_fprop(out)")
    assert ko == k and wf.shape[2] == c, (ko, k, tuple(wf.shape), c)
    s1, s2, rows = _stat_args(stats)
    d = _conv_desc(n, h, w, c, x_cs, ho, wo, k, y_cs, kh, kw, stride, pad_t, pad_l, int(relu), stat_rows=rows)
    with _timed("fprop", 2.0 * n * ho * wo * k * (c_true or c) * kh * kw,
                2.0 * (n * h * w * c + n * ho * wo * k + k * c * kh * kw)):
        call("msp_conv_fprop", C.byref(d), _p(x), _p(wf), _p(bias), _p(out), s1, s2, _stream())
    return out


def conv_dgrad(dy, wd, x_shape, kh, kw, stride, pad_t, pad_l, out=None, accumulate=False, c_true=None):
    """dx = conv_transpose(dy, w); x_shape = (N, H, W, C8)."""
    n, ho, wo, k, y_cs = _chk_nhwc(dy, "conv_dgrad(dy)")
    _, h, w, c = x_shape
    if out is None:
        out = new_act(n, h, w, c, dy.device)
        accumulate = False
    _, _, _, _, x_cs = _chk_nhwc(out, "conv_dgrad(out)")
    d = _conv_desc(n, h, w, c, x_cs, ho, wo, k, y_cs, kh, kw, stride, pad_t, pad_l)
    with _timed("dgrad", 2.0 * n * ho * wo * k * (c_true or c) * kh * kw,
                2.0 * (n * h * w * c + n * ho * wo * k + k * c * kh * kw)):
        call("msp_conv_dgrad", C.byref(d), _p(dy), _p(wd), _p(out), int(accumulate), _stream())
    return out


def upconv2x_fprop(x, wf9, bias, k, relu=True, out=None, c_true=None):
    """nearest x2 -> conv 2x2 'same' (+bias, ReLU) on the LOW-RES x (N, H, W, C8) -> (N, 2H, 2W, K)."""
    n, h, w, c, x_cs = _chk_nhwc(x, "upconv2x(x)")
    if out is None:
        out = new_act(n, 2 * h, 2 * w, k, x.device)
    _, _, _, ko, y_cs = _chk_nhwc(out, "upconv2x(out)")
    assert ko == k and wf9.shape[2] == c and wf9.shape[1] == 9
    d = _conv_desc(n, h, w, c, x_cs, 2 * h, 2 * w, k, y_cs, 2, 2, 1, 0, 0, int(relu))
    # 9 folded taps per 2x2 output block = 9/4 per output pixel: the FLOPs actually executed
    with _timed("fprop", 2.0 * n * h * w * 9 * k * (c_true or c), 2.0 * (n * h * w * c + 4 * n * h * w * k + 9 * k * c)):
        call("msp_upconv2x_fprop", C.byref(d), _p(x), _p(wf9), _p(bias), _p(out), _stream())
    return out


def upconv2x_dgrad(dy, wd9, x_shape, c_true=None):
    n, h2, w2, k, y_cs = _chk_nhwc(dy, "upconv2x_dgrad(dy)")
    _, h, w, c = x_shape
    dx = new_act(n, h, w, c, dy.device)
    d = _conv_desc(n, h, w, c, c, h2, w2, k, y_cs, 2, 2, 1, 0, 0)
    with _timed("dgrad", 2.0 * n * h * w * 9 * k * (c_true or c), 2.0 * (n * h * w * c + 4 * n * h * w * k + 9 * k * c)):
        call("msp_upconv2x_dgrad", C.byref(d), _p(dy), _p(wd9), _p(dx), 0, _stream())
    return dx


def upconv2x_wgrad(x, dy, weight: torch.Tensor):
    """Weight gradient of the folded up-conv in the layout of the ORIGINAL (K, C, 2, 2) filter: one wgrad launch per
    output-parity class on the low-res x, the class gradients unpacked ([K, C, 1+a, 1+b]) and mapped back through the
    (linear) fold.  With a leaf Parameter the result goes straight into `weight.grad` at the end of the backward pass
    (ops._WgradQueue); otherwise it is returned."""
    n, h, w, c, x_cs = _chk_nhwc(x, "upconv2x_wgrad(x)")
    _, h2, w2, k, y_cs = _chk_nhwc(dy, "upconv2x_wgrad(dy)")
    c_true = weight.shape[1]
    sink = _WGRAD_SINK and weight.is_leaf and weight.dtype == torch.float32 and weight.is_contiguous() and \
        (weight.grad is None or (weight.grad.dtype == torch.float32 and weight.grad.is_contiguous()))
    side = _WGRAD_STREAM if (_CONV_TIMELINE is None and sink) else None
    classes = []
    if side is not None:
        side.wait_stream(torch.cuda.current_stream())
    for a in (0, 1):
        for b in (0, 1):
            d = _conv_desc(n, h, w, c, x_cs, h, w, k, y_cs, 1 + a, 1 + b, 1, 0, 0)
            splits = _lib.lib.msp_upconv2x_wgrad_splits(C.byref(d))
            _lib.check(0 if splits >= 1 else splits, "msp_upconv2x_wgrad_splits")
            taps = (1 + a) * (1 + b)
            flops, nbytes = 2.0 * n * h * w * taps * k * c_true, 2.0 * (n * h * w * c + n * h * w * k) + 4.0 * k * c_true * taps
            if side is not None:
                with torch.cuda.stream(side):
                    part = torch.empty((splits, k, taps, c), dtype=torch.float32, device=x.device)
                    call("msp_upconv2x_wgrad_class", C.byref(d), _p(x), _p(dy), a, b, _p(part), side.cuda_stream)
            else:
                part = torch.empty((splits, k, taps, c), dtype=torch.float32, device=x.device)
                with _timed("wgrad", flops, nbytes):
                    call("msp_upconv2x_wgrad_class", C.byref(d), _p(x), _p(dy), a, b, _p(part), _stream())
            g = torch.empty((k, c_true, 1 + a, 1 + b), dtype=torch.float32, device=x.device)
            classes.append((d, part, g))
    if side is not None:
        x.record_stream(side)
        dy.record_stream(side)
    if sink:
        acc = weight.grad is not None
        if not acc:
            weight.grad = torch.empty_like(weight, memory_format=torch.contiguous_format)
        dst = weight.grad
        for d, part, g in classes:
            _WGRAD_QUEUE.push(d, part, g, c_true, False)
        gs = [g for _, _, g in classes]
        _WGRAD_QUEUE.after_unpack(lambda: call("msp_unfold_upconv_wgrad", _p(gs[0]), _p(gs[1]), _p(gs[2]), _p(gs[3]), k,
                                               c_true, _p(dst), int(acc), _stream()), keep=(gs, dst))
        return None
    for d, part, g in classes:
        it = _WGRAD_QUEUE.make_item(d, part, g, c_true, False)
        call("msp_unpack_wgrad_batched", 1, (_lib.UnpackItem * 1)(it), _stream())
    dw = torch.empty((k, c_true, 2, 2), dtype=torch.float32, device=x.device)
    call("msp_unfold_upconv_wgrad", _p(classes[0][2]), _p(classes[1][2]), _p(classes[2][2]), _p(classes[3][2]), k, c_true,
         _p(dw), 0, _stream())
    return dw


def conv_transpose_fprop(x, wd, bias, c_out, kh, kw, stride, pad, relu=False, out=None, k_true=None):
    """nn.ConvTranspose2d forward = data gradient of the conv it transposes (msp_conv_transpose_fprop).
    x (N, Hi, Wi, K8) -> (N, Ho, Wo, C8) with Ho = (Hi - 1) * stride - 2 * pad + kh."""
    n, hi, wi, k, x_cs = _chk_nhwc(x, "conv_transpose(x)")
    ho, wo = (hi - 1) * stride - 2 * pad + kh, (wi - 1) * stride - 2 * pad + kw
    if out is None:
        out = new_act(n, ho, wo, c_out, x.device)
    _, _, _, _, y_cs = _chk_nhwc(out, "conv_transpose(out)")
    # descriptor of the transposed convolution's "parent": input (N, ho, wo, c_out) -> output (N, hi, wi, k)
    d = _conv_desc(n, ho, wo, c_out, y_cs, hi, wi, k, x_cs, kh, kw, stride, pad, pad)
    with _timed("dgrad", 2.0 * n * hi * wi * (k_true or k) * c_out * kh * kw,
                2.0 * (n * hi * wi * k + n * ho * wo * c_out + k * c_out * kh * kw)):
        call("msp_conv_transpose_fprop", C.byref(d), _p(x), _p(wd), _p(bias), int(relu), _p(out), _stream())
    return out


def _wgrad(d: ConvDesc, x, dy, c_true, flops, nbytes=0.0) -> torch.Tensor:
    splits = _lib.lib.msp_conv_wgrad_splits(C.byref(d))
    _lib.check(0 if splits >= 1 else splits, "msp_conv_wgrad_splits")
    taps = d.KH if d.win_px else d.KH * d.KW
    cw = 64 if d.win_px else d.C
    part = torch.empty((splits, d.K, taps, cw), dtype=torch.float32, device=x.device)
    dw = torch.empty((d.K, c_true, d.KH, d.KW), dtype=torch.float32, device=x.device)
    with _timed("wgrad", flops, nbytes):
        call("msp_conv_wgrad", C.byref(d), _p(x), _p(dy), _p(part), _stream())
        call("msp_unpack_wgrad", C.byref(d), _p(part), c_true, _p(dw), _stream())
    return dw


# ------------------------------------------------------------------------------------------------
# weight gradients straight into `param.grad`, unpacked by ONE kernel per backward pass
# ------------------------------------------------------------------------------------------------
class _WgradQueue:
    """Split-K partials of the wgrad launches of the running backward pass, waiting for msp_unpack_wgrad_batched.

    Per-layer unpack launches cost ~8 us each for microseconds of work (5 % of the R50 U-Net step).  A convolution whose
    weight is a leaf Parameter therefore writes its gradient into `weight.grad` itself (allocated here when absent, ADDED
    to when present: gradient accumulation, all-reduce bucket views) instead of returning it to autograd, and only
    queues the fixed-order split sum; `flush()` — queued as an end-of-backward callback of the autograd engine, and called
    by parallel.GradReducer before it reduces a bucket — issues one kernel for everything queued.  Nobody can read
    `weight.grad` in between: autograd never sees these gradients."""

    def __init__(self):
        self.items, self.keep, self.callback_queued = [], [], False
        self.post = []          # calls issued after the unpack launches of a flush (the up-conv's unfold)

    def after_unpack(self, fn, keep=None):
        self.post.append(fn)
        self.keep.append(keep)

    def push(self, d: ConvDesc, part, dst, c_true, accumulate):
        self.items.append(self.make_item(d, part, dst, c_true, accumulate))
        self.keep.append((part, dst))
        if not self.callback_queued:
            try:
                torch.autograd.Variable._execution_engine.queue_callback(self.flush)
                self.callback_queued = True
            except RuntimeError:
                self.flush()                    # not inside a backward pass: nothing to defer to
                return
        if _WGRAD_STREAM is not None and _CONV_TIMELINE is None and len(self.items) >= _WGRAD_CHUNK:
            # side-stream mode: unpack in chunks DURING the backward pass, behind the wgrad kernels on their own stream —
            # one unpack of everything at the end sits exposed between the backward pass and the optimizer (0.5 ms of the
            # 10 ms R50 U-Net step); the end-of-backward callback takes the rest and joins the stream
            self.flush(on_side=True)

    @staticmethod
    def make_item(d: ConvDesc, part, dst, c_true, accumulate):
        it = _lib.UnpackItem()
        it.partials, it.dst = part.data_ptr(), dst.data_ptr()
        it.splits, it.K, it.C_true = part.shape[0], d.K, c_true
        it.split_stride = part.stride(0)
        if d.win_px:
            it.rowwin_KH, it.rowwin_KW, it.rowwin_cpp, it.taps, it.Cpad = d.KH, d.KW, d.C, d.KH, 64
        else:
            it.rowwin_KH, it.taps, it.Cpad = 0, d.KH * d.KW, d.C
        it.accumulate = int(accumulate)
        return it

    def flush(self, on_side: bool = False):
        """`on_side`: issue the unpack ON the wgrad side stream (parallel.GradReducer: a bucket's unpack and all-reduce
        then follow the wgrad kernels they depend on without stalling the dgrad / BatchNorm-backward chain of the main
        stream; the caller has made the side stream wait for the main stream).  Otherwise the current stream joins the
        side stream first."""
        self.callback_queued = False
        items, self.items = self.items, []
        keep, self.keep = self.keep, []
        post, self.post = self.post, []
        side = _WGRAD_STREAM if (on_side and _WGRAD_STREAM is not None) else None
        if side is None and _WGRAD_STREAM is not None:
            cur = torch.cuda.current_stream()
            cur.wait_stream(_WGRAD_STREAM)          # every queued wgrad kernel / side-stream unpack has finished
            for kp in keep:
                if kp is not None and isinstance(kp[0], torch.Tensor):
                    kp[0].record_stream(cur)

        def issue():
            for lo in range(0, len(items), 96):
                chunk = items[lo:lo + 96]
                arr = (_lib.UnpackItem * len(chunk))(*chunk)
                call("msp_unpack_wgrad_batched", len(chunk), arr, _stream())
            for fn in post:
                fn()

        if side is not None:
            with torch.cuda.stream(side):
                issue()
        else:
            issue()
        del keep


# queued weight gradients per side-stream unpack DURING the backward pass; off by default: measured on one box,
# 20 per chunk was 2.8 % slower on the R50 U-Net step than one unpack at the end (10.14 vs 9.86 ms; the chunks' blocks
# compete with the dgrad / BatchNorm-backward chain for the SMs), 0.6 % slower on ResNet-50
_WGRAD_CHUNK = int(os.environ.get("MSP_WGRAD_CHUNK", "1000000"))
_WGRAD_QUEUE = _WgradQueue()
_WGRAD_SINK = os.environ.get("MSP_WGRAD_SINK", "1") != "0"    # 0: return dW to autograd, one unpack launch per layer
_WGRAD_STREAM: Optional[torch.cuda.Stream] = None


def set_wgrad_stream(stream: Optional[torch.cuda.Stream]) -> None:
    """Run the wgrad kernels of the sink path on `stream` (None: the current stream).  The weight gradients are leaves
    of the backward pass: nothing downstream waits for them, while the dgrad / BatchNorm-backward chain is serial.  On
    a side stream they fill the SMs the chain's small kernels leave idle (small batches: most launches of the R50
    U-Net at batch 24 use a fraction of the 148 SMs).  The stream forks from the current stream before every launch
    and is joined by `flush_wgrad()`; CUDA-graph capture records the fork / join as graph branches."""
    global _WGRAD_STREAM
    flush_wgrad()
    _WGRAD_STREAM = stream


def flush_wgrad(on_side: bool = False) -> None:
    _WGRAD_QUEUE.flush(on_side)


def wgrad_stream() -> Optional[torch.cuda.Stream]:
    return _WGRAD_STREAM if _CONV_TIMELINE is None else None


def join_wgrad_stream() -> None:
    """The current stream waits for everything issued on the wgrad side stream (end of a step / graph capture)."""
    if _WGRAD_STREAM is not None:
        torch.cuda.current_stream().wait_stream(_WGRAD_STREAM)


def wgrad_into_param(d: ConvDesc, x, dy, weight: torch.Tensor, c_true, flops, nbytes=0.0) -> bool:
    """Launch the wgrad kernel of one convolution and queue the unpack into `weight.grad`.  Returns False when the
    parameter does not qualify (the caller then takes the autograd route)."""
    if not _WGRAD_SINK or not weight.is_leaf or weight.dtype != torch.float32 or not weight.is_contiguous():
        return False
    g = weight.grad
    if g is not None and (g.dtype != torch.float32 or not g.is_contiguous() or g.shape != weight.shape
                          or g.device != weight.device):
        return False
    splits = _lib.lib.msp_conv_wgrad_splits(C.byref(d))
    _lib.check(0 if splits >= 1 else splits, "msp_conv_wgrad_splits")
    taps = d.KH if d.win_px else d.KH * d.KW
    cw = 64 if d.win_px else d.C
    side = _WGRAD_STREAM if _CONV_TIMELINE is None else None     # instrumented runs time every kernel in stream order
    if side is not None:
        side.wait_stream(torch.cuda.current_stream())             # x and dy are complete
        with torch.cuda.stream(side):
            part = torch.empty((splits, d.K, taps, cw), dtype=torch.float32, device=x.device)
            call("msp_conv_wgrad", C.byref(d), _p(x), _p(dy), _p(part), side.cuda_stream)
        x.record_stream(side)
        dy.record_stream(side)
    else:
        part = torch.empty((splits, d.K, taps, cw), dtype=torch.float32, device=x.device)
        with _timed("wgrad", flops, nbytes):
            call("msp_conv_wgrad", C.byref(d), _p(x), _p(dy), _p(part), _stream())
    accumulate = g is not None
    if g is None:
        g = torch.empty_like(weight, memory_format=torch.contiguous_format)
        weight.grad = g
    _WGRAD_QUEUE.push(d, part, g, c_true, accumulate)
    return True


def conv_wgrad_param(x, dy, weight, kh, kw, stride, pad_t, pad_l) -> bool:
    n, h, w, c, x_cs = _chk_nhwc(x, "conv_wgrad(x)")
    _, ho, wo, k, y_cs = _chk_nhwc(dy, "conv_wgrad(dy)")
    c_true = weight.shape[1]
    d = _conv_desc(n, h, w, c, x_cs, ho, wo, k, y_cs, kh, kw, stride, pad_t, pad_l)
    return wgrad_into_param(d, x, dy, weight, c_true, 2.0 * n * ho * wo * k * c_true * kh * kw,
                            2.0 * (n * h * w * c + n * ho * wo * k) + 4.0 * k * c_true * kh * kw)


def conv_wgrad_rowwin_param(xw, w_img, dy, weight, kh, kw, stride, pad_t, pad_l, win_px) -> bool:
    n, ho, wo, k, y_cs = _chk_nhwc(dy, "conv_wgrad_rowwin(dy)")
    d = _rowwin_desc(xw, w_img, k, kh, kw, stride, pad_t, pad_l, ho, wo, y_cs, win_px)
    c_true = weight.shape[1]
    return wgrad_into_param(d, xw, dy, weight, c_true, 2.0 * n * ho * wo * k * c_true * kh * kw,
                            2.0 * (xw.numel() + n * ho * wo * k) + 4.0 * k * c_true * kh * kw)


def conv_wgrad(x, dy, c_true, kh, kw, stride, pad_t, pad_l) -> torch.Tensor:
    """-> dW in the OIHW fp32 layout of nn.Conv2d.weight.grad (deterministic split-K: per-split partials
    summed in a fixed order by msp_unpack_wgrad)."""
    n, h, w, c, x_cs = _chk_nhwc(x, "conv_wgrad(x)")
    _, ho, wo, k, y_cs = _chk_nhwc(dy, "conv_wgrad(dy)")
    d = _conv_desc(n, h, w, c, x_cs, ho, wo, k, y_cs, kh, kw, stride, pad_t, pad_l)
    return _wgrad(d, x, dy, c_true, 2.0 * n * ho * wo * k * c_true * kh * kw,
                  2.0 * (n * h * w * c + n * ho * wo * k) + 4.0 * k * c_true * kh * kw)


# ------------------------------------------------------------------------------------------------
# row-window convolution: the tiny-channel first layer (7x7/2 stem, 3x3 first block)
# ------------------------------------------------------------------------------------------------
def rowwin_geometry(c_true: int, kw: int, stride: int):
    """-> (win_px, cpp) or None when the layer does not qualify (see msp_conv.cu)."""
    if stride not in (1, 2):
        return None
    if c_true <= 8 and kw <= 8:
        return 8, 8          # 8 pixels x 8 channels per window
    if c_true <= 16 and kw <= 4:
        return 4, 16         # 4 pixels x 16 channels per window
    return None


def nchw_to_rowwin(x: torch.Tensor, cpp: int, pad_l: int, wp: int) -> torch.Tensor:
    """fp32 (or bf16) NCHW -> bf16 [N, H, Wp, cpp], image column w at w + pad_l, zero elsewhere."""
    x = x.contiguous()
    if x.dtype not in (torch.float32, _BF16):
        x = x.float()
    n, c, h, w = x.shape
    y = torch.empty((n, h, wp, cpp), dtype=_BF16, device=x.device)
    fn = "msp_nchw_bf16_to_rowwin_bf16" if x.dtype == _BF16 else "msp_nchw_f32_to_rowwin_bf16"
    call(fn, _p(x), n, c, h, w, cpp, pad_l, wp, _p(y), _stream())
    return y


def pack_weights_rowwin(w: torch.Tensor, win_px: int) -> torch.Tensor:
    w = w.detach()
    if w.dtype != torch.float32 or not w.is_contiguous():
        w = w.float().contiguous()
    k, c, kh, kw = w.shape
    out = torch.empty((k, kh, 64), dtype=_BF16, device=w.device)
    call("msp_pack_weights_rowwin", _p(w), k, c, kh, kw, win_px, _p(out), _stream())
    return out


def _rowwin_desc(xw, w_img, k, kh, kw, stride, pad_t, pad_l, ho, wo, y_cs, win_px, relu=0, stat_rows=0):
    n, h, wp, cpp = xw.shape
    return _conv_desc(n, h, w_img, cpp, cpp, ho, wo, k, y_cs, kh, kw, stride, pad_t, pad_l, relu, win_px, wp,
                      stat_rows)


def conv_fprop_rowwin(xw, w_img, wr, bias, k, kh, kw, stride, pad_t, pad_l, ho, wo, win_px, relu=False,
                      out=None, stats=None, c_true=None):
    n = xw.shape[0]
    if out is None:
        out = new_act(n, ho, wo, k, xw.device)
    _, _, _, _, y_cs = _chk_nhwc(out, "conv_fprop_rowwin(out)")
    s1, s2, rows = _stat_args(stats)
    d = _rowwin_desc(xw, w_img, k, kh, kw, stride, pad_t, pad_l, ho, wo, y_cs, win_px, int(relu), stat_rows=rows)
    with _timed("fprop", 2.0 * n * ho * wo * k * (c_true or xw.shape[3]) * kh * kw,
                2.0 * (xw.numel() + n * ho * wo * k)):
        call("msp_conv_fprop", C.byref(d), _p(xw), _p(wr), _p(bias), _p(out), s1, s2, _stream())
    return out


def conv_wgrad_rowwin(xw, w_img, dy, c_true, kh, kw, stride, pad_t, pad_l, win_px) -> torch.Tensor:
    n, ho, wo, k, y_cs = _chk_nhwc(dy, "conv_wgrad_rowwin(dy)")
    d = _rowwin_desc(xw, w_img, k, kh, kw, stride, pad_t, pad_l, ho, wo, y_cs, win_px)
    return _wgrad(d, xw, dy, c_true, 2.0 * n * ho * wo * k * c_true * kh * kw)


def channel_sum(x: torch.Tensor) -> torch.Tensor:
    n, h, w, c, cs = _chk_nhwc(x, "channel_sum")
    out = torch.empty((c,), dtype=torch.float32, device=x.device)
    ws, rows = None, 0
    if deterministic():
        rows = 4 * sm_count(x.device)
        ws = torch.empty((rows, c), dtype=torch.float32, device=x.device)
    call("msp_channel_sum", _p(x), n * h * w, c, cs, _p(out), _p(ws), rows, _stream())
    return out


# ------------------------------------------------------------------------------------------------
# BatchNorm (+ activation + residual)
# ------------------------------------------------------------------------------------------------
def bn_finalize(stats, count, eps, momentum, running_mean=None, running_var=None, reset=False):
    """`stats`: [2, C] sums, or the [rows, 2, C] per-CTA workspace of the deterministic mode (rows added in fixed order)."""
    c = stats.shape[-1]
    s1, s2, rows = _stat_args(stats)
    mi = torch.empty((2, c), dtype=torch.float32, device=stats.device)
    call("msp_bn_finalize", s1, s2, c, float(count), float(eps),
         float(momentum), mi[0].data_ptr(), mi[1].data_ptr(), _p(running_mean), _p(running_var), int(reset),
         max(rows, 1), _stream())
    return mi


def bn_eval_stats(running_mean, running_var, eps):
    c = running_mean.shape[0]
    mi = torch.empty((2, c), dtype=torch.float32, device=running_mean.device)
    mi[0].copy_(running_mean)
    call("msp_bn_eval_prepare", _p(running_var), c, float(eps), mi[1].data_ptr(), _stream())
    return mi


def _bn_desc(x, y, act, res, r_stride):
    n, h, w, c, x_cs = _chk_nhwc(x, "bn_act(x)")
    _, _, _, _, y_cs = _chk_nhwc(y, "bn_act(y)")
    r_c = r_cs = 0
    if res is not None:
        _, _, _, r_c, r_cs = _chk_nhwc(res, "bn_act(residual)")
        r_c = min(r_c, c)
    return BnActDesc(n, h, w, c, x_cs, y_cs, act, r_c, r_cs, r_stride)


def bn_act_fwd(x, mi, gamma, beta, act, residual=None, r_stride=1, sample_scale=None, out=None):
    if out is None:
        out = new_act(x.shape[0], x.shape[1], x.shape[2], x.shape[3], x.device)
    d = _bn_desc(x, out, act, residual, r_stride)
    call("msp_bn_act_fwd", C.byref(d), _p(x), mi[0].data_ptr(), mi[1].data_ptr(), _p(gamma), _p(beta),
         _p(sample_scale), _p(residual), _p(out), _stream())
    return out


def bn_act_bwd_reduce(x, y, dy, mi, act, sample_scale=None, gamma=None, beta=None):
    """`y` may be None for BatchNorm -> ReLU without shortcut / sample scale: the mask is recomputed from x."""
    d = _bn_desc(x, dy if y is None else y, act, None, 1)
    _chk_nhwc(dy, "bn_act_bwd(dy)")
    if y is not None:
        assert dy.stride(2) == y.stride(2), "dy must share the pixel stride of y"
    sums = torch.empty((2, x.shape[3]), dtype=torch.float32, device=x.device)
    ws, rows = None, 0
    if deterministic():
        rows = 2 * sm_count(x.device)
        ws = torch.empty((rows, 2, x.shape[3]), dtype=torch.float32, device=x.device)
    call("msp_bn_act_bwd_reduce", C.byref(d), _p(x), _p(y), _p(dy), mi[0].data_ptr(), mi[1].data_ptr(),
         _p(gamma), _p(beta), _p(sample_scale), sums[0].data_ptr(), sums[1].data_ptr(), _p(ws), rows, _stream())
    return sums


def bn_act_bwd_reduce_rows(x, y, dy, mi, act, sample_scale=None, gamma=None, beta=None):
    """First stage of the BatchNorm-backward sums only: ([rows, 2, C] workspace, rows written).  For the SyncBN path, whose
    exchange kernel adds the rows itself (parallel.PeerAllReduce.stats_exchange) — in either reduction mode: the
    per-block rows cost nothing extra and spare the atomics' memset."""
    d = _bn_desc(x, dy if y is None else y, act, None, 1)
    _chk_nhwc(dy, "bn_act_bwd(dy)")
    if y is not None:
        assert dy.stride(2) == y.stride(2), "dy must share the pixel stride of y"
    rows = 2 * sm_count(x.device)
    ws = torch.empty((rows, 2, x.shape[3]), dtype=torch.float32, device=x.device)
    used = C.c_int(0)
    call("msp_bn_act_bwd_reduce_rows", C.byref(d), _p(x), _p(y), _p(dy), mi[0].data_ptr(), mi[1].data_ptr(),
         _p(gamma), _p(beta), _p(sample_scale), _p(ws), rows, C.byref(used), _stream())
    return ws, used.value


def bn_act_bwd_apply(x, y, dy, mi, gamma, act, sums, count, residual_like=None, r_stride=1,
                     sample_scale=None, dres=None, dres_accumulate=False, beta=None):
    n, h, w, c = x.shape
    if x.stride(2) == c:
        dx = new_act(n, h, w, c, x.device)
    else:  # x is a channel slice of a wider buffer: dx must share its pixel stride
        dx = torch.empty((n, h, w, x.stride(2)), dtype=_BF16, device=x.device)[..., :c]
    d = _bn_desc(x, dy if y is None else y, act, residual_like if dres is None else dres, r_stride)
    call("msp_bn_act_bwd_apply", C.byref(d), _p(x), _p(y), _p(dy), mi[0].data_ptr(), mi[1].data_ptr(),
         _p(gamma), _p(beta), _p(sample_scale), sums[0].data_ptr(), sums[1].data_ptr(), float(count), _p(dx),
         _p(dres), int(dres_accumulate), _stream())
    return dx


# ------------------------------------------------------------------------------------------------
# pooling / resampling / elementwise
# ------------------------------------------------------------------------------------------------
def maxpool_fwd(x, k, stride, pad, want_idx=True):
    n, h, w, c, cs = _chk_nhwc(x, "maxpool(x)")
    ho, wo = (h + 2 * pad - k) // stride + 1, (w + 2 * pad - k) // stride + 1
    y = new_act(n, ho, wo, c, x.device)
    idx = torch.empty((n, ho, wo, c), dtype=torch.uint8, device=x.device) if want_idx else None
    call("msp_maxpool_fwd", _p(x), n, h, w, c, cs, k, stride, pad, _p(y), _p(idx), ho, wo, c, _stream())
    return y, idx


def maxpool_bwd(idx, dy, x_shape, k, stride, pad):
    n, h, w, c = x_shape
    _, ho, wo, _, dy_cs = _chk_nhwc(dy, "maxpool_bwd(dy)")
    dx = new_act(n, h, w, c, dy.device)
    call("msp_maxpool_bwd", _p(idx), _p(dy), n, h, w, c, k, stride, pad, ho, wo, dy_cs, _p(dx), c, 0,
         _stream())
    return dx


def upsample2x_fwd(x, out=None):
    n, h, w, c, cs = _chk_nhwc(x, "upsample2x(x)")
    if out is None:
        out = new_act(n, 2 * h, 2 * w, c, x.device)
    call("msp_upsample2x_fwd", _p(x), n, h, w, c, cs, _p(out), out.stride(2), _stream())
    return out


def upsample2x_bwd(dy):
    n, h2, w2, c, cs = _chk_nhwc(dy, "upsample2x_bwd(dy)")
    dx = new_act(n, h2 // 2, w2 // 2, c, dy.device)
    call("msp_upsample2x_bwd", _p(dy), n, h2 // 2, w2 // 2, c, cs, _p(dx), c, _stream())
    return dx


def upsample_bilinear2x_fwd(x, out=None):
    n, h, w, c, cs = _chk_nhwc(x, "upsample_bilinear2x(x)")
    if out is None:
        out = new_act(n, 2 * h, 2 * w, c, x.device)
    call("msp_upsample_bilinear2x_fwd", _p(x), n, h, w, c, cs, _p(out), out.stride(2), _stream())
    return out


def upsample_bilinear2x_bwd(dy):
    n, h2, w2, c, cs = _chk_nhwc(dy, "upsample_bilinear2x_bwd(dy)")
    dx = new_act(n, h2 // 2, w2 // 2, c, dy.device)
    call("msp_upsample_bilinear2x_bwd", _p(dy), n, h2 // 2, w2 // 2, c, cs, _p(dx), c, _stream())
    return dx


def avgpool_fwd(x):
    n, h, w, c, cs = _chk_nhwc(x, "avgpool(x)")
    y = new_act(n, 1, 1, c, x.device)
    call("msp_avgpool_fwd", _p(x), n, h * w, c, cs, _p(y), _stream())
    return y


def avgpool_bwd(dy, h, w):
    n, _, _, c, _ = _chk_nhwc(dy, "avgpool_bwd(dy)")
    dyc = dy if dy.stride(2) == c else dy.contiguous()
    dx = new_act(n, h, w, c, dy.device)
    call("msp_avgpool_bwd", _p(dyc), n, h * w, c, _p(dx), c, _stream())
    return dx


def copy_channels(x, out):
    n, h, w, c, cs = _chk_nhwc(x, "copy_channels(x)")
    _, _, _, co, ocs = _chk_nhwc(out, "copy_channels(out)")
    assert co == c
    call("msp_copy_channels", _p(x), n * h * w, c, cs, _p(out), ocs, _stream())
    return out


def _ew2(name, a, b, out=None):
    n, h, w, c, a_cs = _chk_nhwc(a, name + "(a)")
    _, _, _, _, b_cs = _chk_nhwc(b, name + "(b)")
    if out is None:
        out = new_act(n, h, w, c, a.device)
    call(name, _p(a), _p(b), n * h * w, c, a_cs, b_cs, _p(out), out.stride(2), _stream())
    return out


def add_relu(a, b, out=None):
    return _ew2("msp_add_relu_fwd", a, b, out)


def add(a, b, out=None):
    return _ew2("msp_add", a, b, out)


def relu_bwd(y, dy, out=None):
    return _ew2("msp_relu_bwd", y, dy, out)


def gate_mul_fwd(skip, p, out=None):
    n, h, w, c, s_cs = _chk_nhwc(skip, "gate_mul(skip)")
    _, _, _, _, p_cs = _chk_nhwc(p, "gate_mul(p)")
    if out is None:
        out = new_act(n, h, w, c, skip.device)
    call("msp_gate_mul_fwd", _p(skip), _p(p), n, h, w, c, s_cs, p_cs, _p(out), out.stride(2), _stream())
    return out


def gate_mul_bwd(skip, p, dy):
    n, h, w, c, s_cs = _chk_nhwc(skip, "gate_mul_bwd(skip)")
    _, _, _, _, p_cs = _chk_nhwc(p, "gate_mul_bwd(p)")
    _, _, _, _, dy_cs = _chk_nhwc(dy, "gate_mul_bwd(dy)")
    dskip = new_act(n, h, w, c, skip.device)
    dp = new_act(n, h // 2, w // 2, c, skip.device)
    call("msp_gate_mul_bwd", _p(skip), _p(p), _p(dy), n, h, w, c, s_cs, p_cs, dy_cs, _p(dskip), c, 0,
         _p(dp), c, _stream())
    return dskip, dp


# ------------------------------------------------------------------------------------------------
# heads / losses
# ------------------------------------------------------------------------------------------------
HEAD_ACT = {None: 0, "none": 0, "sigmoid": 1, "softmax": 2}


def final_conv_act_fwd(x, w2d, bias, act: int, want_logits=False):
    n, h, w, c, cs = _chk_nhwc(x, "final_conv(x)")
    k = w2d.shape[0]
    prob = torch.empty((n, k, h, w), dtype=torch.float32, device=x.device)
    logits = torch.empty_like(prob) if want_logits else None
    call("msp_final_conv_act_fwd", _p(x), n, h, w, c, cs, _p(w2d), _p(bias), k, act, _p(logits), _p(prob),
         _stream())
    return prob, logits


def final_conv_act_bwd(x, w2d, act: int, prob, dprob, need_dx=True, has_bias=True):
    n, h, w, c, cs = _chk_nhwc(x, "final_conv_bwd(x)")
    k = w2d.shape[0]
    dx = new_act(n, h, w, c, x.device) if need_dx else None
    ws, rows = None, 0
    if deterministic():
        # one contiguous [K*C + K] result (dw | db) so that the fixed-order row reduction is a single launch
        flat = torch.empty((k * c + k,), dtype=torch.float32, device=x.device)
        dw, db_all = flat[:k * c].view(k, c), flat[k * c:]
        rows = 4 * sm_count(x.device)
        ws = torch.empty((rows, k * c + k), dtype=torch.float32, device=x.device)
        call("msp_final_conv_act_bwd", _p(x), n, h, w, c, cs, _p(w2d), k, act, _p(prob), _p(dprob), _p(dx), c,
             _p(dw), _p(db_all), _p(ws), rows, _stream())
        return dx, dw, (db_all if has_bias else None)
    dw = torch.empty((k, c), dtype=torch.float32, device=x.device)
    db = torch.empty((k,), dtype=torch.float32, device=x.device) if has_bias else None
    call("msp_final_conv_act_bwd", _p(x), n, h, w, c, cs, _p(w2d), k, act, _p(prob), _p(dprob), _p(dx), c,
         _p(dw), _p(db), None, 0, _stream())
    return dx, dw, db


def dice_sums(prob, mask, two_class, label_offset, batchwise):
    n, cp = prob.shape[0], prob.shape[1]
    hw = prob[0, 0].numel()
    ceff = 2 if two_class else cp
    g = 1 if batchwise else n
    sums = torch.empty((g, ceff, 3), dtype=torch.float64, device=prob.device)
    call("msp_dice_sums", _p(prob), _p(mask), n, cp, hw, int(two_class), int(label_offset), int(batchwise),
         _p(sums), _stream())
    return sums


def dice_finalize(sums, class_start, eps):
    g, ceff, _ = sums.shape
    coef = torch.empty((g, ceff, 2), dtype=torch.float32, device=sums.device)
    loss = torch.empty((), dtype=torch.float32, device=sums.device)
    call("msp_dice_finalize", _p(sums), g, ceff, int(class_start), float(eps), _p(coef), _p(loss), _stream())
    return loss, coef


def dice_bwd(prob, mask, two_class, label_offset, batchwise, coef, gscale=1.0, gscale_dev=None):
    n, cp = prob.shape[0], prob.shape[1]
    hw = prob[0, 0].numel()
    dprob = torch.empty_like(prob)
    call("msp_dice_bwd", _p(prob), _p(mask), n, cp, hw, int(two_class), int(label_offset), int(batchwise),
         _p(coef), float(gscale), _p(gscale_dev), _p(dprob), _stream())
    return dprob


def sum_to_mean(loss_sum, scale):
    out = torch.empty((), dtype=torch.float32, device=loss_sum.device)
    call("msp_scale_to_float", _p(loss_sum), float(scale), _p(out), _stream())
    return out


def ce_prob(prob, label, smooth, gscale=1.0, gscale_dev=None, want_loss=True, want_grad=False):
    n, c = prob.shape[0], prob.shape[1]
    hw = prob[0, 0].numel()
    ls = torch.empty((1,), dtype=torch.float64, device=prob.device) if want_loss else None
    dprob = torch.empty_like(prob) if want_grad else None
    call("msp_ce_prob_fwd_bwd", _p(prob), _p(label), n, c, hw, float(smooth), float(gscale), _p(gscale_dev),
         _p(ls), _p(dprob), _stream())
    return ls, dprob


def bce(prob, target, clamp_log, gscale=1.0, gscale_dev=None, want_loss=True, want_grad=False):
    ls = torch.empty((1,), dtype=torch.float64, device=prob.device) if want_loss else None
    dprob = torch.empty_like(prob) if want_grad else None
    call("msp_bce_fwd_bwd", _p(prob), _p(target), prob.numel(), int(clamp_log), float(gscale),
         _p(gscale_dev), _p(ls), _p(dprob), _stream())
    return ls, dprob


def softmax_ce(logits, label, smooth, gscale=1.0, gscale_dev=None, want_loss=True, want_grad=False):
    n, c = logits.shape
    ls = torch.empty((1,), dtype=torch.float64, device=logits.device) if want_loss else None
    dl = torch.empty_like(logits) if want_grad else None
    call("msp_softmax_ce_fwd_bwd", _p(logits), _p(label), n, c, float(smooth), float(gscale),
         _p(gscale_dev), _p(ls), _p(dl), _stream())
    return ls, dl


def softmax_ce_soft(logits, target, smooth, gscale=1.0, gscale_dev=None, want_loss=True, want_grad=False):
    """(N, C) logits against (N, C) class-probability targets."""
    n, c = logits.shape
    ls = torch.empty((1,), dtype=torch.float64, device=logits.device) if want_loss else None
    dl = torch.empty_like(logits) if want_grad else None
    call("msp_softmax_ce_soft_fwd_bwd", _p(logits), _p(target), n, c, float(smooth), float(gscale),
         _p(gscale_dev), _p(ls), _p(dl), _stream())
    return ls, dl


def softmax_ce_spatial(logits, label, smooth, gscale=1.0, gscale_dev=None, want_loss=True, want_grad=False):
    """(N, C, *spatial) logits against (N, prod(spatial)) class indices."""
    n, c = logits.shape[0], logits.shape[1]
    hw = logits[0, 0].numel()
    ls = torch.empty((1,), dtype=torch.float64, device=logits.device) if want_loss else None
    dl = torch.empty_like(logits) if want_grad else None
    call("msp_softmax_ce_spatial_fwd_bwd", _p(logits), _p(label), n, c, hw, float(smooth), float(gscale),
         _p(gscale_dev), _p(ls), _p(dl), _stream())
    return ls, dl


# ------------------------------------------------------------------------------------------------
# metrics / robustness
# ------------------------------------------------------------------------------------------------
def confusion_binary(pred, target, thr, per_channel):
    """pred fp32 (N, C, *spatial); target same shape (int64 or fp32).  -> int64 [6] or [C, 6] =
    TP, TN, FP, FN, positives, NaN targets."""
    n, c = pred.shape[0], pred.shape[1]
    hw = pred[0, 0].numel() if n > 0 else 0
    out = torch.empty((c, 6) if per_channel else (6,), dtype=torch.int64, device=pred.device)
    is_float = target.dtype == torch.float32
    call("msp_confusion_binary", _p(pred), _p(target), int(is_float), n, c, hw, float(thr), int(per_channel),
         _p(out), _stream())
    return out


def confusion_multiclass(pred, target, onehot):
    n, c = pred.shape[0], pred.shape[1]
    hw = pred[0, 0].numel() if n > 0 else 0
    cm = torch.empty((c, c), dtype=torch.int64, device=pred.device)
    call("msp_confusion_multiclass", _p(pred), _p(target), int(onehot), n, c, hw, _p(cm), _stream())
    return cm


def topk_hits(pred, label, k):
    n, c = pred.shape[0], pred.shape[1]
    hw = pred[0, 0].numel() if n > 0 else 0
    hits = torch.empty((1,), dtype=torch.int64, device=pred.device)
    call("msp_topk_hits", _p(pred), _p(label), n, c, hw, k, _p(hits), _stream())
    return hits


def rowpair_distances(q, k, pooled_hw=0):
    """q, k fp32 [N, D] (or [N, C, hw] with pooled_hw = hw).  -> fp32 [6, N]."""
    n = q.shape[0]
    d = q.shape[1] if pooled_hw > 1 else q[0].numel()
    out = torch.empty((6, n), dtype=torch.float32, device=q.device)
    call("msp_rowpair_distances", _p(q), _p(k), n, d, int(pooled_hw), _p(out), _stream())
    return out


def triplet_hinge(dist, margins):
    n = dist.shape[1]
    m = margins.numel()
    out = torch.empty((m, 3, n), dtype=torch.float32, device=dist.device)
    call("msp_triplet_hinge", _p(dist), n, _p(margins), m, _p(out), _stream())
    return out


launch_count = _lib.launch_count
