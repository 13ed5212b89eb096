"""Drop-in conversion of the reference's nn.Modules to the B200 path.

`convert(model)` leaves the module tree, its Parameters / buffers and therefore `state_dict()`,
`parameters()`, `.train()/.eval()`, `.requires_grad_()` and the optimizer untouched, and replaces only
what happens inside `forward`: an interpreter walks the reference-structured module tree and issues the
fused sm_100a kernels of libmsp_b200.so on bf16 NHWC activations.

Recognised structures (by class name and attribute layout, so that the same code drives both the real
reference modules and the oracle restatement in oracle/ref_models.py):
  DeepResNet / ResBlock / BottleNeckBlock / BasicBlock / DropPath   classification/models.py:9-325
  UNet / UNet_encoder / UNet_decoder                                 segmentation/models/unet_models.py:39-688
  ConvBlock / UpConvBlock / AttentionBlock / ConcatBlock             segmentation/models/blocks.py:419-635
  Model wrapper (`.model`)                                           model/model.py:18-75
  FeedForwardModel (sequential `.layers`: encoder + pooled linear head)  config/pretraining/*/*.yaml
  nn.Conv2d / ConvTranspose2d / BatchNorm2d / ReLU / Sigmoid / MaxPool2d / Upsample(nearest | bilinear, x2) /
  Sequential / Identity
Anything else raises UnsupportedModule: there is no PyTorch fallback on this path.
"""
from __future__ import annotations

import os
import types
from typing import List, Optional, Tuple

import torch
from torch import nn

from . import functional as Fn
from . import ops


class UnsupportedModule(NotImplementedError):
    pass


# Folded up-convolution (functional._UpConv2x): 0 = never (materialise the x4 tensor like the reference), 2 = always,
# 1 (default) = when the layer is large enough to pay for its extra launches: the fold runs 9/16 of the FLOPs on a 4x
# smaller operand but as 4 + 1 + 4 kernels instead of 1 + 1 + 1 (+ 2 up-sample kernels); at the R50 U-Net's batch 24 every
# up-conv is a 15-30 us launch for a few us of work and the fold measured 3 % SLOWER over the step (2 033 vs 2 094 img/s,
# profiles/r02_experiments.txt), on the 1024^2 basic U-Net the up-convs are the largest tensors of the network.
_UPCONV_FOLD = int(os.environ.get("MSP_UPCONV_FOLD", "1"))
_BRANCH_STREAMS = os.environ.get("MSP_BRANCH_STREAMS", "1") != "0"
_BRANCH_MAX_PIXELS = int(os.environ.get("MSP_BRANCH_MAX_PIXELS", "1000000000000"))   # (sweep: always on is best)
_UPCONV_FOLD_MIN_FLOPS = 1.5e11      # unfused forward FLOPs (2 * 4NHW * K * 4C) above which the fold is used in mode 1


class ExecContext:
    """Per-model execution options: `group` = process group for SyncBN statistics (None = local).
    During a CUDA-graph capture (graphs.GraphedStep) the DropPath per-sample scales live in static device buffers
    that are refilled from the CPU generator before every replay."""

    def __init__(self, group=None):
        self.group = group
        self._static = None      # list of (DropPath module, n, device tensor) in forward order while capturing
        self._replay = []
        self.counters = []       # num_batches_tracked buffers touched by the running forward
        self.pack_cache = ops.WeightPackCache()   # bf16 operand copies of the conv weights, one repack kernel per step
        self._branch = {}        # device index -> (stream, stream) for independent decoder chains

    def branch_streams(self, device, pixels: int):
        """Two side streams for the independent chains of an attention decoder level (run_unet_decoder), or None:
        off with MSP_BRANCH_STREAMS=0, on the CPU, while bench.py's per-kernel instrumentation attributes device time
        to convolution calls in stream order, and — single process — on levels large enough to fill the GPU by
        themselves (`pixels` = N * H * W of the level)."""
        if not _BRANCH_STREAMS or torch.device(device).type != "cuda" or ops.conv_timeline_active():
            return None
        if self.group is None and pixels > _BRANCH_MAX_PIXELS:
            return None
        key = torch.device(device).index
        st = self._branch.get(key)
        if st is None:
            st = (torch.cuda.Stream(device=device), torch.cuda.Stream(device=device))
            self._branch[key] = st
        return st

    def begin_forward(self):
        self.pack_cache.begin_step()
        ops.set_active_pack_cache(self.pack_cache)

    def stats_for(self, bn):
        """Persistent, self-cleaning [2, C] accumulator of the conv epilogue's BatchNorm sums (zeroed once here, then
        by msp_bn_finalize after every use)."""
        buf = getattr(bn, "_msp_stats", None)
        dev = bn.weight.device if bn.weight is not None else bn.running_mean.device
        if buf is None or buf.device != dev or buf.shape[-1] != bn.num_features or \
                (buf.dim() == 3) != ops.deterministic():
            buf = ops.new_stats(bn.num_features, dev)      # [2, C], or [SMs, 2, C] rows in deterministic mode
            bn._msp_stats = buf
        return buf

    def finish_forward(self):
        ops.set_active_pack_cache(None)
        self.pack_cache.end_step()
        if not torch.cuda.is_current_stream_capturing():
            self.pack_cache.build_table()
        if self.counters:
            torch._foreach_add_(self.counters, 1)
            self.counters = []

    def begin_static_droppath(self):
        self._static = []
        self._flat_dev, self._flat_used = None, 0

    def end_static_droppath(self):
        self._replay, self._static = self._static or [], None
        # per replay ONE pinned -> device copy of all the masks (15 pageable copies of 24 floats were 0.58 ms of the
        # 10.5 ms R50 U-Net step); a small ring of pinned buffers because the host runs several replays ahead
        self._flat_host = [torch.empty(max(self._flat_used, 1), dtype=torch.float32).pin_memory() for _ in range(4)] \
            if self._replay else []
        self._flat_event, self._flat_turn = [None] * 4, 0

    def droppath_scale(self, dp, n, device):
        """classification/models.py:320-325: Bernoulli(keep) per sample on the CPU generator in training (no 1/keep
        rescale), keep_prob in eval."""
        if dp.training:
            s = torch.bernoulli(dp.keep_prob * torch.ones((n, 1, 1, 1))).reshape(n)
        else:
            s = torch.full((n,), float(dp.keep_prob))
        if self._static is not None:
            if self._flat_dev is None:
                self._flat_dev = torch.empty(1 << 16, dtype=torch.float32, device=device)
            if self._flat_used + n > self._flat_dev.numel() or self._flat_dev.device != torch.device(device):
                raise UnsupportedModule("DropPath under graph capture: more than 65536 mask elements per step")
            off = self._flat_used
            self._flat_used += (n + 3) // 4 * 4          # 16-byte aligned slices
            self._static.append((dp, n, off))
            return self._flat_dev[off:off + n]          # filled by refresh_droppath() before each replay
        return s.to(device=device, dtype=torch.float32, non_blocking=True)

    def refresh_droppath(self):
        if not self._replay:
            return
        k = self._flat_turn
        self._flat_turn = (k + 1) % len(self._flat_host)
        if self._flat_event[k] is not None:
            self._flat_event[k].synchronize()           # the copy that last read this pinned buffer has run
        host = self._flat_host[k]
        for dp, n, off in self._replay:                 # the reference's draws, in its order, on the CPU generator
            if dp.training:
                host[off:off + n] = torch.bernoulli(dp.keep_prob * torch.ones((n, 1, 1, 1))).reshape(n)
            else:
                host[off:off + n] = float(dp.keep_prob)
        self._flat_dev[:host.numel()].copy_(host, non_blocking=True)
        ev = torch.cuda.Event()
        ev.record()
        self._flat_event[k] = ev


def context_of(model):
    return getattr(_unwrap(model), "_msp_ctx", None)


class RawInput:
    """The network input as the reference hands it over (fp32 NCHW), not yet converted: the first
    convolution consumes it directly through the row-window path (functional.input_conv2d)."""

    def __init__(self, t: torch.Tensor):
        self.t = t


def _as_nhwc(x):
    return Fn.to_nhwc(x.t) if isinstance(x, RawInput) else x


def _name(m) -> str:
    return type(m).__name__


def _unwrap(m):
    """model.Model wrapper -> the wrapped module (model/model.py:62)."""
    while _name(m) == "Model" and hasattr(m, "model"):
        m = m.model
    return m


def _pad_of(conv: nn.Conv2d):
    p = conv.padding
    if isinstance(p, str):
        if p == "same":
            return "same"
        if p == "valid":
            return 0
        raise UnsupportedModule(f"conv padding {p!r}")
    if p[0] != p[1]:
        return (int(p[0]), int(p[1]))
    return int(p[0])


def _check_conv(conv: nn.Conv2d):
    if conv.groups != 1 or tuple(conv.dilation) != (1, 1) or conv.stride[0] != conv.stride[1] \
            or conv.kernel_size[0] != conv.kernel_size[1] or conv.padding_mode != "zeros":
        raise UnsupportedModule(f"convolution {conv} (groups/dilation/anisotropic) is outside the B200 path")
    if conv.stride[0] not in (1, 2):
        raise UnsupportedModule(f"convolution stride {conv.stride}")


_ACT_OF = {"ReLU": ops.ACT_RELU, "Sigmoid": ops.ACT_SIGMOID, "Identity": ops.ACT_NONE}


def _is_noop(m) -> bool:
    n = _name(m)
    if n == "Identity":
        return True
    if n in ("Dropout", "Dropout2d"):
        if m.p == 0 or not m.training:
            return True
        raise UnsupportedModule("training-mode Dropout2d is not on the B200 path")
    return False


def conv_bn_act(ctx: ExecContext, x, conv: nn.Conv2d, bn: Optional[nn.BatchNorm2d] = None, act: int = 0,
                residual=None, r_stride: int = 1, sample_scale=None, link_in=None, link_out=None, out=None,
                conv_stream=None):
    """Conv2d [-> BatchNorm2d] [-> ReLU | Sigmoid], with optional fused residual / per-sample scale.
    `out` (conv without BatchNorm only): channel slice of a wider NHWC buffer the epilogue writes into (zero-copy
    torch.cat).  `conv_stream`: a side stream that already waits for x; the convolution is issued there and the current
    stream joins it before the BatchNorm (the statistics exchange of a data-parallel run stays on the current stream)."""
    _check_conv(conv)
    stride, padding = conv.stride[0], _pad_of(conv)
    if isinstance(x, RawInput):
        if out is not None:
            raise UnsupportedModule("the network's first convolution cannot write into a concat slice")
        conv2d = lambda _x, *a, out=None, **k: Fn.input_conv2d(x.t, *a, **k)
    elif link_in is not None:
        conv2d = lambda *a, **k: Fn.conv2d(*a, link=link_in, **k)
    else:
        conv2d = Fn.conv2d
    if bn is None:
        if residual is not None or sample_scale is not None or act == ops.ACT_SIGMOID:
            raise UnsupportedModule("residual / sigmoid epilogue without BatchNorm")
        y, _ = conv2d(x, conv.weight, conv.bias, stride, padding, relu=(act == ops.ACT_RELU), out=out)
        return y
    if out is not None:
        raise UnsupportedModule("conv_bn_act(out=...) with BatchNorm")
    training = bn.training or bn.running_mean is None
    # the conv adds its bias in the epilogue, but the bias *gradient* is produced by the BatchNorm node
    bias = conv.bias.detach() if conv.bias is not None else None
    if conv_stream is not None:
        cur = torch.cuda.current_stream()
        with torch.cuda.stream(conv_stream):
            y, stats = conv2d(x, conv.weight, bias, stride, padding, relu=False,
                              want_stats=ctx.stats_for(bn) if training else False)
        cur.wait_stream(conv_stream)
        y.record_stream(cur)            # allocated on the side stream's pool, consumed here
    else:
        y, stats = conv2d(x, conv.weight, bias, stride, padding, relu=False,
                          want_stats=ctx.stats_for(bn) if training else False)
    return Fn.bn_act(y, stats, bn, act=act, residual=residual, r_stride=r_stride,
                     sample_scale=sample_scale, group=ctx.group, conv_bias=conv.bias, persistent_stats=True,
                     counters=ctx.counters, link=link_out)


def run_sequence(ctx: ExecContext, mods: List[nn.Module], x):
    """Greedy fusion over a flat list of leaf modules."""
    i, n = 0, len(mods)
    while i < n:
        m = _unwrap(mods[i])
        nm = _name(m)
        if _is_noop(m):
            i += 1
            continue
        if nm not in ("Conv2d", "Sequential", "ConvBlock"):
            x = _as_nhwc(x)
        if nm == "Conv2d":
            bn, act, j = None, ops.ACT_NONE, i + 1
            if j < n and _name(_unwrap(mods[j])) == "BatchNorm2d":
                bn = _unwrap(mods[j])
                j += 1
            if j < n and _name(_unwrap(mods[j])) in ("ReLU", "Sigmoid"):
                act = _ACT_OF[_name(_unwrap(mods[j]))]
                j += 1
            x = conv_bn_act(ctx, x, m, bn, act)
            i = j
        elif nm == "ConvTranspose2d":
            # north star "conv/transposed-conv layers" (the reference's own up-sampling is nearest + Conv2d): the data
            # gradient kernel of the convolution it transposes, bias / ReLU in the epilogue
            if m.groups != 1 or tuple(m.dilation) != (1, 1) or tuple(m.output_padding) != (0, 0) or \
                    m.stride[0] != m.stride[1] or m.stride[0] not in (1, 2) or m.kernel_size[0] != m.kernel_size[1] or \
                    m.padding[0] != m.padding[1] or m.kernel_size[0] < m.stride[0] or m.out_channels % 8 or \
                    m.in_channels % 8:
                raise UnsupportedModule(f"{m} is outside the B200 path")
            j, relu = i + 1, False
            if j < n and _name(_unwrap(mods[j])) == "ReLU":
                relu, j = True, j + 1
            x = Fn.conv_transpose2d(x, m.weight, m.bias, m.stride[0], m.padding[0], relu=relu)
            i = j
        elif nm == "MaxPool2d":
            k = m.kernel_size if isinstance(m.kernel_size, int) else m.kernel_size[0]
            s = m.stride if isinstance(m.stride, int) else m.stride[0]
            p = m.padding if isinstance(m.padding, int) else m.padding[0]
            if m.dilation not in (1, (1, 1)) or m.ceil_mode or m.return_indices:
                raise UnsupportedModule(f"{m}")
            x = Fn.maxpool2d(x, k, s or k, p)
            i += 1
        elif nm == "Upsample":
            sf = m.scale_factor if not isinstance(m.scale_factor, tuple) else m.scale_factor[0]
            if sf is None or float(sf) != 2.0 or (isinstance(m.scale_factor, tuple) and len(set(m.scale_factor)) != 1):
                raise UnsupportedModule(f"{m}: only x2 up-sampling is on the B200 path")
            if m.mode == "nearest":
                x = Fn.upsample2x(x)
            elif m.mode == "bilinear" and not m.align_corners:
                x = Fn.upsample_bilinear2x(x)
            else:
                raise UnsupportedModule(f"{m}: nearest and bilinear (align_corners=False) x2 up-sampling only")
            i += 1
        elif nm == "Sequential":
            x = run_sequence(ctx, list(m.children()), x)
            i += 1
        else:
            x = run_module(ctx, m, x)
            i += 1
    return x


# ------------------------------------------------------------------------------------------------
# classification/models.py
# ------------------------------------------------------------------------------------------------
def _drop_path_scale(ctx, block, n: int, device) -> Optional[torch.Tensor]:
    dp = getattr(block, "drop_path", None)
    if dp is None or _name(dp) == "Identity":
        return None
    if _name(dp) != "DropPath" or not hasattr(dp, "keep_prob"):
        raise UnsupportedModule(f"drop path module {dp}")
    return ctx.droppath_scale(dp, n, device)


def run_res_unit(ctx: ExecContext, blk, x):
    """BottleNeckBlock.forward (models.py:277-290) / BasicBlock.forward (:203-212): the last BN, the
    DropPath multiply, the strided zero-filled shortcut, the add and the ReLU are ONE kernel."""
    bottleneck = hasattr(blk, "conv3")
    convs = [blk.conv1, blk.conv2] + ([blk.conv3] if bottleneck else [])
    bns = [blk.bn1, blk.bn2] + ([blk.bn3] if bottleneck else [])
    r_stride = max(c.stride[0] for c in convs)
    if convs[-1].out_channels < convs[0].in_channels:
        raise UnsupportedModule("residual block that narrows its input")
    scale = _drop_path_scale(ctx, blk, x.shape[0], x.device)
    # the shortcut gradient goes straight from the last BatchNorm's backward into the first conv's dgrad epilogue
    link = Fn.ResidualLink() if torch.is_grad_enabled() and x.requires_grad else None
    y = x
    for i, (conv, bn) in enumerate(zip(convs[:-1], bns[:-1])):
        y = conv_bn_act(ctx, y, conv, bn, ops.ACT_RELU, link_in=link if i == 0 else None)
    return conv_bn_act(ctx, y, convs[-1], bns[-1], ops.ACT_RELU, residual=x, r_stride=r_stride,
                       sample_scale=scale, link_out=link)


def run_deep_resnet(ctx: ExecContext, m, x, return_skip_vals: bool = False):
    """DeepResNet.forward (models.py:89-103) on NHWC activations; returns NHWC tensors."""
    if getattr(m, "version", "v1") != "v1":
        raise UnsupportedModule("DeepResNet v2 (pre-activation) is not on the B200 path")
    y = run_sequence(ctx, list(m.stem.children()), x)
    skips = [y]
    y = run_sequence(ctx, [m.max_pool], y)
    for level in m.levels:
        for blk in level:
            if _name(blk) not in ("BottleNeckBlock", "BasicBlock"):
                raise UnsupportedModule(f"residual unit {_name(blk)}")
            y = run_res_unit(ctx, blk, y)
        skips.append(y)
    cls = m.classifier
    if _name(cls) != "Identity":
        y = run_head(ctx, list(cls.children()), y)
    return (y, skips[:-1]) if return_skip_vals else y


def run_head(ctx: ExecContext, mods: List[nn.Module], y):
    """AdaptiveAvgPool2d(1) -> Flatten -> Linear (classification/models.py:73-77; the tail of the pretraining YAMLs'
    layer list, config/pretraining/resnet50/simple.yaml:27-33) on an NHWC feature map: the pooled vector stays an
    (N, 1, 1, C) activation and the Linear layer is a 1x1 convolution on the same tap-GEMM kernel."""
    mods = [_unwrap(c) for c in mods]
    if [_name(c) for c in mods] != ["AdaptiveAvgPool2d", "Flatten", "Linear"]:
        raise UnsupportedModule(f"classifier head {[_name(c) for c in mods]}")
    pool, flat, lin = mods
    if pool.output_size not in (1, (1, 1)) or (flat.start_dim, flat.end_dim) != (1, -1):
        raise UnsupportedModule(f"classifier head {pool} / {flat}")
    if lin.out_features % 8 or lin.in_features != y.shape[3]:
        raise UnsupportedModule(f"{lin}: the head's GEMM needs out_features % 8 == 0 and in_features == the "
                                f"encoder's {y.shape[3]} channels on the B200 path")
    y = Fn.global_avgpool(y)
    w4 = lin.weight.view(lin.out_features, lin.in_features, 1, 1)
    y, _ = Fn.conv2d(y, w4, lin.bias, 1, 0)
    return y


def run_feed_forward(ctx: ExecContext, m, x):
    """The SEQUENTIAL compound model of the pretraining YAMLs (`model.FeedForwardModel: {layers: [...]}`,
    config/pretraining/resnet50/simple.yaml:23-33; the class that produced the published `layers.0.` encoder
    checkpoints, segmentation/models/unet_models.py:570-571, is not in the reference checkout — SURVEY.md App. C):
    layers[0] = the encoder, the remaining layers = pooling / flatten / linear head."""
    layers = [_unwrap(l) for l in m.layers]
    if not layers or _name(layers[0]) != "DeepResNet":
        raise UnsupportedModule(f"sequential model starting with {_name(layers[0]) if layers else None}")
    y = run_deep_resnet(ctx, layers[0], x)
    if _name(layers[0].classifier) != "Identity":
        if len(layers) > 1:
            raise UnsupportedModule("layers after an encoder that already carries a head")
        return y
    if len(layers) > 1:
        y = run_head(ctx, layers[1:], y)
    return y


# ------------------------------------------------------------------------------------------------
# segmentation/models
# ------------------------------------------------------------------------------------------------
def run_conv_block(ctx, m, x):
    return run_sequence(ctx, list(m.block.children()), x)


def run_upconv_block(ctx, m, x, out=None):
    """UpConvBlock.forward (blocks.py:537-539): nearest x2 -> Conv2d(k=2, 'same') -> ReLU.  `out`: the leading channel
    slice of the level's concat buffer — the conv epilogue writes x_up where torch.cat (blocks.py:628,635) would copy it."""
    mods = [_unwrap(c) for c in m.convup.children()]
    if [_name(c) for c in mods] == ["Upsample", "Conv2d", "ReLU"]:
        up, conv = mods[0], mods[1]
        sf = up.scale_factor if not isinstance(up.scale_factor, tuple) else up.scale_factor[0]
        big = isinstance(x, torch.Tensor) and \
            2.0 * 4 * x.shape[0] * x.shape[1] * x.shape[2] * conv.out_channels * 4 * conv.in_channels >= _UPCONV_FOLD_MIN_FLOPS
        foldable = ((_UPCONV_FOLD == 2 or (_UPCONV_FOLD == 1 and big)) and up.mode == "nearest"
                    and sf is not None and float(sf) == 2.0
                    and tuple(conv.kernel_size) == (2, 2) and tuple(conv.stride) == (1, 1) and conv.padding == "same"
                    and conv.groups == 1 and tuple(conv.dilation) == (1, 1) and conv.out_channels % 8 == 0
                    and not isinstance(x, RawInput))
        if foldable:
            # nearest x2 -> 2x2 'same' conv -> ReLU folded onto the low-res input (functional._UpConv2x)
            return Fn.upconv2x(x, conv.weight, conv.bias, relu=True, out=out)
        if out is not None:
            xu = run_sequence(ctx, [up], x)
            return conv_bn_act(ctx, xu, conv, None, ops.ACT_RELU, out=out)
    return run_sequence(ctx, mods, x)


def _upconv_out_channels(m) -> Optional[int]:
    mods = [_unwrap(c) for c in m.convup.children()] if hasattr(m, "convup") else []
    if [_name(c) for c in mods] == ["Upsample", "Conv2d", "ReLU"] and mods[1].out_channels % 8 == 0:
        return mods[1].out_channels
    return None


def run_attention_block(ctx, m, x, x_up, skip, buf=None, ws_stream=None, up_stream=None):
    """AttentionBlock.forward (blocks.py:620-628).  `buf`: the concat buffer whose leading slice already IS x_up.
    `ws_stream` / `up_stream` (ExecContext.branch_streams): the W_s convolution of the skip tensor is issued on the first,
    the up-convolution that produces x_up is still running on the second — the three chains of a decoder level
    (up-conv; gating signal g -> W_g; W_s) are independent until the gate and fill the GPU side by side at small batch."""
    if ws_stream is not None:
        ws_stream.wait_stream(torch.cuda.current_stream())      # skip is complete
    g = run_module(ctx, _unwrap(m.gs_block), x)
    wg_conv, wg_bn = list(m.W_g.children())
    ws_conv, ws_bn = list(m.W_s.children())
    psi = list(m.psi.children())
    if [_name(c) for c in psi] != ["Conv2d", "BatchNorm2d", "Sigmoid"]:
        raise UnsupportedModule(f"attention psi {m.psi}")
    g1 = conv_bn_act(ctx, g, wg_conv, wg_bn, ops.ACT_NONE)
    # relu(BN(W_s(skip)) + g1): the add + ReLU ride on the BN-apply kernel
    p = conv_bn_act(ctx, skip, ws_conv, ws_bn, ops.ACT_RELU, residual=g1, conv_stream=ws_stream)
    p = conv_bn_act(ctx, p, psi[0], psi[1], ops.ACT_SIGMOID)
    if up_stream is not None:
        torch.cuda.current_stream().wait_stream(up_stream)       # x_up is complete
    return Fn.gate_concat(x_up, skip, p, buf=buf)


def run_unet_encoder(ctx, m, x, return_skip_vals=False):
    """UNet_encoder.forward (unet_models.py:200-236)."""
    if getattr(m, "res_con", False) or getattr(m, "layer_scale", False):
        raise UnsupportedModule("U-Net residual connections / layer scaling are not on the B200 path")
    skips = []
    x = run_sequence(ctx, [m.first_block], x)
    for unit in m.down_layers:
        for j in range(m.width):
            x = run_module(ctx, _unwrap(unit[f"conv{j}"]), x)
        skips.append(x)
        if "downsampl" in unit:
            x = run_sequence(ctx, [unit["downsampl"]], x)
    for j in range(m.width):
        x = run_module(ctx, _unwrap(m.bottom_block[f"conv{j}"]), x)
    return (x, skips) if return_skip_vals else x


def run_unet_decoder(ctx, m, x, skips: list, final_act=None):
    """UNet_decoder.forward (unet_models.py:367-390) + the final activation of UNet.forward (:685-686).
    Returns the fp32 NCHW prediction."""
    if getattr(m, "res_con", False) or getattr(m, "layer_scale", False):
        raise UnsupportedModule("U-Net residual connections / layer scaling are not on the B200 path")
    skips = list(skips)
    for i, unit in enumerate(m.up_layers):
        up = _unwrap(unit["upsampl"])
        if i < m.skip_con_nr:
            skip = skips.pop()
            mix = _unwrap(unit["mixing"])
            if _name(mix) not in ("ConcatBlock", "AttentionBlock"):
                raise UnsupportedModule(f"mixing block {_name(mix)}")
            # zero-copy torch.cat((x_up, .), dim=1) (blocks.py:628,635): the up-conv's epilogue writes x_up straight into
            # the leading channel slice of the concat buffer, the gate product / the skip goes into the trailing slice
            buf, ca = None, _upconv_out_channels(up) if _name(up) == "UpConvBlock" else None
            streams = ctx.branch_streams(skip.device, skip.shape[0] * skip.shape[1] * skip.shape[2]) \
                if (_name(mix) == "AttentionBlock" and ca is not None and skip.shape[3] % 8 == 0) else None
            if ca is not None and skip.shape[3] % 8 == 0:
                buf = ops.new_act(skip.shape[0], skip.shape[1], skip.shape[2], ca + skip.shape[3], skip.device)
                if streams is not None:
                    streams[0].wait_stream(torch.cuda.current_stream())      # x (and buf's allocation) are complete
                    with torch.cuda.stream(streams[0]):
                        x_up = run_upconv_block(ctx, up, x, out=buf[..., :ca])
                else:
                    x_up = run_upconv_block(ctx, up, x, out=buf[..., :ca])
            else:
                x_up = run_module(ctx, up, x)
            if _name(mix) == "ConcatBlock":
                x = Fn.concat(x_up, skip, buf=buf)
            elif streams is not None:
                x = run_attention_block(ctx, mix, x, x_up, skip, buf=buf, ws_stream=streams[1], up_stream=streams[0])
            else:
                x = run_attention_block(ctx, mix, x, x_up, skip, buf=buf)
        else:
            x = run_module(ctx, up, x)
        for j in range(m.width):
            x = run_module(ctx, _unwrap(unit[f"conv{j}"]), x)
    fb = _unwrap(m.final_block)
    act_name = "none" if final_act is None else _name(final_act)
    if _name(fb) == "Conv2d" and fb.kernel_size == (1, 1) and fb.out_channels <= 8 and \
            act_name in ("none", "Sigmoid", "Softmax", "Identity"):
        if act_name == "Softmax" and final_act.dim != 1:
            raise UnsupportedModule("softmax over a dimension other than 1")
        act = {"none": 0, "Identity": 0, "Sigmoid": 1, "Softmax": 2}[act_name]
        return Fn.final_conv_act(x, fb.weight, fb.bias, act)
    raise UnsupportedModule(f"final block {fb} with activation {final_act}")


def run_unet(ctx, m, x):
    """UNet.forward (unet_models.py:681-688)."""
    enc = _unwrap(m.encoder)
    y, skips = run_module(ctx, enc, x, return_skip_vals=True)
    return run_unet_decoder(ctx, m.decoder, y, skips, m.final_act)


_RUNNERS = {
    "DeepResNet": run_deep_resnet,
    "UNet_encoder": run_unet_encoder,
    "ConvBlock": run_conv_block,
    "UpConvBlock": run_upconv_block,
    "BottleNeckBlock": run_res_unit,
    "BasicBlock": run_res_unit,
    "UNet": run_unet,
}


def run_module(ctx, m, x, **kw):
    m = _unwrap(m)
    nm = _name(m)
    if nm in _RUNNERS:
        return _RUNNERS[nm](ctx, m, x, **kw)
    if nm in ("Conv2d", "ConvTranspose2d", "MaxPool2d", "Upsample", "Sequential", "Identity"):
        return run_sequence(ctx, [m], x)
    raise UnsupportedModule(f"module {nm} has no B200 implementation (no PyTorch fallback on this path)")


# ------------------------------------------------------------------------------------------------
# public entry point
# ------------------------------------------------------------------------------------------------
def _require_cuda(x):
    if not x.is_cuda:
        raise RuntimeError("medsegpretrainimagenet_b200: the converted model runs on CUDA (sm_100a) "
                           "tensors only; there is no CPU fallback")


def _forward_deep_resnet(self, x, return_skip_vals=False, *args, **kwargs):
    _require_cuda(x)
    ctx = self._msp_ctx
    ctx.begin_forward()
    out = run_deep_resnet(ctx, self, RawInput(x), return_skip_vals=return_skip_vals)
    ctx.finish_forward()
    y, skips = out if return_skip_vals else (out, None)
    if _name(self.classifier) != "Identity":
        y = Fn.to_nchw(y).flatten(1)
    else:
        y = Fn.to_nchw(y)
    if return_skip_vals:
        return y, [Fn.to_nchw(s) for s in skips]
    return y


def _forward_unet_encoder(self, x, return_skip_vals=False):
    _require_cuda(x)
    self._msp_ctx.begin_forward()
    out = run_unet_encoder(self._msp_ctx, self, RawInput(x), return_skip_vals=return_skip_vals)
    self._msp_ctx.finish_forward()
    if return_skip_vals:
        return Fn.to_nchw(out[0]), [Fn.to_nchw(s) for s in out[1]]
    return Fn.to_nchw(out)


def _forward_unet(self, x):
    _require_cuda(x)
    self._msp_ctx.begin_forward()
    out = run_unet(self._msp_ctx, self, RawInput(x))
    self._msp_ctx.finish_forward()
    return out


def _forward_feed_forward(self, x, *args, **kwargs):
    _require_cuda(x)
    ctx = self._msp_ctx
    ctx.begin_forward()
    y = run_feed_forward(ctx, self, RawInput(x))
    ctx.finish_forward()
    y = Fn.to_nchw(y)
    layers = [_unwrap(l) for l in self.layers]
    has_head = len(layers) > 1 or _name(layers[0].classifier) != "Identity"
    return y.flatten(1) if has_head else y


_TOP_LEVEL = {"DeepResNet": _forward_deep_resnet, "UNet_encoder": _forward_unet_encoder,
              "UNet": _forward_unet, "FeedForwardModel": _forward_feed_forward}


def convert(model: nn.Module, group=None) -> nn.Module:
    """Route `model`'s forward through the B200 kernels (in place; returns `model`).

    `model` may be the reference's `model.Model` wrapper or the bare network.  Parameters, buffers and
    the module tree are shared, so state dicts, optimizers and checkpoints are unaffected.
    `group`: torch.distributed process group whose ranks synchronise BatchNorm statistics."""
    inner = _unwrap(model)
    nm = _name(inner)
    if nm not in _TOP_LEVEL:
        raise UnsupportedModule(f"cannot convert top-level module {nm}")
    inner._msp_ctx = ExecContext(group)
    inner.forward = types.MethodType(_TOP_LEVEL[nm], inner)
    inner._msp_converted = True
    return model


def is_converted(model: nn.Module) -> bool:
    return bool(getattr(_unwrap(model), "_msp_converted", False))
