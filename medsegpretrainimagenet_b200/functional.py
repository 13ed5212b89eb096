"""torch.autograd.Function nodes over the C-ABI kernels.

Every node consumes / produces bf16 NHWC activations (see ops.py) except the two boundary nodes
(`to_nhwc`, `to_nchw`) and the segmentation head, which speak the reference's fp32 NCHW.  Parameters
stay the reference's fp32 `nn.Parameter`s: gradients are produced in their native OIHW fp32 layout so
`param.grad` is what `clip_grad_norm_` / the optimizer of train_model.py:93-107 expect.
"""
from __future__ import annotations

from typing import Optional

import torch
import torch.distributed as dist

from . import ops


def _peer(group):
    """The peer-memory communicator of `group` when enabled (parallel.enable_peer_allreduce), else None."""
    from .parallel import peer_allreduce_for
    return peer_allreduce_for(group)


def _grad_sink_ok(p) -> bool:
    """True when a kernel may add this parameter's gradient into `p.grad` itself (None: nothing to write)."""
    if p is None:
        return True
    g = p.grad
    return (p.is_leaf and g is not None and g.dtype == torch.float32 and g.is_contiguous() and g.shape == p.shape
            and g.device == p.device)


def _allreduce_sum(t: torch.Tensor, group) -> None:
    """SyncBN / global-Dice exchange: a small NCCL all-reduce, only when a process group is active."""
    from .parallel import allreduce_small_sum_
    allreduce_small_sum_(t, group)


# ------------------------------------------------------------------------------------------------
# boundary
# ------------------------------------------------------------------------------------------------
class _ToNHWC(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        ctx.c = x.shape[1]
        return ops.nchw_to_nhwc(x)

    @staticmethod
    def backward(ctx, dy):
        return ops.nhwc_to_nchw(dy.contiguous() if dy.stride(3) != 1 else dy, ctx.c)


class _ToNCHW(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, c):
        ctx.cpad = x.shape[3]
        return ops.nhwc_to_nchw(x, c)

    @staticmethod
    def backward(ctx, dy):
        return ops.nchw_to_nhwc(dy, ctx.cpad), None


def to_nhwc(x: torch.Tensor) -> torch.Tensor:
    return _ToNHWC.apply(x)


def to_nchw(x: torch.Tensor, c: Optional[int] = None) -> torch.Tensor:
    return _ToNCHW.apply(x, x.shape[3] if c is None else c)


# ------------------------------------------------------------------------------------------------
# convolution
# ------------------------------------------------------------------------------------------------
class ResidualLink:
    """Carries the shortcut gradient of a residual unit from its last BatchNorm's backward to its first
    convolution's backward, which ADDS its dgrad into that buffer in the kernel epilogue — instead of autograd
    summing two full-size gradient tensors with an extra elementwise kernel (and an extra bf16 rounding)."""
    __slots__ = ("dres",)

    def __init__(self):
        self.dres = None


def _stats_buffer(want_stats, k, device):
    """`want_stats` is False, True (fresh zeroed [2, K] accumulator) or a persistent accumulator that
    msp_bn_finalize(reset) leaves zeroed for the next step."""
    if want_stats is False or want_stats is None:
        return None
    if want_stats is True:
        return ops.new_stats(k, device)
    return want_stats


class _Conv(torch.autograd.Function):
    """y = conv2d(x, weight, bias) (+ReLU); optionally returns the fused per-channel (sum, sum of
    squares) of y as a second, non-differentiable output."""

    @staticmethod
    def forward(ctx, x, weight, bias, stride, padding, relu, want_stats, out, link):
        ctx.link = link
        ctx.set_materialize_grads(False)    # no zero-filled "gradient" of the statistics output per backward pass
        k, c_true, kh, kw = weight.shape
        n, h, w, c8 = x.shape
        ho, wo, pt, pl = ops.conv_out_size(h, w, kh, kw, stride, padding)
        need_dx = ctx.needs_input_grad[0]
        wf, wd = ops.packed_weights(weight, need_dgrad=need_dx)
        stats = _stats_buffer(want_stats, k, x.device)
        y = ops.conv_fprop(x, wf, bias.detach() if bias is not None else None, k, kh, kw, stride, pt, pl,
                           ho, wo, relu=relu, out=out, stats=stats, c_true=c_true)
        ctx.geom = (kh, kw, stride, pt, pl, c_true, relu, bias is not None)
        ctx.x_shape = tuple(x.shape)
        ctx.weight_param = weight if (weight.is_leaf and weight.requires_grad) else None
        ctx.save_for_backward(x, wd, y if relu else None)
        if stats is not None:
            ctx.mark_non_differentiable(stats)
            return y, stats
        return y, None

    @staticmethod
    def backward(ctx, dy, _dstats):
        if dy is None:
            return (None,) * 9
        kh, kw, stride, pt, pl, c_true, relu, has_bias = ctx.geom
        x, wd, y = ctx.saved_tensors
        if dy.stride(3) != 1:
            dy = dy.contiguous()
        if relu:
            dy = ops.relu_bwd(y, dy)
        dx = dw = db = None
        if ctx.needs_input_grad[0]:
            link = ctx.link
            if link is not None and link.dres is not None and tuple(link.dres.shape) == tuple(ctx.x_shape):
                # residual unit: dx = shortcut gradient (already in the buffer) + dgrad, added in the epilogue
                dx = ops.conv_dgrad(dy, wd, ctx.x_shape, kh, kw, stride, pt, pl, out=link.dres, accumulate=True,
                                    c_true=c_true)
                link.dres = None
            else:
                dx = ops.conv_dgrad(dy, wd, ctx.x_shape, kh, kw, stride, pt, pl, c_true=c_true)
        if ctx.needs_input_grad[1]:
            # a leaf Parameter's gradient goes straight into `weight.grad` (one unpack kernel per backward pass,
            # ops._WgradQueue); autograd gets None for it
            wp = ctx.weight_param
            if wp is None or not ops.conv_wgrad_param(x, dy, wp, kh, kw, stride, pt, pl):
                dw = ops.conv_wgrad(x, dy, c_true, kh, kw, stride, pt, pl)
        if has_bias and ctx.needs_input_grad[2]:
            db = ops.channel_sum(dy)
        return dx, dw, db, None, None, None, None, None, None


class _InputConv(torch.autograd.Function):
    """First convolution of a network, straight from the reference's fp32 NCHW image: the layout
    conversion writes the W-padded row-window tensor and every filter ROW becomes one 64-deep k-block
    (msp_conv.cu).  The image receives no gradient."""

    @staticmethod
    def forward(ctx, x, weight, bias, stride, padding, relu, want_stats, geom):
        ctx.set_materialize_grads(False)
        k, c_true, kh, kw = weight.shape
        n, _, h, w = x.shape
        win_px, cpp = geom
        ho, wo, pt, pl = ops.conv_out_size(h, w, kh, kw, stride, padding)
        wp = max(w + pl, stride * (wo - 1) + win_px)
        wp += wp & 1
        xw = ops.nchw_to_rowwin(x, cpp, pl, wp)
        wr = ops.pack_weights_rowwin(weight, win_px)
        stats = _stats_buffer(want_stats, k, x.device)
        y = ops.conv_fprop_rowwin(xw, w, wr, bias.detach() if bias is not None else None, k, kh, kw, stride,
                                  pt, pl, ho, wo, win_px, relu=relu, stats=stats, c_true=c_true)
        ctx.geom = (kh, kw, stride, pt, pl, c_true, relu, bias is not None, win_px, w)
        ctx.weight_param = weight if (weight.is_leaf and weight.requires_grad) else None
        ctx.save_for_backward(xw, y if relu else None)
        if stats is not None:
            ctx.mark_non_differentiable(stats)
            return y, stats
        return y, None

    @staticmethod
    def backward(ctx, dy, _dstats):
        if dy is None:
            return (None,) * 8
        kh, kw, stride, pt, pl, c_true, relu, has_bias, win_px, w_img = ctx.geom
        xw, y = ctx.saved_tensors
        if ctx.needs_input_grad[0]:
            raise RuntimeError("medsegpretrainimagenet_b200: the input image of the first convolution "
                               "cannot require a gradient on the B200 path")
        if dy.stride(3) != 1:
            dy = dy.contiguous()
        if relu:
            dy = ops.relu_bwd(y, dy)
        dw = db = None
        if ctx.needs_input_grad[1]:
            wp = ctx.weight_param
            if wp is None or not ops.conv_wgrad_rowwin_param(xw, w_img, dy, wp, kh, kw, stride, pt, pl, win_px):
                dw = ops.conv_wgrad_rowwin(xw, w_img, dy, c_true, kh, kw, stride, pt, pl, win_px)
        if has_bias and ctx.needs_input_grad[2]:
            db = ops.channel_sum(dy)
        return None, dw, db, None, None, None, None, None


def input_conv2d(x_nchw, weight, bias=None, stride=1, padding=0, relu=False, want_stats=False):
    """conv2d on the fp32 NCHW network input; falls back to to_nhwc + conv2d when the layer does not fit
    the row-window scheme."""
    geom = ops.rowwin_geometry(weight.shape[1], weight.shape[3], stride)
    if geom is None or x_nchw.requires_grad:
        return conv2d(to_nhwc(x_nchw), weight, bias, stride, padding, relu, want_stats)
    return _InputConv.apply(x_nchw, weight, bias, stride, padding, relu, want_stats, geom)


def conv2d(x, weight, bias=None, stride=1, padding=0, relu=False, want_stats=False, out=None, link=None):
    return _Conv.apply(x, weight, bias, stride, padding, relu, want_stats, out, link)


class _UpConv2x(torch.autograd.Function):
    """UpConvBlock's nn.Upsample(scale_factor=2) -> Conv2d(k=2, padding='same') [-> ReLU] (blocks.py:531-539) as ONE
    folded operation on the low-res input (csrc/msp_conv.cu, "Folded up-convolution"): the x4 tensor of the reference is
    never formed, 9/16 of its FLOPs run, forward / dgrad / wgrad are the same tap-GEMM and wgrad kernels."""

    @staticmethod
    def forward(ctx, x, weight, bias, relu, out):
        k, c_true, kh, kw = weight.shape
        assert (kh, kw) == (2, 2)
        wf9, wd9 = ops.packed_folded_weights(weight)
        y = ops.upconv2x_fprop(x, wf9, bias.detach() if bias is not None else None, k, relu=relu, out=out,
                               c_true=c_true)
        ctx.cfg = (c_true, relu, bias is not None)
        ctx.x_shape = tuple(x.shape)
        ctx.weight_ref = weight
        ctx.save_for_backward(x, wd9, y if relu else None)
        return y

    @staticmethod
    def backward(ctx, dy):
        c_true, relu, has_bias = ctx.cfg
        x, wd9, y = ctx.saved_tensors
        if dy.stride(3) != 1:
            dy = dy.contiguous()
        if relu:
            dy = ops.relu_bwd(y, dy)
        dx = dw = db = None
        if ctx.needs_input_grad[0]:
            dx = ops.upconv2x_dgrad(dy, wd9, ctx.x_shape, c_true=c_true)
        if ctx.needs_input_grad[1]:
            dw = ops.upconv2x_wgrad(x, dy, ctx.weight_ref)        # None when it went straight into weight.grad
        if has_bias and ctx.needs_input_grad[2]:
            db = ops.channel_sum(dy)
        return dx, dw, db, None, None


def upconv2x(x, weight, bias=None, relu=True, out=None):
    return _UpConv2x.apply(x, weight, bias, relu, out)


class _ConvTranspose(torch.autograd.Function):
    """y = conv_transpose2d(x, weight, bias) (+ReLU), weight (in_channels, out_channels, kh, kw) as nn.ConvTranspose2d
    keeps it.  The transposed convolution is the data gradient of the convolution `parent` with OIHW weight = this very
    tensor (O = in_channels): forward = parent's dgrad kernel (bias / ReLU in its epilogue), input gradient = parent's
    fprop, weight gradient = parent's wgrad with the roles of the two activations swapped — all three the same tcgen05
    tap-GEMM kernels the convolutions use."""

    @staticmethod
    def forward(ctx, x, weight, bias, stride, padding, relu):
        cin, cout, kh, kw = weight.shape
        if cout % 8:
            raise ValueError("conv_transpose2d: out_channels must be a multiple of 8 on the B200 path")
        wf, wd = ops.packed_weights(weight, need_dgrad=True)      # parent conv: K = cin, C = cout
        y = ops.conv_transpose_fprop(x, wd, bias.detach() if bias is not None else None, ops.ceil8(cout), kh, kw, stride,
                                     padding, relu=relu, k_true=cin)
        ctx.geom = (kh, kw, stride, padding, cin, cout, relu, bias is not None)
        ctx.x_shape = tuple(x.shape)
        ctx.weight_param = weight if (weight.is_leaf and weight.requires_grad) else None
        ctx.save_for_backward(x, wf, y if relu else None)
        return y

    @staticmethod
    def backward(ctx, dy):
        kh, kw, stride, pad, cin, cout, relu, has_bias = ctx.geom
        x, wf, y = ctx.saved_tensors
        if dy.stride(3) != 1:
            dy = dy.contiguous()
        if relu:
            dy = ops.relu_bwd(y, dy)
        n, hi, wi, k8 = ctx.x_shape
        dx = dw = db = None
        if ctx.needs_input_grad[0]:
            dx = ops.conv_fprop(dy, wf, None, k8, kh, kw, stride, pad, pad, hi, wi, c_true=cout)
        if ctx.needs_input_grad[1]:
            wp = ctx.weight_param
            if wp is None or not ops.conv_wgrad_param(dy, x, wp, kh, kw, stride, pad, pad):
                dw = ops.conv_wgrad(dy, x, cout, kh, kw, stride, pad, pad)[:cin]
        if has_bias and ctx.needs_input_grad[2]:
            db = ops.channel_sum(dy)[:cout]
        return dx, dw, db, None, None, None


def conv_transpose2d(x, weight, bias=None, stride=1, padding=0, relu=False):
    return _ConvTranspose.apply(x, weight, bias, stride, padding, relu)


# ------------------------------------------------------------------------------------------------
# BatchNorm + activation + residual + per-sample scale
# ------------------------------------------------------------------------------------------------
class _BnAct(torch.autograd.Function):
    """out = act( s[n] * BN(x) + residual ).  Train mode: batch statistics from `stats` (the conv
    epilogue's sums), running buffers updated in place like nn.BatchNorm2d (momentum, unbiased
    variance).  Eval mode: running statistics."""

    @staticmethod
    def forward(ctx, x, stats, gamma, beta, residual, sample_scale, running_mean, running_var, training,
                momentum, eps, act, r_stride, group, conv_bias, persistent_stats, link):
        ctx.link = link
        n, h, w, c = x.shape
        count = n * h * w
        if training:
            if group is not None and dist.is_initialized() and dist.get_world_size(group) > 1:
                # SyncBN: [2C] fp32 sums over NVLink peer memory / NCCL; every rank holds the same per-GPU batch (weak
                # scaling).  Deterministic mode: the per-CTA rows are first added in fixed order on this rank.
                count = count * dist.get_world_size(group)
                par = _peer(group)
                if par is not None and 2 * c <= par.max_floats:
                    # ONE launch: rows -> local sums -> sums over the ranks (NVLink peer memory) -> mean / invstd /
                    # running statistics (csrc/msp_p2p.cu), instead of reduce_rows -> all-reduce -> bn_finalize
                    mi = torch.empty((2, c), dtype=torch.float32, device=x.device)
                    par.stats_exchange(stats, stats.shape[0] if stats.dim() == 3 else 1, c, reset=persistent_stats,
                                       finalize=(count, eps, momentum, mi, running_mean, running_var))
                else:
                    if stats.dim() == 3:
                        stats = ops.reduce_rows(stats, reset=persistent_stats)
                        persistent_stats = False
                    _allreduce_sum(stats, group)
                    mi = ops.bn_finalize(stats, count, eps, momentum, running_mean, running_var, reset=persistent_stats)
            else:
                mi = ops.bn_finalize(stats, count, eps, momentum, running_mean, running_var, reset=persistent_stats)
        else:
            mi = ops.bn_eval_stats(running_mean, running_var, eps)
        y = ops.bn_act_fwd(x, mi, gamma.detach() if gamma is not None else None,
                           beta.detach() if beta is not None else None, act, residual=residual,
                           r_stride=r_stride, sample_scale=sample_scale)
        ctx.cfg = (training, act, r_stride, count, group, residual is not None,
                   tuple(residual.shape) if residual is not None else None)
        # BatchNorm -> ReLU without shortcut / sample scale: the backward recomputes the mask from x and never reads y
        mask_from_x = act == ops.ACT_RELU and residual is None and sample_scale is None
        ctx.y_stride = y.stride(2)
        ctx.params = (gamma, beta)      # the Parameters themselves: the SyncBN path adds their gradients into `.grad`
        ctx.save_for_backward(x, None if mask_from_x else y, mi, gamma, sample_scale, beta if mask_from_x else None)
        return y

    @staticmethod
    def backward(ctx, dy):
        training, act, r_stride, count, group, has_res, res_shape = ctx.cfg
        x, y, mi, gamma, sscale, beta = ctx.saved_tensors
        n, h, w, c = x.shape
        if dy.stride(3) != 1 or (y is not None and dy.stride(2) != y.stride(2)):
            full = torch.empty((n, h, w, ctx.y_stride), dtype=dy.dtype, device=dy.device)[..., :c]
            dy = ops.copy_channels(dy if dy.stride(3) == 1 else dy.contiguous(), full)
        gd = gamma.detach() if gamma is not None else None
        bd = beta.detach() if beta is not None else None
        synced = training and group is not None and dist.is_initialized() and dist.get_world_size(group) > 1
        par = _peer(group) if synced else None
        fused = par is not None and 2 * c <= par.max_floats
        # dgamma / dbeta are the LOCAL sums (the gradient reducer averages parameters' gradients)
        if fused:
            # per-block rows -> local sums + sums over the ranks in ONE launch (no reduce_rows, no copies)
            ws, rows = ops.bn_act_bwd_reduce_rows(x, y, dy, mi, act, sample_scale=sscale, gamma=gd, beta=bd)
            sums = torch.empty((2, c), dtype=torch.float32, device=x.device)
            gamma_p, beta_p = ctx.params
            want_g = gamma_p is not None and ctx.needs_input_grad[2]
            want_b = beta_p is not None and ctx.needs_input_grad[3]
            if _grad_sink_ok(gamma_p if want_g else None) and _grad_sink_ok(beta_p if want_b else None):
                # dgamma / dbeta ADDED straight into param.grad (the reducer's bucket views, accumulated micro-batches):
                # autograd gets None and launches no accumulate kernel per parameter (ops._WgradQueue's idea)
                par.stats_exchange(ws, rows, c, global_out=sums, local_add=True,
                                   local_halves=(beta_p.grad if want_b else None, gamma_p.grad if want_g else None))
                dbeta = dgamma = None
            else:
                local = torch.empty((2, c), dtype=torch.float32, device=x.device)
                par.stats_exchange(ws, rows, c, local_out=local, global_out=sums)
                dbeta, dgamma = local[0], local[1]
        else:
            sums = ops.bn_act_bwd_reduce(x, y, dy, mi, act, sample_scale=sscale, gamma=gd, beta=bd)
            # only the SyncBN all-reduce below overwrites `sums` in place, so a copy is needed in that case alone
            dbeta, dgamma = (sums[0].clone(), sums[1].clone()) if synced else (sums[0], sums[1])
        dbias = None
        if ctx.needs_input_grad[14]:
            # gradient of the producing conv's bias = sum over pixels of dx.  Train mode: BatchNorm
            # removes the mean, so it is exactly zero; eval mode: gamma * invstd * sum(g).  Either way
            # it comes from the fp32 sums instead of a second pass over the bf16 dx.
            if training:
                dbias = torch.zeros(c, dtype=torch.float32, device=x.device)
            else:
                dbias = dbeta * mi[1] if gamma is None else dbeta * mi[1] * gamma.detach()
        if synced and not fused:
            _allreduce_sum(sums, group)
        elif not training:
            # frozen statistics: dx = gamma * invstd * s * g (no mean / projection terms)
            sums = torch.zeros_like(sums)
        dres = None
        need_res = has_res and ctx.needs_input_grad[4]
        if need_res:
            rn, rh, rw, rc = res_shape
            dres = ops.new_act(rn, rh, rw, rc, x.device, zero=(r_stride > 1 or rc > c))
        dx = ops.bn_act_bwd_apply(x, y, dy, mi, gd, act, sums, count, dres=dres, r_stride=r_stride,
                                  sample_scale=sscale, beta=bd)
        if not ctx.needs_input_grad[0]:
            dx = None
        if dres is not None and ctx.link is not None:
            ctx.link.dres, dres = dres, None    # consumed by the unit's first convolution (ResidualLink)
        return (dx, None, dgamma if gamma is not None and ctx.needs_input_grad[2] else None,
                dbeta if dbeta is not None and ctx.needs_input_grad[3] else None, dres, None, None, None, None, None, None, None,
                None, None, dbias, None, None)


def bn_act(x, stats, bn: torch.nn.BatchNorm2d, act=ops.ACT_NONE, residual=None, r_stride=1,
           sample_scale=None, group=None, conv_bias=None, persistent_stats=False, counters=None, link=None):
    """`counters`: list collecting the num_batches_tracked buffers to bump (one fused add at the end of the forward,
    converter._finish_forward) instead of one tiny launch per layer; None = bump immediately."""
    training = bn.training or bn.running_mean is None
    momentum = 0.1 if bn.momentum is None else bn.momentum
    if training and bn.num_batches_tracked is not None:
        if counters is None:
            bn.num_batches_tracked.add_(1)
        else:
            counters.append(bn.num_batches_tracked)
    return _BnAct.apply(x, stats, bn.weight, bn.bias, residual, sample_scale,
                        bn.running_mean if bn.track_running_stats else None,
                        bn.running_var if bn.track_running_stats else None, training, momentum, bn.eps,
                        act, r_stride, group, conv_bias, persistent_stats, link)


# ------------------------------------------------------------------------------------------------
# pooling / resampling / concat / gate
# ------------------------------------------------------------------------------------------------
class _MaxPool(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, k, stride, pad):
        y, idx = ops.maxpool_fwd(x, k, stride, pad, want_idx=ctx.needs_input_grad[0])
        ctx.cfg = (tuple(x.shape), k, stride, pad)
        ctx.save_for_backward(idx)
        return y

    @staticmethod
    def backward(ctx, dy):
        x_shape, k, stride, pad = ctx.cfg
        (idx,) = ctx.saved_tensors
        if dy.stride(3) != 1:
            dy = dy.contiguous()
        return ops.maxpool_bwd(idx, dy, x_shape, k, stride, pad), None, None, None


def maxpool2d(x, k, stride, pad):
    return _MaxPool.apply(x, k, stride, pad)


class _Upsample2x(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        return ops.upsample2x_fwd(x)

    @staticmethod
    def backward(ctx, dy):
        if dy.stride(3) != 1:
            dy = dy.contiguous()
        return ops.upsample2x_bwd(dy)


def upsample2x(x):
    return _Upsample2x.apply(x)


class _UpsampleBilinear2x(torch.autograd.Function):
    """nn.Upsample(scale_factor=2, mode='bilinear', align_corners=False)."""

    @staticmethod
    def forward(ctx, x):
        return ops.upsample_bilinear2x_fwd(x)

    @staticmethod
    def backward(ctx, dy):
        if dy.stride(3) != 1:
            dy = dy.contiguous()
        return ops.upsample_bilinear2x_bwd(dy)


def upsample_bilinear2x(x):
    return _UpsampleBilinear2x.apply(x)


class _AvgPool(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        ctx.hw = (x.shape[1], x.shape[2])
        return ops.avgpool_fwd(x)

    @staticmethod
    def backward(ctx, dy):
        return ops.avgpool_bwd(dy, *ctx.hw)


def global_avgpool(x):
    return _AvgPool.apply(x)


def _is_leading_slice(a, buf, ca) -> bool:
    return buf is not None and a.data_ptr() == buf.data_ptr() and a.stride() == buf.stride() and a.shape[3] == ca


class _Concat(torch.autograd.Function):
    """torch.cat along channels (blocks.py:628,635).  `buf`: a preallocated (N, H, W, Ca + Cb) buffer whose leading
    channel slice already holds `a` (its producer wrote there): only `b` is copied.  Backward hands out channel-slice
    views of dy."""

    @staticmethod
    def forward(ctx, a, b, buf):
        n, h, w, ca = a.shape
        cb = b.shape[3]
        if _is_leading_slice(a, buf, ca):
            out = buf
        else:
            out = ops.new_act(n, h, w, ca + cb, a.device)
            ops.copy_channels(a, out[..., :ca])
        ops.copy_channels(b, out[..., ca:])
        ctx.ca = ca
        return out

    @staticmethod
    def backward(ctx, dy):
        return dy[..., : ctx.ca], dy[..., ctx.ca:], None


def concat(a, b, buf=None):
    return _Concat.apply(a, b, buf)


class _GateConcat(torch.autograd.Function):
    """cat(x_up, skip * up2(p)) (blocks.py:625-628): the product is written straight into its slice; with `buf` (the
    concat buffer whose leading slice the up-conv's epilogue already filled) nothing is copied at all."""

    @staticmethod
    def forward(ctx, x_up, skip, p, buf):
        n, h, w, ca = x_up.shape
        cb = skip.shape[3]
        if _is_leading_slice(x_up, buf, ca):
            out = buf
        else:
            out = ops.new_act(n, h, w, ca + cb, x_up.device)
            ops.copy_channels(x_up, out[..., :ca])
        ops.gate_mul_fwd(skip, p, out=out[..., ca:])
        ctx.ca = ca
        ctx.save_for_backward(skip, p)
        return out

    @staticmethod
    def backward(ctx, dy):
        skip, p = ctx.saved_tensors
        dskip, dp = ops.gate_mul_bwd(skip, p, dy[..., ctx.ca:])
        return dy[..., : ctx.ca], dskip, dp, None


def gate_concat(x_up, skip, p, buf=None):
    return _GateConcat.apply(x_up, skip, p, buf)


# ------------------------------------------------------------------------------------------------
# segmentation head: 1x1 conv (<= 8 classes) + sigmoid / softmax, fp32 NCHW out
# ------------------------------------------------------------------------------------------------
class _FinalConvAct(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, weight, bias, act):
        k = weight.shape[0]
        w2d = weight.detach().reshape(k, -1)
        c_true = w2d.shape[1]
        if c_true != x.shape[3]:  # input channels stored zero-padded to a multiple of 8
            w2d = torch.nn.functional.pad(w2d, (0, x.shape[3] - c_true))
        w2d = w2d.contiguous()
        prob, _ = ops.final_conv_act_fwd(x, w2d, bias.detach() if bias is not None else None, act)
        ctx.cfg = (act, c_true, tuple(weight.shape), bias is not None)
        ctx.save_for_backward(x, w2d, prob)
        return prob

    @staticmethod
    def backward(ctx, dprob):
        act, c_true, wshape, has_bias = ctx.cfg
        x, w2d, prob = ctx.saved_tensors
        dx, dw, db = ops.final_conv_act_bwd(x, w2d, act, prob, dprob.contiguous(),
                                            need_dx=ctx.needs_input_grad[0], has_bias=has_bias)
        dw = dw[:, :c_true].reshape(wshape)
        return dx, dw, db, None


def final_conv_act(x, weight, bias, act: int):
    return _FinalConvAct.apply(x, weight, bias, act)
