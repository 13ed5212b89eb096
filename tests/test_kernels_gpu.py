"""GPU parity of every C-ABI kernel against plain fp32 PyTorch on the CPU (same seeded inputs, already
rounded to bf16 where the kernel consumes bf16, so the only differences are accumulation order and the
bf16 rounding of the outputs).  Tolerances are written next to each check."""
import math

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

DEV = "cuda"


def _ops():
    from medsegpretrainimagenet_b200 import ops
    return ops


def _bf(x):
    return x.to(torch.bfloat16).float()


def _to_nhwc(x_nchw, cpad=None):
    """CPU fp32 NCHW -> CUDA bf16 NHWC (host-side permutation, independent of the kernel under test)."""
    n, c, h, w = x_nchw.shape
    cpad = cpad or (c + 7) // 8 * 8
    y = torch.zeros((n, h, w, cpad), dtype=torch.bfloat16)
    y[..., :c] = x_nchw.permute(0, 2, 3, 1).to(torch.bfloat16)
    return y.to(DEV)


def _from_nhwc(y, c=None):
    y = y.float().cpu()
    if c is not None:
        y = y[..., :c]
    return y.permute(0, 3, 1, 2).contiguous()


def _close(got, ref, rel, what):
    scale = ref.abs().max().item() + 1e-12
    err = (got - ref).abs().max().item()
    assert err <= rel * scale, f"{what}: max abs err {err:.4g} vs scale {scale:.4g} (rel {err / scale:.3g} > {rel})"


def _ref_conv(x, w, b, stride, padding):
    if padding == "same":
        kh, kw = w.shape[2:]
        x = F.pad(x, ((kw - 1) // 2, kw - 1 - (kw - 1) // 2, (kh - 1) // 2, kh - 1 - (kh - 1) // 2))
        padding = 0
    return F.conv2d(x, w, b, stride=stride, padding=padding)


CONV_CASES = [
    # N, H, W, C, K, k, stride, padding
    (2, 16, 16, 64, 64, 3, 1, 1),
    (2, 14, 14, 256, 256, 3, 1, 1),
    (1, 56, 56, 64, 64, 3, 1, 1),
    (2, 8, 8, 256, 64, 1, 1, 0),
    (4, 7, 7, 2048, 512, 1, 1, 0),
    (2, 16, 16, 128, 128, 3, 2, 1),
    (3, 28, 28, 128, 128, 3, 2, 1),
    (2, 64, 64, 3, 64, 7, 2, 3),
    (2, 64, 64, 1, 64, 7, 2, 3),
    (2, 16, 16, 64, 32, 2, 1, "same"),
    (2, 16, 16, 64, 128, 2, 2, 0),
    (1, 4, 256, 16, 16, 3, 1, 1),
    (2, 32, 32, 32, 16, 3, 1, 1),
    (16, 1, 1, 2048, 1000, 1, 1, 0),
    (2, 20, 24, 96, 40, 3, 1, 1),
    (1, 130, 130, 16, 16, 3, 1, 1),
    # >= 65536 output pixels with 16 / 32 channels: the warp-level MMA kernels of csrc/msp_narrow.cu (ragged tiles, all
    # four channel combinations, even 'same' filter of the up-conv)
    (2, 190, 200, 16, 16, 3, 1, 1),
    (1, 256, 264, 32, 32, 3, 1, 1),
    (2, 192, 176, 32, 16, 2, 1, "same"),
    (1, 260, 256, 16, 32, 3, 1, 1),
]


@pytest.mark.parametrize("case", CONV_CASES, ids=lambda c: "x".join(map(str, c)))
def test_conv_fprop_dgrad_wgrad(case):
    ops = _ops()
    n, h, w, c, k, ks, stride, padding = case
    g = torch.Generator().manual_seed(100 + CONV_CASES.index(case))
    x = _bf(torch.randn((n, c, h, w), generator=g))
    wt = _bf(torch.randn((k, c, ks, ks), generator=g) / math.sqrt(c * ks * ks))
    bias = torch.randn((k,), generator=g)
    x.requires_grad_(True)
    wt.requires_grad_(True)
    ref = _ref_conv(x, wt, bias, stride, padding)
    dy = _bf(torch.randn(ref.shape, generator=g))
    ref.backward(dy)

    ho, wo, pt, pl = ops.conv_out_size(h, w, ks, ks, stride, padding)
    assert (ho, wo) == tuple(ref.shape[2:])
    xd = _to_nhwc(x.detach())
    wf, wd = ops.pack_weights(wt.detach().to(DEV))
    stats = torch.zeros((2, k), dtype=torch.float32, device=DEV)
    y = ops.conv_fprop(xd, wf, bias.to(DEV), k, ks, ks, stride, pt, pl, ho, wo, stats=stats)
    torch.cuda.synchronize()
    got = _from_nhwc(y)
    # bf16 output rounding (2^-9 relative) + fp32 accumulation-order noise
    _close(got, ref.detach(), 6e-3, "fprop")
    # fused BatchNorm statistics are the sums of the bf16-rounded outputs
    yb = y.float().cpu().reshape(-1, k)
    _close(stats[0].cpu(), yb.sum(0), 1e-3, "ch_sum")
    _close(stats[1].cpu(), (yb * yb).sum(0), 1e-3, "ch_sqsum")

    # ReLU epilogue without bias
    y2 = ops.conv_fprop(xd, wf, None, k, ks, ks, stride, pt, pl, ho, wo, relu=True)
    _close(_from_nhwc(y2), torch.relu(ref.detach() - bias.view(1, -1, 1, 1)), 6e-3, "fprop+relu")

    dyd = _to_nhwc(dy)
    dx = ops.conv_dgrad(dyd, wd, tuple(xd.shape), ks, ks, stride, pt, pl)
    _close(_from_nhwc(dx, c), x.grad, 6e-3, "dgrad")
    # accumulate form: dx2 = base + dgrad
    base = _bf(torch.randn((n, c, h, w), generator=g))
    dx2 = _to_nhwc(base)
    ops.conv_dgrad(dyd, wd, tuple(xd.shape), ks, ks, stride, pt, pl, out=dx2, accumulate=True)
    _close(_from_nhwc(dx2, c), x.grad + base, 1e-2, "dgrad(accumulate)")

    dw = ops.conv_wgrad(xd, dyd, c, ks, ks, stride, pt, pl)
    # fp32 accumulation over up to N*Ho*Wo pixels, fixed-order fp32 sum of the split-K partials
    _close(dw.cpu(), wt.grad, 2e-3, "wgrad")


def test_narrow_wgrad_on_channel_slices_and_last_kernel_name():
    """csrc/msp_narrow.cu wgrad reading x and dy as channel slices of wider buffers (concat buffers: pixel stride >
    channels), repeated bit for bit (fixed-order reduction), and reported under its own kernel name."""
    from medsegpretrainimagenet_b200 import _lib
    ops = _ops()
    g = torch.Generator().manual_seed(77)
    n, h, w, c, k = 2, 200, 190, 16, 32
    x = _bf(torch.randn((n, c, h, w), generator=g))
    dy = _bf(torch.randn((n, k, h, w), generator=g))
    wt = torch.zeros((k, c, 3, 3), requires_grad=True)
    F.conv2d(x, wt, None, padding=1).backward(dy)
    xbuf = torch.zeros((n, h, w, 48), dtype=torch.bfloat16, device=DEV)
    dybuf = torch.zeros((n, h, w, 40), dtype=torch.bfloat16, device=DEV)
    xbuf[..., 16:32] = _to_nhwc(x)
    dybuf[..., 8:40] = _to_nhwc(dy)
    a = ops.conv_wgrad(xbuf[..., 16:32], dybuf[..., 8:40], c, 3, 3, 1, 1, 1)
    assert _lib.lib.msp_conv_last_kernel().decode() == "narrow_wgrad_kernel"
    b = ops.conv_wgrad(xbuf[..., 16:32], dybuf[..., 8:40], c, 3, 3, 1, 1, 1)
    assert torch.equal(a, b)
    _close(a.cpu(), wt.grad, 2e-3, "narrow wgrad on slices")


ROWWIN_CASES = [
    # N, H, W, C, K, k, stride, padding   (first convolutions: tiny channel counts)
    (2, 64, 64, 3, 64, 7, 2, 3),       # DeepResNet stem (classification/models.py:43-46)
    (3, 50, 46, 1, 64, 7, 2, 3),       # 1-channel stem (cfg1), ragged tiles
    (2, 40, 36, 3, 64, 3, 1, 1),       # UNet_encoder first block (unet_models.py:440)
    (1, 33, 130, 3, 32, 3, 1, 1),      # wide rows: several W tiles
    (2, 16, 16, 12, 16, 3, 1, 1),      # 9..16 channels -> 4-pixel windows
    (2, 18, 20, 5, 24, 2, 1, "same"),  # even kernel, right/bottom padding only
]


@pytest.mark.parametrize("case", ROWWIN_CASES, ids=lambda c: "x".join(map(str, c)))
def test_conv_rowwin_fprop_wgrad(case):
    """Row-window path of the first convolution: fp32 NCHW image in, same results as F.conv2d."""
    ops = _ops()
    n, h, w, c, k, ks, stride, padding = case
    g = torch.Generator().manual_seed(300 + ROWWIN_CASES.index(case))
    x = _bf(torch.randn((n, c, h, w), generator=g))
    wt = _bf(torch.randn((k, c, ks, ks), generator=g) / math.sqrt(c * ks * ks))
    bias = torch.randn((k,), generator=g)
    wt.requires_grad_(True)
    ref = _ref_conv(x, wt, bias, stride, padding)
    dy = _bf(torch.randn(ref.shape, generator=g))
    ref.backward(dy)
    ho, wo, pt, pl = ops.conv_out_size(h, w, ks, ks, stride, padding)
    win_px, cpp = ops.rowwin_geometry(c, ks, stride)
    wp = max(w + pl, stride * (wo - 1) + win_px)
    wp += wp & 1
    xw = ops.nchw_to_rowwin(x.to(DEV), cpp, pl, wp)
    # the padded layout itself (bit-exact: x is bf16-representable)
    chk = torch.zeros((n, h, wp, cpp))
    chk[:, :, pl:pl + w, :c] = x.permute(0, 2, 3, 1)
    assert torch.equal(xw.float().cpu(), chk)
    # a bf16 NCHW batch (BASELINE cfg2 feeds bf16 images) gives the identical row-window tensor
    assert torch.equal(ops.nchw_to_rowwin(x.to(DEV).to(torch.bfloat16), cpp, pl, wp), xw)
    wr = ops.pack_weights_rowwin(wt.detach().to(DEV), win_px)
    stats = torch.zeros((2, k), dtype=torch.float32, device=DEV)
    y = ops.conv_fprop_rowwin(xw, w, wr, bias.to(DEV), k, ks, ks, stride, pt, pl, ho, wo, win_px, stats=stats)
    _close(_from_nhwc(y), ref.detach(), 6e-3, "rowwin fprop")
    yb = y.float().cpu().reshape(-1, k)
    _close(stats[0].cpu(), yb.sum(0), 1e-3, "rowwin ch_sum")
    _close(stats[1].cpu(), (yb * yb).sum(0), 1e-3, "rowwin ch_sqsum")
    dw = ops.conv_wgrad_rowwin(xw, w, _to_nhwc(dy), c, ks, ks, stride, pt, pl, win_px)
    _close(dw.cpu(), wt.grad, 2e-3, "rowwin wgrad")


def test_conv_wgrad_deterministic():
    """Split-K partials are summed in a fixed order: two runs are bit-identical."""
    ops = _ops()
    g = torch.Generator().manual_seed(9)
    x = _to_nhwc(_bf(torch.randn((4, 64, 28, 28), generator=g)))
    dy = _to_nhwc(_bf(torch.randn((4, 128, 28, 28), generator=g)))
    a = ops.conv_wgrad(x, dy, 64, 3, 3, 1, 1, 1)
    b = ops.conv_wgrad(x, dy, 64, 3, 3, 1, 1, 1)
    assert torch.equal(a, b)


VARIANT_CASES = [
    # N, H, W, C, K, k, stride, padding
    (3, 28, 28, 128, 128, 3, 1, 1),    # halo applies (streaming B); pair applies to the tap path
    (2, 14, 14, 256, 256, 3, 1, 1),    # 256-wide tiles
    (2, 30, 30, 64, 64, 3, 1, 1),      # resident weights in halo mode
    (5, 14, 14, 512, 256, 1, 1, 0),    # flat 1x1, odd number of M tiles for the CTA pair
    (3, 28, 28, 128, 128, 3, 2, 1),    # stride 2: tap / pair only
]


@pytest.mark.parametrize("policy", [(0, 0), (1, 0), (2, 0), (0, 2), (0, 1)], ids=lambda p: f"pair{p[0]}_halo{p[1]}")
@pytest.mark.parametrize("case", VARIANT_CASES, ids=lambda c: "x".join(map(str, c)))
def test_conv_kernel_variants(case, policy):
    """Single-CTA tap-GEMM, cta_group::2 CTA-pair tap-GEMM and the halo-reuse kernel give the same results."""
    ops = _ops()
    n, h, w, c, k, ks, stride, padding = case
    g = torch.Generator().manual_seed(500 + VARIANT_CASES.index(case))
    x = _bf(torch.randn((n, c, h, w), generator=g))
    wt = _bf(torch.randn((k, c, ks, ks), generator=g) / math.sqrt(c * ks * ks))
    bias = torch.randn((k,), generator=g)
    x.requires_grad_(True)
    ref = _ref_conv(x, wt, bias, stride, padding)
    dy = _bf(torch.randn(ref.shape, generator=g))
    ref.backward(dy)
    ho, wo, pt, pl = ops.conv_out_size(h, w, ks, ks, stride, padding)
    xd, dyd = _to_nhwc(x.detach()), _to_nhwc(dy)
    wf, wd = ops.pack_weights(wt.to(DEV))
    ops.set_conv_policy(*policy)
    try:
        stats = torch.zeros((2, k), dtype=torch.float32, device=DEV)
        y = ops.conv_fprop(xd, wf, bias.to(DEV), k, ks, ks, stride, pt, pl, ho, wo, stats=stats)
        dx = ops.conv_dgrad(dyd, wd, tuple(xd.shape), ks, ks, stride, pt, pl)
        torch.cuda.synchronize()
    finally:
        ops.set_conv_policy(-1, -1)
    _close(_from_nhwc(y), ref.detach(), 6e-3, "fprop")
    yb = y.float().cpu().reshape(-1, k)
    _close(stats[0].cpu(), yb.sum(0), 1e-3, "ch_sum")
    _close(stats[1].cpu(), (yb * yb).sum(0), 1e-3, "ch_sqsum")
    _close(_from_nhwc(dx, c), x.grad, 6e-3, "dgrad")


def test_conv_concat_slices():
    """Operands that are channel slices of wider buffers (zero-copy concat, blocks.py:628,635)."""
    ops = _ops()
    g = torch.Generator().manual_seed(7)
    n, h, w, c, k = 2, 16, 16, 32, 64
    x = _bf(torch.randn((n, c, h, w), generator=g))
    wt = _bf(torch.randn((k, c, 3, 3), generator=g) / 17.0)
    ref = F.conv2d(x, wt, None, padding=1)
    xin = torch.zeros((n, h, w, 96), dtype=torch.bfloat16, device=DEV)
    xin[..., 40:72] = _to_nhwc(x)
    out = torch.full((n, h, w, 160), 3.0, dtype=torch.bfloat16, device=DEV)
    wf, _ = ops.pack_weights(wt.to(DEV), need_dgrad=False)
    ops.conv_fprop(xin[..., 40:72], wf, None, k, 3, 3, 1, 1, 1, h, w, out=out[..., 64:128])
    _close(_from_nhwc(out[..., 64:128]), ref, 6e-3, "fprop into slice")
    assert (out[..., :64] == 3).all() and (out[..., 128:] == 3).all(), "wrote outside the channel slice"


def test_layout_roundtrip():
    ops = _ops()
    g = torch.Generator().manual_seed(3)
    x = torch.randn((3, 5, 17, 23), generator=g)
    y = ops.nchw_to_nhwc(x.to(DEV))
    assert y.shape == (3, 17, 23, 8)
    assert torch.equal(y[..., :5].cpu(), x.permute(0, 2, 3, 1).to(torch.bfloat16))
    assert (y[..., 5:] == 0).all()
    back = ops.nhwc_to_nchw(y, 5)
    assert torch.equal(back.cpu(), _bf(x))


@pytest.mark.parametrize("shape", [(4, 8, 8, 64), (2, 7, 9, 24), (2, 4, 4, 2048)])
@pytest.mark.parametrize("act", [0, 1, 2])
def test_bn_act_fwd_bwd(shape, act):
    ops = _ops()
    n, h, w, c = shape
    g = torch.Generator().manual_seed(11 + act)
    x = _bf(torch.randn((n, c, h, w), generator=g) * 2 + 0.5)
    gamma = torch.rand((c,), generator=g) + 0.5
    beta = torch.randn((c,), generator=g) * 0.1
    res = _bf(torch.randn((n, c, h, w), generator=g))
    scale = (torch.rand((n,), generator=g) > 0.3).float()
    rm, rv = torch.zeros(c), torch.ones(c)
    xr = x.clone().requires_grad_(True)
    gr, br, rr = gamma.clone().requires_grad_(True), beta.clone().requires_grad_(True), res.clone().requires_grad_(True)
    bn = F.batch_norm(xr, rm, rv, gr, br, training=True, momentum=0.1, eps=1e-5)
    pre = bn * scale.view(-1, 1, 1, 1) + rr
    ref = torch.relu(pre) if act == 1 else (torch.sigmoid(pre) if act == 2 else pre)
    dy = _bf(torch.randn(ref.shape, generator=g))
    ref.backward(dy)

    xd = _to_nhwc(x)
    flat = xd.float().reshape(-1, c)
    stats = torch.stack([flat.sum(0), (flat * flat).sum(0)])
    rmd, rvd = torch.zeros(c, device=DEV), torch.ones(c, device=DEV)
    mi = ops.bn_finalize(stats, n * h * w, 1e-5, 0.1, rmd, rvd)
    _close(rmd.cpu(), rm, 1e-4, "running_mean")
    _close(rvd.cpu(), rv, 1e-4, "running_var")
    resd = _to_nhwc(res)
    y = ops.bn_act_fwd(xd, mi, gamma.to(DEV), beta.to(DEV), act, residual=resd, sample_scale=scale.to(DEV))
    _close(_from_nhwc(y), ref.detach(), 6e-3, "bn_act_fwd")

    dyd = _to_nhwc(dy)
    sums = ops.bn_act_bwd_reduce(xd, y, dyd, mi, act, sample_scale=scale.to(DEV))
    # dgamma / dbeta: the activation mask comes from the bf16 output, so allow bf16-level noise
    _close(sums[0].cpu(), br.grad, 2e-2, "dbeta")
    _close(sums[1].cpu(), gr.grad, 2e-2, "dgamma")
    dres = ops.new_act(n, h, w, c, DEV)
    dx = ops.bn_act_bwd_apply(xd, y, dyd, mi, gamma.to(DEV), act, sums, n * h * w, dres=dres,
                              sample_scale=scale.to(DEV))
    _close(_from_nhwc(dx), xr.grad, 2e-2, "bn dx")
    _close(_from_nhwc(dres), rr.grad, 2e-2, "residual grad")


def test_bn_relu_backward_mask_from_x_is_identical():
    """BatchNorm -> ReLU without shortcut: the backward may skip reading the stored activation and recompute the ReLU
    mask from x with the forward's own expression — results must be BIT-identical to the stored-activation path."""
    ops = _ops()
    g = torch.Generator().manual_seed(6)
    n, h, w, c = 3, 9, 7, 48
    x = _to_nhwc(_bf(torch.randn((n, c, h, w), generator=g) * 2))
    dy = _to_nhwc(_bf(torch.randn((n, c, h, w), generator=g)))
    gamma = (torch.rand((c,), generator=g) + 0.5).to(DEV)
    beta = (torch.randn((c,), generator=g) * 0.3).to(DEV)
    flat = x.float().reshape(-1, c)
    stats = torch.stack([flat.sum(0), (flat * flat).sum(0)])
    mi = ops.bn_finalize(stats, n * h * w, 1e-5, 0.1)
    y = ops.bn_act_fwd(x, mi, gamma, beta, ops.ACT_RELU)
    s_ref = ops.bn_act_bwd_reduce(x, y, dy, mi, ops.ACT_RELU, gamma=gamma, beta=beta)
    s_new = ops.bn_act_bwd_reduce(x, None, dy, mi, ops.ACT_RELU, gamma=gamma, beta=beta)
    dx_ref = ops.bn_act_bwd_apply(x, y, dy, mi, gamma, ops.ACT_RELU, s_ref, n * h * w, beta=beta)
    dx_new = ops.bn_act_bwd_apply(x, None, dy, mi, gamma, ops.ACT_RELU, s_ref, n * h * w, beta=beta)
    assert torch.equal(dx_ref, dx_new)
    # the per-channel sums are accumulated with atomics across blocks (order varies): compare within fp32 noise
    _close(s_new.cpu(), s_ref.cpu(), 1e-5, "sums")
    # omitting y is refused where the mask cannot be recomputed (shortcut / sigmoid)
    import medsegpretrainimagenet_b200._lib as L
    with pytest.raises(L.MspError):
        ops.bn_act_bwd_reduce(x, None, dy, mi, ops.ACT_SIGMOID, gamma=gamma, beta=beta)


def test_bn_zero_fill_strided_shortcut():
    """ResNet shortcut: stride-2 sub-sampling + zero channel fill (classification/models.py:257-274)."""
    ops = _ops()
    g = torch.Generator().manual_seed(5)
    n, h, w, c, rc = 2, 6, 6, 32, 16
    x = _bf(torch.randn((n, c, h, w), generator=g))
    res = _bf(torch.randn((n, rc, 2 * h, 2 * w), generator=g))
    short = torch.cat([res[:, :, ::2, ::2], torch.zeros((n, c - rc, h, w))], 1)
    ref = torch.relu(F.batch_norm(x, None, None, None, None, training=True) + short)
    xd = _to_nhwc(x)
    flat = xd.float().reshape(-1, c)
    mi = ops.bn_finalize(torch.stack([flat.sum(0), (flat * flat).sum(0)]), n * h * w, 1e-5, 0.1)
    y = ops.bn_act_fwd(xd, mi, None, None, 1, residual=_to_nhwc(res), r_stride=2)
    _close(_from_nhwc(y), ref, 6e-3, "strided zero-fill shortcut")
    dy = _bf(torch.randn(ref.shape, generator=g))
    dyd = _to_nhwc(dy)
    sums = ops.bn_act_bwd_reduce(xd, y, dyd, mi, 1)
    dres = torch.zeros((n, 2 * h, 2 * w, rc), dtype=torch.bfloat16, device=DEV)
    ops.bn_act_bwd_apply(xd, y, dyd, mi, None, 1, sums, n * h * w, dres=dres, r_stride=2)
    gmask = dy * (ref > 0)
    exp = torch.zeros((n, rc, 2 * h, 2 * w))
    exp[:, :, ::2, ::2] = gmask[:, :rc]
    _close(_from_nhwc(dres), exp, 1e-2, "strided shortcut grad")


@pytest.mark.parametrize("k,s,p,hw", [(3, 2, 1, (14, 14)), (2, 2, 0, (14, 14)), (3, 2, 1, (15, 13)), (2, 2, 0, (15, 13)),
                                      (3, 1, 1, (9, 11)), (5, 3, 2, (17, 12))])
def test_maxpool(k, s, p, hw):
    """(3,2,1) and (2,2,0) run the fixed-geometry kernels (ResNet stem / U-Net encoder), the rest the generic ones."""
    ops = _ops()
    g = torch.Generator().manual_seed(9)
    x = _bf(torch.randn((2, 16, *hw), generator=g)).requires_grad_(True)
    ref = F.max_pool2d(x, k, s, p)
    dy = _bf(torch.randn(ref.shape, generator=g))
    ref.backward(dy)
    xd = _to_nhwc(x.detach())
    y, idx = ops.maxpool_fwd(xd, k, s, p)
    assert torch.equal(_from_nhwc(y), ref.detach())
    dx = ops.maxpool_bwd(idx, _to_nhwc(dy), tuple(xd.shape), k, s, p)
    _close(_from_nhwc(dx), x.grad, 1e-2, "maxpool bwd")


def test_upsample_avgpool_gate_elementwise():
    ops = _ops()
    g = torch.Generator().manual_seed(13)
    x = _bf(torch.randn((2, 16, 6, 5), generator=g)).requires_grad_(True)
    up = F.interpolate(x, scale_factor=2, mode="nearest")
    dy = _bf(torch.randn(up.shape, generator=g))
    up.backward(dy)
    xd = _to_nhwc(x.detach())
    assert torch.equal(_from_nhwc(ops.upsample2x_fwd(xd)), up.detach())
    _close(_from_nhwc(ops.upsample2x_bwd(_to_nhwc(dy))), x.grad, 1e-2, "upsample bwd")

    ap = ops.avgpool_fwd(xd)
    _close(_from_nhwc(ap), x.detach().mean((2, 3), keepdim=True), 6e-3, "avgpool")
    dap = _bf(torch.randn((2, 16, 1, 1), generator=g))
    _close(_from_nhwc(ops.avgpool_bwd(_to_nhwc(dap), 6, 5)), (dap / 30).expand(2, 16, 6, 5), 1e-2, "avgpool bwd")

    skip = _bf(torch.randn((2, 16, 12, 10), generator=g)).requires_grad_(True)
    p = _bf(torch.rand((2, 16, 6, 5), generator=g)).requires_grad_(True)
    out = skip * F.interpolate(p, scale_factor=2, mode="nearest")
    out.backward(dy)
    sd, pd = _to_nhwc(skip.detach()), _to_nhwc(p.detach())
    _close(_from_nhwc(ops.gate_mul_fwd(sd, pd)), out.detach(), 6e-3, "gate fwd")
    dskip, dp = ops.gate_mul_bwd(sd, pd, _to_nhwc(dy))
    _close(_from_nhwc(dskip), skip.grad, 1e-2, "gate dskip")
    _close(_from_nhwc(dp), p.grad, 1e-2, "gate dp")

    a, b = _bf(torch.randn((2, 8, 3, 3), generator=g)), _bf(torch.randn((2, 8, 3, 3), generator=g))
    ad, bd = _to_nhwc(a), _to_nhwc(b)
    _close(_from_nhwc(ops.add_relu(ad, bd)), torch.relu(a + b), 6e-3, "add_relu")
    _close(_from_nhwc(ops.add(ad, bd)), a + b, 6e-3, "add")
    _close(_from_nhwc(ops.relu_bwd(ad, bd)), b * (a > 0), 1e-6, "relu_bwd")
    _close(ops.channel_sum(ad).cpu(), a.sum((0, 2, 3)), 1e-3, "channel_sum")
    buf = torch.zeros((2, 3, 3, 24), dtype=torch.bfloat16, device=DEV)
    ops.copy_channels(ad, buf[..., 8:16])
    assert torch.equal(buf[..., 8:16].float().cpu(), a.permute(0, 2, 3, 1))


@pytest.mark.parametrize("k,act", [(1, "sigmoid"), (4, "softmax"), (5, "sigmoid"), (2, None)])
@pytest.mark.parametrize("c", [16, 64])
def test_final_conv_act(k, act, c):
    ops = _ops()
    g = torch.Generator().manual_seed(17 + k)
    n, h, w = 2, 24, 20
    x = _bf(torch.randn((n, c, h, w), generator=g)).requires_grad_(True)
    wt = (torch.randn((k, c), generator=g) / math.sqrt(c)).requires_grad_(True)
    b = torch.randn((k,), generator=g).requires_grad_(True)
    logits = F.conv2d(x, wt.view(k, c, 1, 1), b)
    ref = torch.sigmoid(logits) if act == "sigmoid" else (torch.softmax(logits, 1) if act == "softmax" else logits)
    dp = torch.randn(ref.shape, generator=g)
    ref.backward(dp)
    xd = _to_nhwc(x.detach())
    a = ops.HEAD_ACT[act]
    prob, _ = ops.final_conv_act_fwd(xd, wt.detach().to(DEV), b.detach().to(DEV), a)
    # fp32 math on identical bf16 inputs
    _close(prob.cpu(), ref.detach(), 1e-5, "final conv prob")
    dx, dw, db = ops.final_conv_act_bwd(xd, wt.detach().to(DEV), a, prob, dp.to(DEV))
    _close(_from_nhwc(dx), x.grad, 6e-3, "final conv dx")
    _close(dw.cpu(), wt.grad, 1e-4, "final conv dw")
    _close(db.cpu(), b.grad, 1e-4, "final conv db")


# ------------------------------------------------------------------------------------------------
# deterministic mode (torch.use_deterministic_algorithms; every downstream YAML of the reference sets it)
# ------------------------------------------------------------------------------------------------
@pytest.fixture()
def deterministic_mode():
    torch.use_deterministic_algorithms(True, warn_only=True)
    try:
        yield
    finally:
        torch.use_deterministic_algorithms(False)


@pytest.mark.parametrize("cin, cout, ks, hw, n", [(64, 64, 3, 28, 16), (256, 512, 1, 14, 32), (16, 16, 3, 64, 4),
                                                  (32, 32, 3, 48, 3), (16, 16, 3, 136, 4), (32, 16, 3, 150, 3)])
def test_deterministic_statistics_rows_match_the_atomics_and_repeat_exactly(deterministic_mode, cin, cout, ks, hw, n):
    """Conv epilogue statistics as per-CTA rows [SMs, 2, K] added in fixed order == the float-atomics accumulators
    (to fp32 rounding), and two runs give the same BITS; the narrow layers go through the multi-tile epilogue."""
    from medsegpretrainimagenet_b200 import ops
    g = torch.Generator().manual_seed(0)
    x = torch.randn((n, hw, hw, cin), generator=g).to(torch.bfloat16).to(DEV)
    w = (torch.randn((cout, cin, ks, ks), generator=g) * 0.1).to(DEV)
    wf, _ = ops.pack_weights(w, need_dgrad=False)
    pad = ks // 2
    totals, outs = [], []
    for _ in range(2):
        stats = ops.new_stats(cout, DEV)
        assert stats.dim() == 3 and stats.shape[0] == ops.sm_count(DEV)
        y = ops.conv_fprop(x, wf, None, cout, ks, ks, 1, pad, pad, hw, hw, stats=stats)
        totals.append(ops.reduce_rows(stats).clone())
        outs.append(y.clone())
        mi = ops.bn_finalize(stats, n * hw * hw, 1e-5, 0.1, reset=True)
        assert float(stats.abs().max()) == 0.0                       # self-cleaning rows
    assert torch.equal(totals[0], totals[1]) and torch.equal(outs[0], outs[1])
    yf = outs[0].float().reshape(-1, cout)
    ref = torch.stack([yf.sum(0), (yf * yf).sum(0)])
    assert ((totals[0] - ref).abs().max() / ref.abs().max()).item() <= 1e-3
    m_ref = yf.mean(0)
    assert ((mi[0] - m_ref).abs().max()).item() <= 1e-3 * max(1.0, m_ref.abs().max().item())
    torch.use_deterministic_algorithms(False)
    st = torch.zeros((2, cout), device=DEV)
    ops.conv_fprop(x, wf, None, cout, ks, ks, 1, pad, pad, hw, hw, stats=st)
    assert ((st - totals[0]).abs().max() / totals[0].abs().max()).item() <= 1e-5


def test_deterministic_bn_backward_sums_bias_and_head_gradients(deterministic_mode):
    from medsegpretrainimagenet_b200 import ops
    g = torch.Generator().manual_seed(1)
    n, hw, c = 6, 40, 64
    x = torch.randn((n, hw, hw, c), generator=g).to(torch.bfloat16).to(DEV)
    dy = torch.randn((n, hw, hw, c), generator=g).to(torch.bfloat16).to(DEV)
    mi = torch.stack([torch.zeros(c), torch.ones(c)]).to(DEV)
    gamma, beta = torch.ones(c, device=DEV), torch.zeros(c, device=DEV)
    a = ops.bn_act_bwd_reduce(x, None, dy, mi, ops.ACT_RELU, gamma=gamma, beta=beta).clone()
    b = ops.bn_act_bwd_reduce(x, None, dy, mi, ops.ACT_RELU, gamma=gamma, beta=beta).clone()
    assert torch.equal(a, b)
    gm = (dy.float() * (x.float() > 0)).reshape(-1, c)
    ref = torch.stack([gm.sum(0), (gm * x.float().reshape(-1, c)).sum(0)])
    assert ((a - ref).abs().max() / ref.abs().max()).item() <= 1e-3
    s1, s2 = ops.channel_sum(dy).clone(), ops.channel_sum(dy).clone()
    assert torch.equal(s1, s2)
    assert ((s1 - dy.float().reshape(-1, c).sum(0)).abs().max()).item() <= 1e-2
    k = 4
    w2d, prob = torch.randn((k, c), generator=g).to(DEV), torch.softmax(torch.randn((n, k, hw, hw), generator=g), 1).to(DEV)
    dprob = torch.randn((n, k, hw, hw), generator=g).to(DEV)
    r1 = ops.final_conv_act_bwd(x, w2d, 2, prob, dprob)
    r2 = ops.final_conv_act_bwd(x, w2d, 2, prob, dprob)
    assert torch.equal(r1[1], r2[1]) and torch.equal(r1[2], r2[2]) and torch.equal(r1[0], r2[0])
    torch.use_deterministic_algorithms(False)
    r0 = ops.final_conv_act_bwd(x, w2d, 2, prob, dprob)
    assert ((r0[1] - r1[1]).abs().max() / r0[1].abs().max()).item() <= 1e-4
    assert ((r0[2] - r1[2]).abs().max() / r0[2].abs().max()).item() <= 1e-4


def test_batched_weight_gradient_unpack_and_tiled_pack_match_the_per_layer_kernels():
    """ops._WgradQueue (one unpack launch per backward, straight into param.grad, accumulate mode) against the per-layer
    path, and the tiled transposing weight pack against the element-wise one."""
    import os
    from medsegpretrainimagenet_b200 import functional as Fn, ops
    g = torch.Generator().manual_seed(2)
    x = torch.randn((3, 20, 20, 24), generator=g).to(torch.bfloat16).to(DEV).requires_grad_(True)
    ws = [torch.nn.Parameter((torch.randn(s, generator=g) * 0.1).to(DEV)) for s in ((40, 24, 3, 3), (16, 40, 1, 1),
                                                                                   (32, 16, 2, 2))]

    def run():
        for w in ws:
            w.grad = None
        y, _ = Fn.conv2d(x, ws[0], None, 1, 1)
        y, _ = Fn.conv2d(y, ws[1], None, 1, 0)
        y, _ = Fn.conv2d(y, ws[2], None, 1, "same")
        y.float().square().mean().backward()
        return [w.grad.clone() for w in ws]
    sink = run()
    twice = None
    for w in ws:                       # accumulate mode: a second backward adds into the existing gradients
        pass
    y, _ = Fn.conv2d(x, ws[0], None, 1, 1)
    y, _ = Fn.conv2d(y, ws[1], None, 1, 0)
    y, _ = Fn.conv2d(y, ws[2], None, 1, "same")
    y.float().square().mean().backward()
    twice = [w.grad.clone() for w in ws]
    ops._WGRAD_SINK = False
    try:
        plain = run()
    finally:
        ops._WGRAD_SINK = True
    for a, b, c2 in zip(sink, plain, twice):
        assert torch.equal(a, b)
        assert torch.allclose(c2, 2 * a, rtol=1e-6, atol=1e-8)
    # weight packing: batched tiled kernel == per-layer kernel
    cache = ops.WeightPackCache()
    per_layer = [ops.pack_weights(w) for w in ws]
    for w in ws:
        cache.lookup(w, True)
    cache.build_table()
    cache.begin_step()
    for w, (wf, wd) in zip(ws, per_layer):
        bf, bd = cache.lookup(w, True)
        assert torch.equal(bf, wf) and torch.equal(bd, wd)


@pytest.mark.parametrize("cin, cout, k, s, p, hw", [(32, 16, 2, 2, 0, 12), (64, 32, 4, 2, 1, 9), (16, 24, 3, 1, 1, 10),
                                                    (128, 64, 2, 2, 0, 16)])
def test_conv_transpose2d_matches_torch(cin, cout, k, s, p, hw):
    """nn.ConvTranspose2d (north star: transposed-conv layers) on the dgrad / fprop / wgrad tap-GEMM kernels against
    F.conv_transpose2d in fp32 on bf16-representable operands: output within 6e-3 of its range (bf16 storage), input /
    weight / bias gradients within 1e-2 / 2e-3 / 2e-3 of theirs."""
    from medsegpretrainimagenet_b200 import functional as Fn
    g = torch.Generator().manual_seed(0)
    n = 3
    x = torch.randn((n, cin, hw, hw), generator=g).to(torch.bfloat16).float()
    w = (torch.randn((cin, cout, k, k), generator=g) * 0.1).to(torch.bfloat16).float()
    b = torch.randn((cout,), generator=g)
    xr, wr, br = x.clone().requires_grad_(True), w.clone().requires_grad_(True), b.clone().requires_grad_(True)
    yr = F.relu(F.conv_transpose2d(xr, wr, br, stride=s, padding=p))
    gy = torch.randn(yr.shape, generator=g).to(torch.bfloat16).float()
    yr.backward(gy)
    xg = x.to(DEV).requires_grad_(True)
    wg = torch.nn.Parameter(w.to(DEV))
    bg = torch.nn.Parameter(b.to(DEV))
    yn = Fn.conv_transpose2d(Fn.to_nhwc(xg), wg, bg, s, p, relu=True)
    y = Fn.to_nchw(yn, cout)
    assert y.shape == yr.shape
    y.backward(gy.to(DEV))
    _close(y.detach().cpu(), yr.detach(), 6e-3, "conv_transpose y")
    _close(xg.grad.cpu(), xr.grad, 1e-2, "conv_transpose dx")
    _close(wg.grad.cpu(), wr.grad, 2e-3, "conv_transpose dw")
    _close(bg.grad.cpu(), br.grad, 2e-3, "conv_transpose db")


def test_conv_transpose2d_module_is_converted():
    """The converter maps nn.ConvTranspose2d (+ReLU) inside a reference-structured block."""
    from medsegpretrainimagenet_b200 import converter as cv, functional as Fn
    torch.manual_seed(0)
    seq = torch.nn.Sequential(torch.nn.ConvTranspose2d(32, 16, 2, stride=2), torch.nn.ReLU()).to(DEV)
    with torch.no_grad():
        for prm in seq.parameters():
            prm.copy_(prm.to(torch.bfloat16).float())
    x = torch.randn((2, 32, 8, 8), generator=torch.Generator().manual_seed(1)).to(torch.bfloat16).float().to(DEV)
    y = Fn.to_nchw(cv.run_sequence(cv.ExecContext(), list(seq.children()), Fn.to_nhwc(x)))
    _close(y.detach().cpu(), seq(x).detach().cpu(), 6e-3, "ConvTranspose2d module")
    with pytest.raises(cv.UnsupportedModule):
        cv.run_sequence(cv.ExecContext(), [torch.nn.ConvTranspose2d(32, 16, 2, stride=2, output_padding=1).to(DEV)],
                        Fn.to_nhwc(x))


@pytest.mark.parametrize("n, c, h, w", [(2, 16, 5, 7), (1, 64, 1, 9), (3, 8, 12, 12)])
def test_bilinear_upsample_matches_torch(n, c, h, w):
    """nn.Upsample(scale_factor=2, mode='bilinear') forward and backward against F.interpolate (fp32) on
    bf16-representable inputs: outputs within one bf16 rounding (2^-8 relative), gradients likewise."""
    from medsegpretrainimagenet_b200 import converter as cv, functional as Fn
    g = torch.Generator().manual_seed(0)
    x = torch.randn((n, c, h, w), generator=g).to(torch.bfloat16).float()
    xr = x.clone().requires_grad_(True)
    yr = F.interpolate(xr, scale_factor=2, mode="bilinear", align_corners=False)
    gy = torch.randn(yr.shape, generator=g).to(torch.bfloat16).float()
    yr.backward(gy)
    xg = x.to(DEV).requires_grad_(True)
    up = torch.nn.Upsample(scale_factor=2, mode="bilinear")
    y = Fn.to_nchw(cv.run_sequence(cv.ExecContext(), [up], Fn.to_nhwc(xg)))
    y.backward(gy.to(DEV))
    _close(y.detach().cpu(), yr.detach(), 5e-3, "bilinear y")
    _close(xg.grad.cpu(), xr.grad, 5e-3, "bilinear dx")
    with pytest.raises(cv.UnsupportedModule):
        cv.run_sequence(cv.ExecContext(), [torch.nn.Upsample(scale_factor=2, mode="bilinear", align_corners=True)],
                        Fn.to_nhwc(xg))


def test_softmax_ce_soft_labels_and_spatial_logits_match_torch():
    """torch.nn.CrossEntropyLoss with class-probability targets (Mixup / CutMix, advanced.yaml:17,48) and
    F.cross_entropy on spatial logits, both with label smoothing: loss rel <= 1e-5, gradients <= 1e-5 of their range."""
    from medsegpretrainimagenet_b200 import losses
    g = torch.Generator().manual_seed(0)
    for smooth in (0.0, 0.1):
        z = torch.randn((37, 1000), generator=g) * 3
        t = torch.zeros((37, 1000))
        a, b = torch.randint(0, 1000, (37,), generator=g), torch.randint(0, 1000, (37,), generator=g)
        lam = torch.rand((37,), generator=g)
        t[torch.arange(37), a] += lam
        t[torch.arange(37), b] += 1 - lam                         # mixed one-hot pairs
        zr = z.clone().requires_grad_(True)
        lr_ = F.cross_entropy(zr, t, label_smoothing=smooth)
        lr_.backward()
        zg = z.to(DEV).requires_grad_(True)
        lg = losses.TorchCrossEntropyLoss(label_smoothing=smooth)(zg, t.to(DEV))
        (lg * 1.0).backward()
        assert abs(lg.item() - lr_.item()) <= 1e-5 * abs(lr_.item())
        _close(zg.grad.cpu(), zr.grad, 1e-5, "soft CE grad")
        # spatial logits: (N, C, H, W) against (N, H, W) and (N, 1, H, W) class indices
        zs = torch.randn((3, 4, 19, 23), generator=g) * 2
        ys = torch.randint(0, 4, (3, 19, 23), generator=g)
        zsr = zs.clone().requires_grad_(True)
        ls_ = F.cross_entropy(zsr, ys, label_smoothing=smooth)
        ls_.backward()
        for lab in (ys, ys.unsqueeze(1)):
            zsg = zs.to(DEV).requires_grad_(True)
            lsg = losses.CrossEntropyLoss(smooth, apply_softmax=True)(zsg, lab.to(DEV))
            lsg.backward()
            assert abs(lsg.item() - ls_.item()) <= 1e-5 * abs(ls_.item())
            _close(zsg.grad.cpu(), zsr.grad, 1e-5, "spatial CE grad")
        # the hard-label (N, C) / (N, 1) form keeps working through the same entry
        y1 = torch.randint(0, 1000, (37, 1), generator=g)
        want = F.cross_entropy(z, y1.squeeze(1), label_smoothing=smooth)
        got = losses.CrossEntropyLoss(smooth, apply_softmax=True)(z.to(DEV), y1.to(DEV))
        assert abs(got.item() - want.item()) <= 1e-5 * abs(want.item())


@pytest.mark.parametrize("n, cin, cout, h, w", [(2, 64, 32, 7, 9), (1, 2048, 1024, 4, 4), (3, 32, 16, 33, 20),
                                                (2, 128, 64, 16, 16)])
def test_folded_upconv_matches_upsample_then_conv(n, cin, cout, h, w):
    """UpConvBlock (blocks.py:531-539): nearest x2 -> Conv2d(k=2, padding='same') [-> ReLU], folded onto the low-res input
    (four output-parity classes with pre-summed weights), against torch in fp32 on bf16-representable operands:
    output within 6e-3 of its range, dx 1e-2, dW 3e-3, db 2e-3.  The gradients are compared on the LINEAR layer (a ReLU
    mask that flips on a near-zero output moves single terms of these short sums by several per cent of the range); the
    ReLU epilogue and its backward mask are checked separately against a reference that uses the kernel's own mask."""
    from medsegpretrainimagenet_b200 import functional as Fn
    g = torch.Generator().manual_seed(0)
    x = torch.randn((n, cin, h, w), generator=g).to(torch.bfloat16).float()
    wt = (torch.randn((cout, cin, 2, 2), generator=g) * (1.0 / (4 * cin)) ** 0.5).to(torch.bfloat16).float()
    b = torch.randn((cout,), generator=g) * 0.1

    def reference(relu, mask=None):
        xr, wr, br = x.clone().requires_grad_(True), wt.clone().requires_grad_(True), b.clone().requires_grad_(True)
        yr = F.conv2d(F.pad(F.interpolate(xr, scale_factor=2, mode="nearest"), (0, 1, 0, 1)), wr, br)
        if relu:
            yr = yr * mask if mask is not None else F.relu(yr)
        return xr, wr, br, yr
    xr, wr, br, yr = reference(False)
    gy = torch.randn(yr.shape, generator=g).to(torch.bfloat16).float()
    yr.backward(gy)
    xg = x.to(DEV).requires_grad_(True)
    wg, bg = torch.nn.Parameter(wt.to(DEV)), torch.nn.Parameter(b.to(DEV))
    y = Fn.to_nchw(Fn.upconv2x(Fn.to_nhwc(xg), wg, bg, relu=False), cout)
    assert y.shape == yr.shape
    y.backward(gy.to(DEV))
    _close(y.detach().cpu(), yr.detach(), 6e-3, "folded up-conv y")
    _close(xg.grad.cpu(), xr.grad, 1e-2, "folded up-conv dx")
    _close(wg.grad.cpu(), wr.grad, 3e-3, "folded up-conv dW")
    _close(bg.grad.cpu(), br.grad, 2e-3, "folded up-conv db")
    # a second backward ADDS into the existing weight gradient (gradient accumulation / all-reduce bucket views)
    y2 = Fn.to_nchw(Fn.upconv2x(Fn.to_nhwc(xg), wg, bg, relu=False), cout)
    y2.backward(gy.to(DEV))
    _close(wg.grad.cpu(), 2 * wr.grad, 3e-3, "folded up-conv dW accumulated")
    # a non-leaf weight takes the autograd route (the gradient is returned instead of written into .grad)
    wv = wt.to(DEV).requires_grad_(True)
    y3 = Fn.to_nchw(Fn.upconv2x(Fn.to_nhwc(xg), wv * 1.0, None, relu=False), cout)
    y3.backward(gy.to(DEV))
    _close(wv.grad.cpu(), wr.grad, 3e-3, "folded up-conv dW (autograd route)")
    # ReLU epilogue + masked backward, the reference using the kernel's own mask
    xg2 = x.to(DEV).requires_grad_(True)
    wg2, bg2 = torch.nn.Parameter(wt.to(DEV)), torch.nn.Parameter(b.to(DEV))
    yk = Fn.to_nchw(Fn.upconv2x(Fn.to_nhwc(xg2), wg2, bg2, relu=True), cout)
    yk.backward(gy.to(DEV))
    mask = (yk.detach().cpu() > 0).float()
    xr2, wr2, br2, yr2 = reference(True, mask)
    yr2.backward(gy)
    _close(yk.detach().cpu(), F.relu(reference(False)[3]).detach(), 6e-3, "folded up-conv relu(y)")
    _close(xg2.grad.cpu(), xr2.grad, 1e-2, "folded up-conv dx through ReLU")
    _close(wg2.grad.cpu(), wr2.grad, 3e-3, "folded up-conv dW through ReLU")
    _close(bg2.grad.cpu(), br2.grad, 2e-3, "folded up-conv db through ReLU")


def test_unet_with_folded_upconvs_and_zero_copy_concat_equals_the_materialising_path():
    """The attention U-Net with the up-convs folded and x_up written straight into the concat buffers against the same
    network with the x4 tensors materialised (MSP_UPCONV_FOLD=0 path): predictions within bf16 noise (the folded weights
    are summed before their bf16 rounding), gradients aligned."""
    import medsegpretrainimagenet_b200 as b200
    from medsegpretrainimagenet_b200 import converter as cv, models
    torch.manual_seed(0)
    m = models.kaiming_init_(models.resnet18_attention_unet()).to(DEV).train()
    x = torch.rand((4, 1, 64, 64), device=DEV)
    y = (torch.rand((4, 1, 64, 64), device=DEV) < 0.3).long()
    torch.use_deterministic_algorithms(True, warn_only=True)

    def run():
        for p in m.parameters():
            p.grad = None
        for bn in [mod for mod in m.modules() if isinstance(mod, torch.nn.BatchNorm2d)]:
            bn.reset_running_stats()
        pred = m(x)
        loss = b200.losses.DiceLoss()(pred, y)
        loss.backward()
        return pred.detach().clone(), loss.item(), torch.cat([p.grad.flatten() for p in m.parameters()]).double()
    keep = cv._UPCONV_FOLD
    try:
        cv._UPCONV_FOLD = 2            # always fold (the default folds only layers large enough to pay for the launches)
        p1, l1, g1 = run()
        cv._UPCONV_FOLD = 0
        p0, l0, g0 = run()
    finally:
        cv._UPCONV_FOLD = keep
        torch.use_deterministic_algorithms(False)
    rms = ((p1 - p0).double().pow(2).mean().sqrt() / p0.double().pow(2).mean().sqrt()).item()
    cos = (g1 @ g0 / (g1.norm() * g0.norm())).item()
    # the two paths round different intermediate sums to bf16 (folded weights are added before their rounding); through
    # this randomly initialised 40-layer network that noise is amplified like any other bf16 perturbation
    # (tests/test_hotpath_gpu.py measures the same effect against the emulated-storage oracle)
    assert rms <= 5e-2 and abs(l1 - l0) <= 5e-3 * abs(l0) and cos >= 0.9, (rms, l1, l0, cos)
