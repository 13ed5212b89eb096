"""Input side of the step on the device (csrc/msp_input.cu): the uint8 -> /255 -> float32 -> RepeatChannels kernel against
numpy's arithmetic bit for bit, ColorJitter against torchvision on the CPU with the same parameter draw, the sequential
pretraining model against the classifier it restates, and `eval_encoder` (robustness/eval.py:56-69) end to end."""
import copy

import numpy as np
import pytest
import torch

from oracle import ref_models, ref_robustness

pytestmark = pytest.mark.gpu
DEV = "cuda"


@pytest.mark.parametrize("shape, repeats", [((3, 1, 37, 53), 3), ((2, 3, 64, 64), 1), ((5, 1, 7, 9), 2)])
def test_u8_input_kernel_is_numpy_exact(shape, repeats):
    from medsegpretrainimagenet_b200 import transforms as T
    rng = np.random.default_rng(0)
    x = rng.integers(0, 256, shape, dtype=np.uint8)
    # classification/datasets.py:47 (`np.load(f) / 255`, float64) -> RepeatChannels (np.repeat axis 0 of CHW) -> float32
    want = np.stack([np.repeat(img / 255, repeats, axis=0) for img in x]).astype(np.float32)
    got = T.DeviceInput(repeats=repeats)(torch.from_numpy(x).to(DEV)).cpu().numpy()
    assert got.shape == want.shape and np.array_equal(got, want)


@pytest.mark.parametrize("channels", [3, 1])
def test_color_jitter_matches_torchvision_for_every_op_order(channels):
    import itertools
    import torchvision.transforms.functional as F
    from medsegpretrainimagenet_b200 import transforms as T
    g = torch.Generator().manual_seed(1)
    imgs = torch.rand((3, channels, 45, 61), generator=g)
    imgs[0, :, :5, :5] = 0.5                      # gray patch: maxc == minc branch of the hue conversion
    b, c, s, h = 1.07, 0.96, 1.09, -0.043
    worst = 0.0
    for order in itertools.permutations(range(4)):
        ref = imgs.clone()
        for fn_id in order:
            ref = (F.adjust_brightness(ref, b) if fn_id == 0 else F.adjust_contrast(ref, c) if fn_id == 1
                   else F.adjust_saturation(ref, s) if fn_id == 2 else F.adjust_hue(ref, h))
        got = T.color_jitter_apply(imgs.to(DEV), order, b, c, s, h).cpu()
        worst = max(worst, (got - ref).abs().max().item())
    # everything but the contrast mean (a float32 tree sum in torch, a float64 sum here) is computed with identical
    # separately-rounded operations: the deviation is that mean's last bits, scaled by |1 - contrast|
    assert worst <= 2e-6, worst
    # without contrast the result is bit-identical
    ref = F.adjust_hue(F.adjust_saturation(F.adjust_brightness(imgs, b), s), h)
    got = T.color_jitter_apply(imgs.to(DEV), (0, 2, 3, 1), b, None, s, h).cpu()
    assert torch.equal(got, ref)


def test_color_jitter_class_consumes_the_cpu_generator_like_torchvision():
    import torchvision
    from medsegpretrainimagenet_b200 import transforms as T
    imgs = torch.rand((2, 3, 32, 32), generator=torch.Generator().manual_seed(2))
    kw = dict(brightness=0.1, contrast=0.05, hue=0.05, saturation=0.1)        # robustness/eval.py:61-64
    torch.manual_seed(11)
    ref0 = torchvision.transforms.ColorJitter(**kw)(imgs)
    ref1 = torchvision.transforms.ColorJitter(**kw)(imgs)
    torch.manual_seed(11)
    aug = T.ColorJitter(**kw)
    got0, got1 = aug(imgs.to(DEV)).cpu(), aug(imgs.to(DEV)).cpu()
    assert (got0 - ref0).abs().max() <= 2e-6 and (got1 - ref1).abs().max() <= 2e-6
    assert not torch.equal(got0, got1)


def test_sequential_pretraining_model_equals_the_classifier_bit_for_bit():
    """[DeepResNet, AdaptiveAvgPool2d, Flatten, Linear] (config/pretraining/resnet50/simple.yaml:23-33) is the same
    computation as DeepResNet(head=True) (classification/models.py:71-77): identical kernels on identical weights."""
    from medsegpretrainimagenet_b200 import models
    torch.manual_seed(0)
    cls = models.kaiming_init_(models.resnet50_classifier(num_classes=16)).to(DEV)
    ffm = models.resnet50_pretraining_model(num_classes=16).to(DEV)
    sd = {k.replace("classifier.2.", "layers.3.") if k.startswith("classifier") else "layers.0." + k: v
          for k, v in cls.state_dict().items()}
    assert set(sd) == set(ffm.state_dict())
    ffm.layers[0].model.load_state_dict({k[9:]: v for k, v in sd.items() if k.startswith("layers.0.")})
    ffm.layers[3].model.load_state_dict({k[9:]: v for k, v in sd.items() if k.startswith("layers.3.")})
    x = torch.randn((4, 3, 64, 64), generator=torch.Generator().manual_seed(3)).to(DEV)
    for m in (cls, ffm):
        m.eval()
    with torch.no_grad():
        a, b = cls(x), ffm(x)
    assert a.shape == b.shape == (4, 16) and torch.equal(a, b)
    cls.train(), ffm.train()
    torch.use_deterministic_algorithms(True, warn_only=True)      # fixed-order BatchNorm statistics: same bits
    try:
        a, b = cls(x), ffm(x)
        assert torch.equal(a, b)
        a.square().mean().backward()
        b.square().mean().backward()
    finally:
        torch.use_deterministic_algorithms(False)
    ga = cls.classifier[2].weight.grad
    gb = ffm.layers[3].model.weight.grad
    assert torch.equal(ga, gb)
    assert torch.equal(cls.stem[0].weight.grad, ffm.layers[0].model.stem[0].weight.grad)


def test_bench_model_equals_the_converted_oracle():
    """the objects bench.py times (`b200.models.*`) were imported by no test.  Same weights -> the
    b200 model and the converted oracle model run the same kernels: bit-identical outputs."""
    import medsegpretrainimagenet_b200 as b200
    from medsegpretrainimagenet_b200 import models
    for make_mine, make_ref, shape in (
            (lambda: models.resnet50_attention_unet(out_ch=4, final_activation="softmax", stochastic_depth_rate=0),
             lambda: ref_models.resnet50_attention_unet(out_ch=4, final_activation="softmax", stochastic_depth_rate=0),
             (2, 3, 64, 64)),
            (lambda: models.basic_unet(out_ch=5), lambda: ref_models.basic_unet(out_ch=5), (1, 3, 64, 64)),
            (lambda: models.resnet50_classifier(num_classes=24), lambda: ref_models.resnet50_classifier(num_classes=24),
             (2, 3, 64, 64))):
        torch.manual_seed(0)
        ref = ref_models.kaiming_init_(make_ref())
        mine = make_mine()
        assert list(mine.state_dict()) == list(ref.state_dict())
        ref_models.load_flat_state_dict(mine, ref.state_dict())
        conv = b200.convert(copy.deepcopy(ref).to(DEV))
        mine = mine.to(DEV)
        x = torch.rand(shape, generator=torch.Generator().manual_seed(4)).to(DEV)
        conv.eval(), mine.eval()
        with torch.no_grad():
            ya, yb = conv(x), mine(x)
        assert torch.equal(ya, yb)


def test_eval_encoder_pipeline_against_the_oracle():
    """robustness/eval.py:56-69 end to end: ColorJitter x2 -> encoder representations of a level -> triplet score, vs
    the oracle encoder on the CPU fed the SAME augmented images."""
    import medsegpretrainimagenet_b200 as b200
    from medsegpretrainimagenet_b200 import models, robustness as R, transforms as T
    torch.manual_seed(0)
    ref_enc = ref_models.kaiming_init_(ref_models.DeepResNet(bottleneck=False, channel_sizes=(64, 128, 256, 512),
                                                             widths=(2, 2, 2, 2), in_channels=3, bias=False))
    enc = models.DeepResNet(bottleneck=False, channel_sizes=(64, 128, 256, 512), widths=(2, 2, 2, 2), in_channels=3,
                            bias=False)
    ref_models.load_flat_state_dict(enc, ref_enc.state_dict())
    model = models.FeedForwardModel([enc]).to(DEV)
    imgs = torch.rand((12, 3, 64, 64), generator=torch.Generator().manual_seed(5))
    for level, pool in ((-2, True), (-1, True), (1, False), (-2, False)):
        torch.manual_seed(21)
        got = R.eval_encoder(model, imgs, R.Robustness("cosine", 0.5), level=level, pool=pool, batch_size=5, device=DEV)
        torch.manual_seed(21)
        aug = T.ColorJitter(brightness=0.1, contrast=0.05, hue=0.05, saturation=0.1)
        a0, a1 = aug(imgs.to(DEV)).cpu(), aug(imgs.to(DEV)).cpu()
        ref_enc.eval()
        with torch.no_grad():
            reps = []
            for a in (a0, a1):
                y, inner = ref_enc(a, return_skip_vals=True)
                r = (list(inner) + [y])[level]
                reps.append(ref_robustness.pooled(r) if pool else r)
        want = ref_robustness.robustness_scores(reps[0], reps[1], ref_robustness.cosine_distance, 0.5)
        assert got.shape == (12,)
        # bf16 encoder vs fp32 encoder on near-identical views: the distances are ~1e-3, compared absolutely
        assert (got.cpu() - want).abs().max() <= 2e-2, (level, pool, (got.cpu() - want).abs().max())
