"""Model classes with the reference's structure (attribute names -> identical `state_dict()` keys, so the
published encoder checkpoints and the reference's own `load_state_dict` paths apply) whose forward runs
exclusively on the B200 kernels.  They exist so the hot path can be constructed WITHOUT the reference
checkout (bench.py, smoke tests, YAML class paths such as
`medsegpretrainimagenet_b200.models.DeepResNet`); with the reference present, `convert()` /
`patch.install()` route the reference's own modules through the same interpreter instead.

Constructor arguments follow classification/models.py:11-13 (DeepResNet) and, for the U-Net, the
resolved values of segmentation/models/unet_models.py:591-678 as plain keywords.
"""
from __future__ import annotations

from typing import Optional, Sequence

import numpy as np
import torch
from torch import nn

from . import converter as _cv
from . import functional as _Fn


class _B200Module(nn.Module):
    """A structural block: calling it on an NCHW fp32 CUDA tensor runs the B200 path for that block."""

    def forward(self, x, *args, **kwargs):
        _cv._require_cuda(x)
        y = _cv.run_module(_cv.ExecContext(), self, _Fn.to_nhwc(x))
        return _Fn.to_nchw(y)


class Model(_B200Module):
    """model/model.py:18-75 wrapper (`.model`; state_dict / parameters delegate, :248-255)."""

    def __init__(self, inner: nn.Module):
        super().__init__()
        self.model = inner

    def state_dict(self, *args, **kwargs):
        return self.model.state_dict(*args, **kwargs)

    def parameters(self, recurse: bool = True):
        return self.model.parameters(recurse)


class DropPath(nn.Module):
    """classification/models.py:313-325 (parameters only; the multiply is fused into the BN kernel)."""

    def __init__(self, p: float = 0.0):
        super().__init__()
        self.p, self.keep_prob = p, 1 - p


class BasicBlock(_B200Module):
    """classification/models.py:156-212."""

    def __init__(self, in_channels, out_channels, downsample=False, bias=True, drop_probability=0.0):
        super().__init__()
        if out_channels < in_channels:
            raise ValueError("Out channel size should not be smaller than in channel size.")
        self.conv1 = nn.Conv2d(in_channels, out_channels, 3, stride=2 if downsample else 1, padding=1, bias=bias)
        self.bn1 = nn.BatchNorm2d(out_channels)
        self.relu1 = nn.ReLU()
        self.conv2 = nn.Conv2d(out_channels, out_channels, 3, stride=1, padding=1, bias=bias)
        self.bn2 = nn.BatchNorm2d(out_channels)
        self.relu2 = nn.ReLU()
        self.drop_path = nn.Identity() if drop_probability == 0 else DropPath(drop_probability)


class BottleNeckBlock(_B200Module):
    """classification/models.py:230-290."""

    def __init__(self, in_channels, out_channels, downsample=False, bias=True, drop_probability=0.0):
        super().__init__()
        if out_channels < in_channels:
            raise ValueError("Out channel size should not be smaller than in channel size.")
        mid = out_channels // 4
        self.conv1 = nn.Conv2d(in_channels, mid, 1, bias=bias)
        self.bn1 = nn.BatchNorm2d(mid)
        self.relu1 = nn.ReLU()
        self.conv2 = nn.Conv2d(mid, mid, 3, padding=1, stride=2 if downsample else 1, bias=bias)
        self.bn2 = nn.BatchNorm2d(mid)
        self.relu2 = nn.ReLU()
        self.conv3 = nn.Conv2d(mid, out_channels, 1, bias=bias)
        self.bn3 = nn.BatchNorm2d(out_channels)
        self.relu3 = nn.ReLU()
        self.drop_path = nn.Identity() if drop_probability == 0 else DropPath(drop_probability)


class ResBlock(nn.Sequential):
    """classification/models.py:115-154."""

    def __init__(self, size, in_channels, out_channels, version="v1", bottleneck=True, downsample=False,
                 bias=True, drop_probabilities=None):
        unit_size = 3 if bottleneck else 2
        if size % unit_size:
            raise ValueError(f"Size of residual block must be divisible by {unit_size}, but got {size}.")
        n = size // unit_size
        probs = (0,) * n if drop_probabilities is None else drop_probabilities
        if len(probs) != n:
            raise ValueError("Number of drop probabilities given must equal the number of blocks")
        unit = BottleNeckBlock if bottleneck else BasicBlock
        super().__init__(*[unit(in_channels if i == 0 else out_channels, out_channels,
                                downsample=downsample and i == 0, bias=bias, drop_probability=p)
                           for i, p in enumerate(probs)])


class DeepResNet(nn.Module):
    """classification/models.py:9-103 (version 'v1').  `forward(x, return_skip_vals=False)` returns
    fp32 NCHW tensors exactly like the reference: `y` or `(y, [stem, level0, level1, level2])`."""

    def __init__(self, version="v1", bottleneck=True, channel_sizes=(256, 512, 1024, 2048),
                 widths=(3, 4, 6, 3), in_channels=3, base_channel_size=64, bias=True, head=False,
                 stochastic_depth_rate=0, group=None, *args, **kwargs):
        super().__init__()
        if isinstance(version, int):
            version = f"v{version}"
        if version != "v1":
            raise _cv.UnsupportedModule("DeepResNet v2 (pre-activation) is not on the B200 path")
        if len(widths) != len(channel_sizes):
            raise ValueError("Each level of the ResNet needs one channel size and one width")
        self.version, self.bottleneck, self.channel_sizes, self.widths = version, bottleneck, channel_sizes, widths
        self.in_channels, self.base_channel_size, self.bias, self.head = in_channels, base_channel_size, bias, head
        self.stochastic_depth_rate = stochastic_depth_rate
        self.stem = nn.Sequential(nn.Conv2d(in_channels, base_channel_size, 7, stride=2, padding=3, bias=bias),
                                  nn.BatchNorm2d(base_channel_size), nn.ReLU())
        self.max_pool = nn.MaxPool2d(kernel_size=3, stride=2, padding=1)
        unit_size = 3 if bottleneck else 2
        probs = np.linspace(0, stochastic_depth_rate or 0, sum(widths))
        self.levels = nn.ModuleList()
        cin = base_channel_size
        for i, (wd, cout) in enumerate(zip(widths, channel_sizes)):
            lo = sum(widths[:i])
            self.levels.append(ResBlock(wd * unit_size, cin, cout, bottleneck=bottleneck, downsample=bool(i),
                                        bias=bias, drop_probabilities=probs[lo:lo + wd]))
            cin = cout
        if head:
            self.output_size = kwargs["output_size"]
            self.classifier = nn.Sequential(nn.AdaptiveAvgPool2d(output_size=1), nn.Flatten(),
                                            nn.Linear(channel_sizes[-1], kwargs["output_size"]))
        else:
            self.classifier = nn.Identity()
        _cv.convert(self, group=group)


class ConvBlock(_B200Module):
    """segmentation/models/blocks.py:452-492 with ReLU activations."""

    def __init__(self, in_channels, out_channels, size=2, kernel_size=3, padding=1, activations="relu",
                 dropout=False, stride=None, downsample_in_block=False, *args, **kwargs):
        super().__init__()
        if activations != "relu" or dropout:
            raise _cv.UnsupportedModule("ConvBlock: only ReLU activations without dropout are on the B200 path")
        layers = []
        for i in range(size):
            s = stride or (2 if (downsample_in_block and i == size - 1) else 1)
            layers += [nn.Conv2d(in_channels if i == 0 else out_channels, out_channels, kernel_size, stride=s,
                                 padding=padding, bias=True),
                       nn.BatchNorm2d(out_channels), nn.ReLU(inplace=True)]
        self.block = nn.Sequential(*layers)


class UpConvBlock(_B200Module):
    """segmentation/models/blocks.py:513-539."""

    def __init__(self, in_channels, out_channels, activation="relu", kernel_size=2, scale_factor=2, *args, **kwargs):
        super().__init__()
        if activation != "relu" or scale_factor != 2:
            raise _cv.UnsupportedModule("UpConvBlock: nearest x2 + ReLU only")
        self.convup = nn.Sequential(nn.Upsample(scale_factor=scale_factor),
                                    nn.Conv2d(in_channels, out_channels, kernel_size, stride=1, padding="same",
                                              bias=True),
                                    nn.ReLU(inplace=True))


class AttentionBlock(nn.Module):
    """segmentation/models/blocks.py:582-628 with the default 1x1-ConvBlock gating signal."""

    def __init__(self, x_channels, x_up_channels, skip_channels, level_out_channels, *args, **kwargs):
        super().__init__()
        self.gs_block = Model(ConvBlock(x_channels, x_channels, size=1, kernel_size=1, padding=0))
        self.W_g = nn.Sequential(nn.Conv2d(x_channels, x_channels, 1, 1, padding=0, bias=True),
                                 nn.BatchNorm2d(x_channels))
        self.W_s = nn.Sequential(nn.Conv2d(skip_channels, x_channels, 2, 2, padding=0, bias=True),
                                 nn.BatchNorm2d(x_channels))
        self.psi = nn.Sequential(nn.Conv2d(x_channels, skip_channels, 1, 1, padding=0, bias=True),
                                 nn.BatchNorm2d(skip_channels), nn.Sigmoid())
        self.upsample = nn.Upsample(scale_factor=2)
        self.relu = nn.ReLU()

    def get_out_ch(self, x_channels, x_up_channels, skip_channels, level_out_channels):
        return x_up_channels + skip_channels

    def forward(self, x, x_up, skip_val):
        for t in (x, x_up, skip_val):
            _cv._require_cuda(t)
        y = _cv.run_attention_block(_cv.ExecContext(), self, _Fn.to_nhwc(x), _Fn.to_nhwc(x_up),
                                    _Fn.to_nhwc(skip_val))
        return _Fn.to_nchw(y)


class ConcatBlock(nn.Module):
    """segmentation/models/blocks.py:631-635."""

    def __init__(self, **kwargs):
        super().__init__()

    def get_out_ch(self, x_channels, x_up_channels, skip_channels, level_out_channels):
        return x_up_channels + skip_channels

    def forward(self, x, x_up, skip_val):
        return _Fn.to_nchw(_Fn.concat(_Fn.to_nhwc(x_up), _Fn.to_nhwc(skip_val)))


class UNet_encoder(nn.Module):
    """segmentation/models/unet_models.py:64-236 in the shipped configuration (3x3 'same' stem, one ConvBlock
    per level, MaxPool2d(2) between levels, no residual connections / layer scaling)."""

    def __init__(self, in_channel_size=3, depth=4, width=1, channels=None, group=None, *args, **kwargs):
        super().__init__()
        self.depth, self.width = depth, width
        ch = list(channels) if channels not in (None, "default") else [64 * 2 ** i for i in range(depth + 1)]
        if len(ch) < depth + 2:
            ch = [ch[0], *ch]
        self.channels = ch
        self.res_con, self.layer_scale, self.integrated_downsample = False, False, False
        self.first_block = Model(nn.Conv2d(in_channel_size, ch[0], kernel_size=3, padding="same"))
        self.down_layers = nn.ModuleList()
        for i in range(depth):
            unit = {"conv0": Model(ConvBlock(ch[i], ch[i + 1]))}
            for j in range(1, width):
                unit[f"conv{j}"] = Model(ConvBlock(ch[i + 1], ch[i + 1]))
            unit["downsampl"] = Model(nn.MaxPool2d(kernel_size=2))
            self.down_layers.append(nn.ModuleDict(unit))
        bottom = {"conv0": Model(ConvBlock(ch[-2], ch[-1]))}
        for j in range(1, width):
            bottom[f"conv{j}"] = Model(ConvBlock(ch[-1], ch[-1]))
        self.bottom_block = nn.ModuleDict(bottom)
        _cv.convert(self, group=group)


class UNet_decoder(nn.Module):
    """segmentation/models/unet_models.py:254-390 (structure only; executed by UNet.forward)."""

    def __init__(self, channels: Sequence[int], skip_con_channels_list: Sequence[int], output_ch=1, width=1,
                 attention=False, upsample_channel_decrease_ratio=0.5):
        super().__init__()
        self.channels = list(channels)
        self.depth, self.width = len(channels) - 1, width
        self.skip_con_nr = len(skip_con_channels_list)
        self.res_con, self.layer_scale = False, False
        layers = []
        for i in range(self.depth):
            up_out = int(self.channels[i] * upsample_channel_decrease_ratio)
            unit = {"upsampl": Model(UpConvBlock(self.channels[i], up_out))}
            mix_out = up_out
            if i < self.skip_con_nr:
                kw = dict(x_channels=self.channels[i], x_up_channels=up_out,
                          skip_channels=skip_con_channels_list[i], level_out_channels=self.channels[i + 1])
                unit["mixing"] = AttentionBlock(**kw) if attention else ConcatBlock(**kw)
                mix_out = unit["mixing"].get_out_ch(**kw)
            unit["conv0"] = Model(ConvBlock(mix_out, self.channels[i + 1]))
            for j in range(1, width):
                unit[f"conv{j}"] = Model(ConvBlock(self.channels[i + 1], self.channels[i + 1]))
            layers.append(nn.ModuleDict(unit))
        self.up_layers = nn.ModuleList(layers)
        self.final_block = Model(nn.Conv2d(self.channels[-1], output_ch, kernel_size=1))


class UNet(nn.Module):
    """segmentation/models/unet_models.py:591-688.  `encoder=None`: the basic U-Net; otherwise an external
    encoder exposing `forward(x, return_skip_vals=True)` (DeepResNet) with explicit channel lists."""

    def __init__(self, img_ch=3, output_ch=1, depth=4, width=1, channels=None, encoder: Optional[nn.Module] = None,
                 encoder_channels: Optional[Sequence[int]] = None, decoder_channels: Optional[Sequence[int]] = None,
                 skip_con_channels: Optional[Sequence[int]] = None, mixing_block="concatenate",
                 final_activation: Optional[str] = "sigmoid", group=None, *args, **kwargs):
        super().__init__()
        if final_activation not in (None, "sigmoid", "softmax"):
            raise _cv.UnsupportedModule(f"final activation {final_activation!r}")
        self.final_act = {None: None, "sigmoid": nn.Sigmoid(), "softmax": nn.Softmax(dim=1)}[final_activation]
        self.depth, self.width = depth, width
        self.channels = list(channels) if channels not in (None, "default") else [64 * 2 ** i for i in range(depth + 1)]
        if encoder is not None:
            if encoder_channels is None:
                encoder_channels = getattr(encoder, "channel_sizes")
            self.encoder_channels = list(encoder_channels)
            self.encoder = encoder
        else:
            self.encoder_channels = self.channels if encoder_channels is None else list(encoder_channels)
            self.encoder = UNet_encoder(img_ch, depth=len(self.encoder_channels) - 1, width=width,
                                        channels=self.encoder_channels, group=group)
        self.decoder_channels = (self.channels[::-1] if decoder_channels is None
                                 else [self.encoder_channels[-1], *decoder_channels])
        skips = list(skip_con_channels) if skip_con_channels else self.encoder_channels[:-1][::-1]
        attention = mixing_block not in ("concatenate", None) and "Attention" in str(mixing_block)
        if mixing_block not in ("concatenate", None) and not attention:
            raise _cv.UnsupportedModule(f"mixing block {mixing_block!r}")
        self.decoder = UNet_decoder(self.decoder_channels, skips, output_ch=output_ch, width=width,
                                    attention=attention)
        _cv.convert(self, group=group)


def kaiming_init_(model: nn.Module) -> nn.Module:
    """The shipped configs' initialisation (model/model.py:136-198 with torch.nn.init.kaiming_normal_,
    config/downstream/covidqu/unet.yaml:38-43): kaiming-normal on every >=2-D `.weight`, zero bias."""
    for m in model.modules():
        w = getattr(m, "weight", None)
        if isinstance(w, torch.Tensor) and w.dim() >= 2:
            nn.init.kaiming_normal_(w, a=0, mode="fan_in", nonlinearity="relu")
            if getattr(m, "bias", None) is not None:
                nn.init.zeros_(m.bias)
    return model


class FeedForwardModel(nn.Module):
    """The sequential compound model the pretraining YAMLs name (`model.FeedForwardModel: {layers: [...]}`,
    config/pretraining/resnet50/simple.yaml:23-33): `.layers` is a ModuleList applied in order, so an encoder trained
    as its first layer is checkpointed under the `layers.0.` prefix that `UNet.init_weights` strips
    (segmentation/models/unet_models.py:570-571) and `eval_encoder` reads as `model.layers[0]` (robustness/eval.py:58).
    The class shipped in the reference (model/model.py:313-333) takes `threads` and runs its sub-models in PARALLEL
    (SURVEY.md App. C), so those YAMLs cannot be built from the checkout; `patch.install()` puts a class with this
    behaviour in its place.  Each layer is wrapped in `Model` like the reference's CompoundModel does
    (model/model.py:299-305); `state_dict()` keys carry no `.model.` segment (model/model.py:248-249).

    Supported layer lists: [DeepResNet] or [DeepResNet, AdaptiveAvgPool2d(1), Flatten, Linear] — executed as ONE pass
    of the B200 interpreter (the head is a pooled 1x1 convolution on the tap-GEMM kernel), never through torch.nn."""

    def __init__(self, layers: Sequence[nn.Module], group=None, *args, **kwargs):
        super().__init__()
        if layers is None:
            layers = []
        if not isinstance(layers, (tuple, list)):
            layers = [layers]
        self.layers = nn.ModuleList([l if type(l).__name__ == "Model" else Model(l) for l in layers])
        self.pass_all_inputs = [False] * len(self.layers)
        self.PASS_ALL_INPUTS = False
        _cv.convert(self, group=group)


def encoder_state_dict(state_dict: dict, prefix: str = "layers.0") -> dict:
    """unet_models.py:570-571: the entries of a pretraining checkpoint that belong to its first layer."""
    n = len(prefix)
    return {k[n + 1:]: v for k, v in state_dict.items() if k[:n] == prefix}


def load_encoder_checkpoint(unet: nn.Module, checkpoint, strict: bool = True):
    """`UNet.init_weights` (segmentation/models/unet_models.py:555-588) for models built without the reference:
    load the `layers.0.` entries of a pretraining checkpoint (path or state dict) into `unet.encoder`, mapping keys
    saved without the wrappers' `.model.` segments back onto them (:573-577).  Returns (missing, unexpected)."""
    sd = torch.load(checkpoint, map_location="cpu") if isinstance(checkpoint, (str, bytes)) or hasattr(checkpoint, "__fspath__") \
        else checkpoint
    enc_sd = encoder_state_dict(sd)
    encoder = unet.encoder
    missing, _ = encoder.load_state_dict(enc_sd, strict=False)
    for key in list(missing):
        short = key.replace(".model.", ".")
        if short in enc_sd:
            enc_sd[key] = enc_sd.pop(short)
    res = encoder.load_state_dict(enc_sd, strict=strict)
    return list(res.missing_keys), list(res.unexpected_keys)


# the BASELINE.json configurations -----------------------------------------------------------------
def resnet50_classifier(num_classes=1000, in_channels=3, group=None):
    """cfg2 — config/pretraining/resnet50/simple.yaml:24-33."""
    return DeepResNet(bias=False, head=True, output_size=num_classes, in_channels=in_channels, group=group)


def resnet50_pretraining_model(num_classes=1000, in_channels=3, stochastic_depth_rate=0, group=None):
    """cfg2 exactly as its YAML lists it (config/pretraining/resnet50/simple.yaml:23-33): the sequential model
    [DeepResNet(bias=False, v1), AdaptiveAvgPool2d(1), Flatten, Linear(2048, num_classes)]."""
    enc = DeepResNet(bias=False, version="v1", in_channels=in_channels, stochastic_depth_rate=stochastic_depth_rate,
                     group=group)
    return FeedForwardModel([enc, nn.AdaptiveAvgPool2d(output_size=1), nn.Flatten(), nn.Linear(2048, num_classes)],
                            group=group)


def resnet50_attention_unet(out_ch=1, final_activation="sigmoid", in_channels=3, stochastic_depth_rate=0.1,
                            group=None):
    """cfg3 — config/downstream/acdc/resnet50_attention_unet.yaml:26-54."""
    enc = DeepResNet(bias=False, head=False, in_channels=in_channels, stochastic_depth_rate=stochastic_depth_rate,
                     group=group)
    return UNet(img_ch=in_channels, output_ch=out_ch, encoder=enc, encoder_channels=(256, 512, 1024, 2048),
                decoder_channels=(256, 128, 64, 32, 16), skip_con_channels=(1024, 512, 256, 64),
                mixing_block="segmentation.models.blocks.AttentionBlock", final_activation=final_activation,
                group=group)


def resnet18_attention_unet(out_ch=1, final_activation="sigmoid", in_channels=1, group=None):
    """cfg1 — SURVEY.md §8d."""
    enc = DeepResNet(bottleneck=False, channel_sizes=(64, 128, 256, 512), widths=(2, 2, 2, 2),
                     in_channels=in_channels, bias=False, group=group)
    return UNet(img_ch=in_channels, output_ch=out_ch, encoder=enc, encoder_channels=(64, 128, 256, 512),
                decoder_channels=(256, 128, 64, 32, 16), skip_con_channels=(256, 128, 64, 64),
                mixing_block="segmentation.models.blocks.AttentionBlock", final_activation=final_activation,
                group=group)


def basic_unet(out_ch=1, final_activation="sigmoid", in_channels=3, group=None):
    """cfg4 — config/downstream/idrid/unet.yaml."""
    return UNet(img_ch=in_channels, output_ch=out_ch, final_activation=final_activation, group=group)
