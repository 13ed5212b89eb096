"""ORACLE (test infrastructure, never imported by the product): plain fp32 PyTorch restatement of the
reference's encoder / U-Net modules, written against the reference sources cited below and pinned to
them by tests/test_oracle_vs_reference.py (live, when /root/reference is present) and by the golden
vectors under tests/golden/ (tools/make_golden.py).

The restatement keeps the reference's *attribute layout and class names* (so `state_dict()` keys are
identical and `medsegpretrainimagenet_b200.convert` walks it exactly like the real thing) but takes plain
keyword arguments instead of the reference's ConfigDict plumbing.
"""
from __future__ import annotations

from typing import Optional, Sequence

import numpy as np
import torch
from torch import nn


class Model(nn.Module):
    """model/model.py:18-75, 248-255: wrapper whose forward drops every input but `x` and whose
    state_dict / parameters delegate to the wrapped module (hence no `.model.` segment in keys)."""

    def __init__(self, inner: nn.Module):
        super().__init__()
        self.model = inner

    def forward(self, x, *args, **kwargs):
        return self.model(x)

    def state_dict(self, *args, **kwargs):
        return self.model.state_dict(*args, **kwargs)

    def load_state_dict(self, state_dict, strict: bool = True):
        return self.model.load_state_dict(state_dict, strict)

    def parameters(self, recurse: bool = True):
        return self.model.parameters(recurse)


def load_flat_state_dict(module: nn.Module, state_dict) -> None:
    """Load a reference checkpoint (keys without the wrappers' `.model.` segment — the reference patches
    the segment back in at model/model.py:203-208) into a tree that contains `Model` wrappers."""
    own = list(module.named_parameters()) + list(module.named_buffers())
    used = set()
    with torch.no_grad():
        for name, t in own:
            key = ("." + name).replace(".model.", ".")[1:]
            t.copy_(state_dict[key])
            used.add(key)
    extra = set(state_dict) - used
    if extra:
        raise KeyError(f"unexpected keys: {sorted(extra)[:5]} ...")


# ------------------------------------------------------------------------------------------------
# classification/models.py
# ------------------------------------------------------------------------------------------------
class DropPath(nn.Module):
    """classification/models.py:313-325: per-sample Bernoulli(keep) mask drawn on the CPU generator,
    NOT rescaled by 1/keep in training; multiplies by keep in eval."""

    def __init__(self, p: float = 0.0):
        super().__init__()
        self.p = p
        self.keep_prob = 1 - p

    def forward(self, x):
        if self.training:
            shape = (x.shape[0],) + (1,) * (x.dim() - 1)
            return torch.bernoulli(self.keep_prob * torch.ones(shape)).to(x.device) * x
        return self.keep_prob * x


def _shortcut(x, downsample: bool, extra_channels: int):
    """classification/models.py:183-200 / 257-274: AvgPool2d(kernel 1, stride 2) == pure sub-sampling,
    then zero-filled widening (no projection convolution)."""
    if downsample:
        x = x[:, :, ::2, ::2]
    if extra_channels > 0:
        z = torch.zeros((x.shape[0], extra_channels) + tuple(x.shape[2:]), device=x.device, dtype=x.dtype)
        x = torch.cat([x, z], dim=1)
    return x


class BasicBlock(nn.Module):
    """classification/models.py:156-212 (stride on conv1)."""

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
        self._ds, self._extra = bool(downsample), out_channels - in_channels

    def forward(self, x):
        y = self.relu1(self.bn1(self.conv1(x)))
        y = self.bn2(self.conv2(y))
        return self.relu2(self.drop_path(y) + _shortcut(x, self._ds, self._extra))


class BottleNeckBlock(nn.Module):
    """classification/models.py:230-290 (1x1 -> 3x3 carrying the stride -> 1x1)."""

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
        self._ds, self._extra = bool(downsample), out_channels - in_channels

    def forward(self, x):
        y = self.relu1(self.bn1(self.conv1(x)))
        y = self.relu2(self.bn2(self.conv2(y)))
        y = self.bn3(self.conv3(y))
        return self.relu3(self.drop_path(y) + _shortcut(x, self._ds, self._extra))


class ResBlock(nn.Sequential):
    """classification/models.py:115-154: `n` units; only the first changes width / resolution."""

    def __init__(self, n, in_channels, out_channels, bottleneck=True, downsample=False, bias=True,
                 drop_probabilities=None):
        unit = BottleNeckBlock if bottleneck else BasicBlock
        probs = (0,) * n if drop_probabilities is None else drop_probabilities
        super().__init__(*[unit(in_channels if i == 0 else out_channels, out_channels,
                                downsample=downsample and i == 0, bias=bias, drop_probability=p)
                           for i, p in enumerate(probs)])


class DeepResNet(nn.Module):
    """classification/models.py:9-103, version 'v1' (every shipped config, e.g.
    config/downstream/acdc/resnet50_attention_unet.yaml:48)."""

    def __init__(self, version="v1", bottleneck=True, channel_sizes=(256, 512, 1024, 2048),
                 widths=(3, 4, 6, 3), in_channels=3, base_channel_size=64, bias=True, head=False,
                 stochastic_depth_rate=0, output_size=None):
        super().__init__()
        if version not in ("v1", 1):
            raise ValueError("the oracle restates DeepResNet v1 only")
        self.version = "v1"
        self.stem = nn.Sequential(nn.Conv2d(in_channels, base_channel_size, 7, stride=2, padding=3, bias=bias),
                                  nn.BatchNorm2d(base_channel_size), nn.ReLU())
        self.max_pool = nn.MaxPool2d(kernel_size=3, stride=2, padding=1)
        probs = np.linspace(0, stochastic_depth_rate or 0, sum(widths))
        self.levels = nn.ModuleList()
        cin = base_channel_size
        for i, (wd, cout) in enumerate(zip(widths, channel_sizes)):
            lo = sum(widths[:i])
            self.levels.append(ResBlock(wd, cin, cout, bottleneck=bottleneck, downsample=bool(i), bias=bias,
                                        drop_probabilities=probs[lo:lo + wd]))
            cin = cout
        if head:
            self.classifier = nn.Sequential(nn.AdaptiveAvgPool2d(1), nn.Flatten(),
                                            nn.Linear(channel_sizes[-1], output_size))
        else:
            self.classifier = nn.Identity()

    def forward(self, x, return_skip_vals=False, *args, **kwargs):
        y = self.stem(x)
        feats = [y]
        y = self.max_pool(y)
        for level in self.levels:
            y = level(y)
            feats.append(y)
        y = self.classifier(y)
        return (y, feats[:-1]) if return_skip_vals else y


# ------------------------------------------------------------------------------------------------
# segmentation/models/blocks.py
# ------------------------------------------------------------------------------------------------
class ConvBlock(nn.Module):
    """blocks.py:452-492: `size` x [Conv2d(bias) -> BatchNorm2d -> ReLU(inplace)]; stride 2 on the last
    conv when `downsample_in_block`."""

    def __init__(self, in_channels, out_channels, size=2, kernel_size=3, padding=1, stride=None,
                 downsample_in_block=False):
        super().__init__()
        layers = []
        for i in range(size):
            s = stride or (2 if (downsample_in_block and i == size - 1) else 1)
            layers += [nn.Conv2d(in_channels if i == 0 else out_channels, out_channels, kernel_size,
                                 stride=s, padding=padding, bias=True),
                       nn.BatchNorm2d(out_channels), nn.ReLU(inplace=True)]
        self.block = nn.Sequential(*layers)

    def forward(self, x):
        return self.block(x)


class UpConvBlock(nn.Module):
    """blocks.py:513-539: nearest x2 -> Conv2d(k=2, padding='same') -> ReLU (no BatchNorm)."""

    def __init__(self, in_channels, out_channels, kernel_size=2, scale_factor=2):
        super().__init__()
        self.convup = nn.Sequential(nn.Upsample(scale_factor=scale_factor),
                                    nn.Conv2d(in_channels, out_channels, kernel_size, stride=1,
                                              padding="same", bias=True),
                                    nn.ReLU(inplace=True))

    def forward(self, x):
        return self.convup(x)


class AttentionBlock(nn.Module):
    """blocks.py:582-628 (attention gate of Attention U-Net) with the default 1x1 ConvBlock gating
    signal (:559-563)."""

    def __init__(self, x_channels, x_up_channels, skip_channels, level_out_channels):
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
        g1 = self.W_g(self.gs_block(x))
        x1 = self.W_s(skip_val)
        p = self.upsample(self.psi(self.relu(x1 + g1)))
        return torch.cat((x_up, skip_val * p), dim=1)


class ConcatBlock(nn.Module):
    """blocks.py:631-635."""

    def __init__(self, **kwargs):
        super().__init__()

    def get_out_ch(self, x_channels, x_up_channels, skip_channels, level_out_channels):
        return x_up_channels + skip_channels

    def forward(self, x, x_up, skip_val):
        return torch.cat((x_up, skip_val), dim=1)


# ------------------------------------------------------------------------------------------------
# segmentation/models/unet_models.py
# ------------------------------------------------------------------------------------------------
class UNet_encoder(nn.Module):
    """unet_models.py:64-236 in its shipped configuration (config/downstream/*/unet.yaml): 3x3 'same'
    stem conv, one ConvBlock per level changing the channel count, MaxPool2d(2) between levels, no
    residual connections / layer scaling / stochastic depth."""

    def __init__(self, in_channel_size=3, depth=4, width=1, channels=None):
        super().__init__()
        self.depth, self.width = depth, width
        ch = list(channels) if channels is not None else [64 * 2 ** i for i in range(depth + 1)]
        if len(ch) < depth + 2:
            ch = [ch[0], *ch]
        self.channels = ch
        self.res_con, self.layer_scale = False, False
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

    def forward(self, x, return_skip_vals=False):
        skips = []
        x = self.first_block(x)
        for unit in self.down_layers:
            for j in range(self.width):
                x = unit[f"conv{j}"](x)
            skips.append(x)
            x = unit["downsampl"](x)
        for j in range(self.width):
            x = self.bottom_block[f"conv{j}"](x)
        return (x, skips) if return_skip_vals else x


class UNet_decoder(nn.Module):
    """unet_models.py:254-390: per level up-conv -> mixing with the popped (deepest-first) skip ->
    `width` ConvBlocks; levels beyond the number of skips have no mixing; final 1x1 conv."""

    def __init__(self, channels: Sequence[int], skip_con_channels_list: Sequence[int], output_ch=1,
                 width=1, attention=False):
        super().__init__()
        self.channels = list(channels)
        self.depth, self.width = len(channels) - 1, width
        self.skip_con_nr = len(skip_con_channels_list)
        self.res_con, self.layer_scale = False, False
        layers = []
        for i in range(self.depth):
            up_out = int(self.channels[i] * 0.5)
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

    def forward(self, x, skip_values):
        skip_values = list(skip_values)
        for i, unit in enumerate(self.up_layers):
            x_up = unit["upsampl"](x)
            if i < self.skip_con_nr:
                x = unit["mixing"](x=x, x_up=x_up, skip_val=skip_values.pop())
            else:
                x = x_up
            for j in range(self.width):
                x = unit[f"conv{j}"](x)
        return self.final_block(x)


class UNet(nn.Module):
    """unet_models.py:591-688.  `encoder=None` -> the basic U-Net (channels 64..1024, concatenate
    mixing); otherwise an external encoder (DeepResNet) with explicit decoder / skip channels and,
    in the shipped ResNet-50 configs, AttentionBlock mixing."""

    def __init__(self, img_ch=3, output_ch=1, depth=4, width=1, channels=None, encoder: Optional[nn.Module] = None,
                 encoder_channels: Optional[Sequence[int]] = None, decoder_channels: Optional[Sequence[int]] = None,
                 skip_con_channels: Optional[Sequence[int]] = None, attention=False,
                 final_activation: Optional[str] = "sigmoid"):
        super().__init__()
        self.final_act = {None: None, "sigmoid": nn.Sigmoid(), "softmax": nn.Softmax(dim=1)}[final_activation]
        self.channels = list(channels) if channels is not None else [64 * 2 ** i for i in range(depth + 1)]
        if encoder is not None:
            enc_ch = list(encoder_channels)
            self.encoder = encoder
        else:
            enc_ch = self.channels
            self.encoder = UNet_encoder(img_ch, depth=len(enc_ch) - 1, width=width, channels=enc_ch)
        dec_ch = self.channels[::-1] if decoder_channels is None else [enc_ch[-1], *decoder_channels]
        skips = list(skip_con_channels) if skip_con_channels else enc_ch[:-1][::-1]
        self.decoder = UNet_decoder(dec_ch, skips, output_ch=output_ch, width=width, attention=attention)

    def forward(self, x):
        x, skips = self.encoder(x, return_skip_vals=True)
        out = self.decoder(x, skips)
        return out if self.final_act is None else self.final_act(out)


# ------------------------------------------------------------------------------------------------
# weight initialisation of the shipped configs (model/model.py:136-198 with
# torch.nn.init.kaiming_normal_(a=0, mode='fan_in', nonlinearity='relu'),
# config/downstream/covidqu/unet.yaml:38-43): every module owning a >=2-D `.weight` gets
# kaiming-normal weights and zero bias; BatchNorm (1-D weight) keeps its defaults.
# ------------------------------------------------------------------------------------------------
def kaiming_init_(model: nn.Module) -> nn.Module:
    for m in model.modules():
        w = getattr(m, "weight", None)
        if isinstance(w, torch.Tensor) and w.dim() >= 2:
            nn.init.kaiming_normal_(w, a=0, mode="fan_in", nonlinearity="relu")
            if getattr(m, "bias", None) is not None:
                nn.init.zeros_(m.bias)
    return model


# ------------------------------------------------------------------------------------------------
# the BASELINE.json configurations
# ------------------------------------------------------------------------------------------------
def resnet50_classifier(num_classes=1000, in_channels=3):
    """cfg2: config/pretraining/resnet50/simple.yaml:24-33 (DeepResNet -> avgpool -> flatten -> linear;
    SURVEY.md §8c: identical maths to DeepResNet(head=True))."""
    return DeepResNet(bias=False, head=True, output_size=num_classes, in_channels=in_channels)


def resnet50_attention_unet(out_ch=1, final_activation="sigmoid", in_channels=3, stochastic_depth_rate=0.1):
    """cfg3: config/downstream/acdc/resnet50_attention_unet.yaml:26-54."""
    enc = DeepResNet(bias=False, head=False, in_channels=in_channels, stochastic_depth_rate=stochastic_depth_rate)
    return UNet(img_ch=in_channels, output_ch=out_ch, encoder=enc, encoder_channels=(256, 512, 1024, 2048),
                decoder_channels=(256, 128, 64, 32, 16), skip_con_channels=(1024, 512, 256, 64),
                attention=True, final_activation=final_activation)


def resnet18_attention_unet(out_ch=1, final_activation="sigmoid", in_channels=1):
    """cfg1: ResNet-18-shaped DeepResNet encoder + attention decoder."""
    enc = DeepResNet(bottleneck=False, channel_sizes=(64, 128, 256, 512), widths=(2, 2, 2, 2),
                     in_channels=in_channels, bias=False)
    return UNet(img_ch=in_channels, output_ch=out_ch, encoder=enc, encoder_channels=(64, 128, 256, 512),
                decoder_channels=(256, 128, 64, 32, 16), skip_con_channels=(256, 128, 64, 64),
                attention=True, final_activation=final_activation)


def basic_unet(out_ch=1, final_activation="sigmoid", in_channels=3):
    """cfg4: config/downstream/idrid/unet.yaml (UNet defaults, unet_models.py:413-497)."""
    return UNet(img_ch=in_channels, output_ch=out_ch, final_activation=final_activation)
