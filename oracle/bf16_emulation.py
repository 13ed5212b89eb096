"""ORACLE (test infrastructure) — storage-precision emulation of the B200 path on top of the fp32 oracle.

The B200 path keeps activations in bf16 between kernels and does every arithmetic step (convolution
accumulation, BatchNorm statistics / apply, residual add, activation, losses) in fp32.  The fp32 oracle
(oracle/ref_models.py, restating reference classification/models.py and segmentation/models/*.py) run as is
differs from that by one bf16 rounding per stored tensor; on a randomly initialised 50-80 layer ReLU network
those roundings flip ReLU masks and the difference grows chaotically with depth, which says nothing about the
kernels.  `emulate_bf16_storage(model)` installs forward hooks on an oracle model that round exactly the
tensors the B200 path stores in bf16, so the two can be compared tightly over the WHOLE network:

  * the input of every Conv2d / Linear (the stored activation it consumes),
  * the output of every Conv2d / Linear except the segmentation head's final 1x1 convolution (whose fused
    kernel goes straight to the fp32 prediction, reference unet_models.py:442-445, 685-686),
  * the output of every ReLU / Sigmoid (the fused BatchNorm+residual+activation kernel's store),
  * the output of a BatchNorm that is not followed by an activation in the same fused kernel — the `W_g`
    branch of the attention gate (reference blocks.py:598-603, 621).

Weights are expected to be bf16-representable already (round them before cloning the model to the GPU; the
test does).  Gradients flow through the rounding as identity (straight-through), like autograd through a
stored bf16 tensor."""
from __future__ import annotations

import torch
from torch import nn


class _RoundBF16(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        return x.to(torch.bfloat16).to(x.dtype)

    @staticmethod
    def backward(ctx, g):
        return g


def round_bf16(x: torch.Tensor) -> torch.Tensor:
    return _RoundBF16.apply(x)


def round_weights_(model: nn.Module) -> nn.Module:
    """Make every convolution / linear weight bf16-representable in place (BatchNorm affine parameters and
    biases stay fp32: the B200 path consumes them in fp32)."""
    with torch.no_grad():
        for m in model.modules():
            if isinstance(m, (nn.Conv2d, nn.Linear)):
                m.weight.copy_(m.weight.to(torch.bfloat16).float())
    return model


def emulate_bf16_storage(model: nn.Module):
    """Install the rounding hooks; returns the list of hook handles (call .remove() on each to undo)."""
    handles = []
    no_round_out = set()
    for name, m in model.named_modules():
        if name.endswith("final_block") or ".final_block." in name or name.startswith("final_block"):
            for sub in m.modules():
                if isinstance(sub, nn.Conv2d):
                    no_round_out.add(sub)

    def pre(mod, args):
        return (round_bf16(args[0]),) + tuple(args[1:])

    def post(mod, args, out):
        return round_bf16(out)

    for name, m in model.named_modules():
        if isinstance(m, (nn.Conv2d, nn.Linear)):
            handles.append(m.register_forward_pre_hook(pre))
            if m not in no_round_out:
                handles.append(m.register_forward_hook(post))
        elif isinstance(m, (nn.ReLU, nn.Sigmoid)):
            handles.append(m.register_forward_hook(post))
        elif name.split(".")[-1] == "W_g" and isinstance(m, nn.Sequential):
            handles.append(m.register_forward_hook(post))
    return handles
