"""ORACLE (test infrastructure): deterministic weights for golden fixtures.

`fill_state_(module, seed)` overwrites every tensor of `module.state_dict()` in key order from one seeded
generator, so that the real reference model (tools/make_golden.py, build container) and its oracle restatement
(tests, any machine) hold IDENTICAL weights without shipping them: their state_dict keys, order and shapes are
the same (tests/test_oracle_vs_reference.py).  Scales are chosen to keep activations O(1) through 50+ layers:
conv / linear weights ~ N(0, 2/fan_in) (the reference's kaiming_normal_ scheme, model/model.py:142-148),
BatchNorm affine and running statistics perturbed away from their defaults."""
from __future__ import annotations

import math

import torch


def fill_state_(module: torch.nn.Module, seed: int = 0) -> torch.nn.Module:
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        for key, t in module.state_dict().items():
            if not t.is_floating_point():
                continue                                     # num_batches_tracked
            if t.dim() >= 2:
                fan_in = t[0].numel()
                v = torch.randn(t.shape, generator=g) * math.sqrt(2.0 / fan_in)
            elif key.endswith("running_var"):
                v = 1.0 + 0.2 * torch.rand(t.shape, generator=g)
            elif key.endswith("running_mean"):
                v = 0.1 * torch.randn(t.shape, generator=g)
            elif key.endswith("weight"):                     # BatchNorm gamma
                v = 1.0 + 0.1 * torch.randn(t.shape, generator=g)
            else:                                            # biases, BatchNorm beta
                v = 0.1 * torch.randn(t.shape, generator=g)
            t.copy_(v)
    return module
