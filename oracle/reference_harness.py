"""TEST INFRASTRUCTURE ONLY — imports the *real* reference (aielte-research/MedSegPretrainImageNet) from
/root/reference/src so that the oracle restatement (oracle/ref_*.py) can be pinned against it and golden
vectors can be minted (tools/make_golden.py).  /root/reference does not exist on the GPU box: everything
here is reachable only from `-m "not gpu"` tests (which skip when it is absent) and from tools/.

Recipe: stub the plotting / augmentation dependencies that are absent from this
container (never the arithmetic), shim the py3.12 `random.randint(0, 1e16)` default argument of
src/run_experiment.py:35, then build objects exactly as src/run_experiment.py:46-50,111-119,222,282-332.
"""
from __future__ import annotations

import importlib
import os
import random
import sys
import types
from unittest import mock

REFERENCE_SRC = os.environ.get("MSP_REFERENCE_SRC", "/root/reference/src")

_STUBS = [
    "matplotlib", "matplotlib.pyplot", "matplotlib.colors", "mpl_toolkits", "mpl_toolkits.axes_grid1",
    "colorcet", "plotly", "plotly.express", "bokeh", "bokeh.colors", "bokeh.io", "bokeh.layouts",
    "bokeh.models", "bokeh.models.ranges", "bokeh.plotting", "bokeh.transform", "bokeh.palettes",
    "fvcore", "fvcore.nn", "albumentations", "albumentations.augmentations",
    "albumentations.augmentations.geometric", "albumentations.augmentations.geometric.rotate",
    "albumentations.augmentations.crops", "albumentations.augmentations.crops.transforms", "nibabel",
]


def available() -> bool:
    return os.path.isdir(REFERENCE_SRC)


_ready = False


def setup() -> None:
    """Make `import model, loss, metrics, ...` resolve to the reference's packages."""
    global _ready
    if _ready:
        return
    if not available():
        raise RuntimeError(f"reference sources not found at {REFERENCE_SRC}")
    sys.dont_write_bytecode = True
    for name in _STUBS:
        try:
            importlib.import_module(name)
        except Exception:
            sys.modules[name] = mock.MagicMock(name=name)
    try:
        importlib.import_module("timm.models.layers")
    except Exception:
        import torch

        class DropPath(torch.nn.Module):  # Swin-only dependency (blocks.py:2); never on the hot path
            def __init__(self, drop_prob=0.0):
                super().__init__()
                self.drop_prob = drop_prob

            def forward(self, x):
                return x

        timm = types.ModuleType("timm")
        timm_models = types.ModuleType("timm.models")
        layers = types.ModuleType("timm.models.layers")
        layers.DropPath = DropPath
        layers.trunc_normal_ = torch.nn.init.trunc_normal_
        timm.models, timm_models.layers = timm_models, layers
        sys.modules.update({"timm": timm, "timm.models": timm_models, "timm.models.layers": layers})
    if REFERENCE_SRC not in sys.path:
        sys.path.insert(0, REFERENCE_SRC)
    _orig = random.randint
    random.randint = lambda a, b: _orig(int(a), int(b))  # run_experiment.py:35 on Python >= 3.12
    try:
        importlib.import_module("run_experiment")
    finally:
        random.randint = _orig
    _ready = True


def load_config(yaml_rel_path: str, overrides: dict | None = None, grid_index: int = 0):
    """YAML -> expanded, default-filled ConfigDict (run_experiment.py:46-50, 70, 111-119)."""
    setup()
    import yaml
    import utils
    import data
    import model
    import optim
    import metrics
    from utils.config_dict import ConfigDict

    path = yaml_rel_path if os.path.isabs(yaml_rel_path) else os.path.join(
        os.path.dirname(REFERENCE_SRC), "config", yaml_rel_path)
    from utils import config_parser
    grid, _ = config_parser.parse(path)  # YAML lists are sweep axes (config_parser.py:5-16)
    cd = ConfigDict(grid[grid_index])
    for k, v in (overrides or {}).items():
        cd[k] = v
    cd.expand()
    cd.fill_with_defaults(utils.default_dict)
    cd["meta/technical"] = cd["meta/technical"].trim()
    cd["meta/technical/log_to_device"] = False
    data.BalancedDataLoader.fill_kwargs(cd.get_or_update("data/sampling", ConfigDict({})))
    for key in ("model", "training/loss"):
        utils.fill_dict(cd, key)
    model.Model.fill_weight_init_kwargs(cd["model"].value())
    optim.Optimizer.fill_kwargs(cd["training/optimizer"])
    metrics.MetricsCalculator.fill_kwargs(cd)
    return cd.trim()


def build_model(cd, seed: int = 0, init: bool = True):
    """run_experiment.py:277-292."""
    setup()
    import numpy as np
    import torch
    import utils
    import model

    torch.manual_seed(seed)
    np.random.seed(seed)
    random.seed(seed)
    m = utils.create_object_from_dict(cd, key="model", wrapper_class=model.Model)
    if init:
        m.init_weight(cd["model"].value())
        m.freeze_and_unfreeze(cd["model"].value())
    return m


def build_loss(cd):
    setup()
    import utils
    import loss
    return utils.create_object_from_dict(cd, key="training/loss", wrapper_class=loss.Loss)


def build_metrics(cd, loss_fn, class_names=()):
    setup()
    import metrics
    return metrics.MetricsCalculator(cd, validate=True, exp_name="run_1", loss=loss_fn,
                                     class_names=class_names)
