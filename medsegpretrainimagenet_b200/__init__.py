"""B200-native (sm_100a) hot path for aielte-research/MedSegPretrainImageNet.

Importing the package loads libmsp_b200.so (hand-written CUDA kernels behind the C ABI of
include/msp_b200.h) and fails if it has not been built: there is no PyTorch / CPU fallback.

    from medsegpretrainimagenet_b200 import convert
    convert(model)              # reference `model.Model` / UNet / DeepResNet -> B200 kernels, in place
"""
from . import _lib  # noqa: F401  (raises ImportError when the library is missing)
from .converter import ExecContext, UnsupportedModule, convert, is_converted
from . import converter, functional, graphs, losses, metrics, ops, optim, robustness, transforms
from .graphs import GraphedStep
from .host import BatchPrefetcher, ScalarReader

__all__ = ["convert", "is_converted", "ExecContext", "UnsupportedModule", "GraphedStep", "BatchPrefetcher", "ScalarReader", "functional", "graphs", "losses", "metrics",
           "ops", "optim", "robustness", "transforms"]
