"""myfm_b200 — B200-native Gibbs engine for Bayesian Factorization Machines behind the Python API
of tohtsky/myFM (MyFMRegressor / MyFMClassifier / MyFMOrderedProbit, RelationBlock, group_shapes).

The per-iteration hot path runs as hand-written sm_100a CUDA behind a C ABI
(include/myfm_b200.h); see DESIGN.md.
"""
from ._myfm import RelationBlock
from .gibbs import MyFMGibbsClassifier, MyFMGibbsRegressor, MyFMOrderedProbit
from .options import engine_options, get_options, set_options

__version__ = "0.1.0"

MyFMRegressor = MyFMGibbsRegressor
MyFMClassifier = MyFMGibbsClassifier

__all__ = [
    "RelationBlock",
    "MyFMOrderedProbit",
    "MyFMRegressor",
    "MyFMClassifier",
    "MyFMGibbsRegressor",
    "MyFMGibbsClassifier",
    "engine_options",
    "get_options",
    "set_options",
]
