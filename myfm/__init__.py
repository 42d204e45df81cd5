"""`import myfm` drop-in: the reference's package name over the B200 engine (myfm_b200).

Code written against tohtsky/myFM — `from myfm import MyFMRegressor, RelationBlock`,
`from myfm._myfm import create_train_fm`, `from myfm.utils.callbacks import RegressionCallback`,
`myfm.gibbs`, `myfm.base` — resolves to the modules of `myfm_b200` with the same names.  The
variational estimators of the reference (src/myfm/variational.py) are outside the accelerated path
and are not provided.
"""
import importlib as _importlib
import sys as _sys

import myfm_b200 as _impl
from myfm_b200 import *  # noqa: F401,F403
from myfm_b200 import __all__, __version__  # noqa: F401

for _name in ("_myfm", "base", "gibbs", "options", "utils", "utils.callbacks", "utils.callbacks.libfm"):
    _mod = _importlib.import_module("myfm_b200." + _name)
    _sys.modules[__name__ + "." + _name] = _mod
    if "." not in _name:
        globals()[_name] = _mod
del _name, _mod
