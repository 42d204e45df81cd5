from .libfm import (
    ClassificationCallback,
    LibFMLikeCallbackBase,
    OrderedProbitCallback,
    RegressionCallback,
)

__all__ = [
    "LibFMLikeCallbackBase",
    "OrderedProbitCallback",
    "ClassificationCallback",
    "RegressionCallback",
]
