"""helicon_b200: B200-native (sm_100a CUDA) implementation of helicon's denovo3D
solve+score hot path behind the reference's own Python signatures.

    from helicon_b200 import solver_linear_regression as solver   # drop-in module
    from helicon_b200 import search_grid                          # batched grid driver
"""

__version__ = "0.1.0"

from ._lib import HeliconB200Error  # noqa: F401
