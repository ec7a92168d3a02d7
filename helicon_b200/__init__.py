"""helicon_b200: B200-native (sm_100a CUDA) implementation of helicon's denovo3D
solve+score hot path behind the reference's own Python signatures.

    from helicon_b200 import solver_linear_regression as solver   # drop-in module
    from helicon_b200 import search_grid                          # batched grid driver
"""

__version__ = "0.2.0"

from ._lib import HeliconB200Error  # noqa: F401


def search_grid(*args, **kwargs):
    """``helicon_b200.grid.search_grid`` (imported on first use: the grid driver pulls in the CUDA library binding)."""
    from .grid import search_grid as _search_grid

    return _search_grid(*args, **kwargs)
