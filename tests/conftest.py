import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN


@pytest.fixture(autouse=True)
def _fresh_refined_params():
    """``lsq_reconstruct._refined_params`` is function-attribute state that the NEXT process_one_task consumes -- in the
    reference as well (pipeline.py:428-436); tests must not inherit it from one another."""
    import sys

    mod = sys.modules.get("helicon_b200.solver_linear_regression")
    if mod is not None and hasattr(mod.lsq_reconstruct, "_refined_params"):
        mod.lsq_reconstruct._refined_params = {}
    yield
