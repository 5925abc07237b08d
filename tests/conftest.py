import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def _has_gpu():
    try:
        import genlib_b200 as gen
        return gen.lib().genlib_device_count() > 0
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    # `-m gpu` on a machine without a GPU must fail loudly rather than skip silently;
    # plain `pytest` (no -m) on a CPU box skips the GPU tests.
    if config.getoption("-m"):
        return
    if not _has_gpu():
        skip = pytest.mark.skip(reason="no CUDA device")
        for it in items:
            if "gpu" in it.keywords:
                it.add_marker(skip)


@pytest.fixture(scope="session")
def gen():
    import genlib_b200
    return genlib_b200


@pytest.fixture(scope="session")
def ob():
    from oracle import binding
    binding.build()
    return binding


@pytest.fixture(scope="session")
def genea140_oracle(gen, ob):
    """(oracle pedigree, 140x140 float32 oracle matrix, per-step info) -- computed once (~6 s)."""
    p = ob.OraclePedigree.from_csv(gen.genea140)
    phi, steps = p.phi(with_steps=True)
    return p, phi, steps
