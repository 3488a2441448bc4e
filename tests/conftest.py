import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu under gpurun)")


@pytest.fixture(scope="session")
def golden():
    import torch

    def load(name):
        return torch.load(os.path.join(GOLDEN, name + ".pt"), weights_only=False)
    return load


# Largest errors the parity tests observed against the reference / the oracle, printed at the end of the run so that a
# drift towards a tolerance is visible before it becomes a failure (the assertions themselves are unchanged).
OBSERVED = {}


def record_error(kind, value, scale=None):
    """Keep the largest `value` (and the scale it occurred at) per kind of comparison."""
    value = float(value)
    old = OBSERVED.get(kind)
    if old is None or value > old[0]:
        OBSERVED[kind] = (value, None if scale is None else float(scale))


def pytest_terminal_summary(terminalreporter):
    if not OBSERVED:
        return
    terminalreporter.write_sep("-", "largest observed parity errors")
    for kind in sorted(OBSERVED):
        v, sc = OBSERVED[kind]
        terminalreporter.write_line(f"{kind:58s} {v:.3e}" + ("" if sc is None else f"   (at scale {sc:.3e})"))
