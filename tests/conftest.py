import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))
if str(ROOT / "tests") not in sys.path:
    sys.path.insert(0, str(ROOT / "tests"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200, sm_100a); run with -m gpu")


@pytest.fixture(scope="session")
def golden_dir():
    return ROOT / "tests" / "golden"


@pytest.fixture(autouse=True)
def _library_options_are_per_test(request):
    """The perf / A-B switches of the library (ks_set_option) are process-global: every GPU test starts and ends with the defaults,
    so a test that toggles a kernel variant cannot leak it into the next one."""
    if request.node.get_closest_marker("gpu") is None:
        yield
        return
    from kurosiwo_b200.lib import default_ops
    ops = default_ops()
    ops.reset_options()
    yield
    ops.reset_options()
