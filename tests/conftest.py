import importlib
import pathlib
import sys

import pytest

ROOT = pathlib.Path(__file__).resolve().parent.parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def ora():
    import oracle

    return oracle.Oracle()


@pytest.fixture(scope="session")
def ref():
    """The unmodified reference compiled into oracle/_ref (skips where it was never built)."""
    import oracle

    try:
        return oracle.Reference("serial")
    except oracle.ReferenceUnavailable as e:
        pytest.skip(str(e))


@pytest.fixture(scope="session")
def golden_dir():
    import oracle

    if not (oracle.GOLDEN_DIR / "chef-with-trumpet.bmp").exists():
        pytest.skip("reference golden images not staged (oracle/_ref/golden)")
    return oracle.GOLDEN_DIR


@pytest.fixture(scope="session")
def pkg():
    return importlib.import_module("yuv-manipulations-2_b200")


@pytest.fixture(scope="session")
def synth():
    return importlib.import_module("yuv-manipulations-2_b200.synth")


@pytest.fixture(scope="session")
def ctx(pkg):
    c = pkg.Context(0)
    yield c
    c.close()
