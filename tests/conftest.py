import os
import subprocess
import sys

import pytest

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box: pytest -m gpu)")


@pytest.fixture(scope="session")
def oracle_lib():
    """The CPU oracle (test infrastructure).  Built on demand with gcc; never used by the product."""
    from final184_b200 import api as A
    so = os.path.join(REPO, "oracle", "_build", "libf184_oracle.so")
    r = subprocess.run(["make", "-C", os.path.join(REPO, "oracle")], capture_output=True, text=True)
    if r.returncode != 0:        # never fall back to a stale .so: the parity tests would then pass against an out-of-date checker
        pytest.fail("oracle build failed:\n" + r.stdout + r.stderr)
    return A.Library(so, "f184o_", product=False)


@pytest.fixture(scope="session")
def cuda_lib():
    """The product library.  A missing .so or a missing GPU is a hard failure for -m gpu tests, never a skip:
    a GPU test that silently ran elsewhere would void the parity claim."""
    from final184_b200 import api as A
    return A.load_library()


@pytest.fixture(scope="session")
def proc_scene():
    from final184_b200 import scene as S
    return S.procedural_scene(seed=1)


@pytest.fixture(scope="session")
def cams():
    from final184_b200 import scene as S
    return dict(main=S.fixture_constants("main"), shadow=S.fixture_constants("shadow"), voxel=S.fixture_constants("voxel"))
