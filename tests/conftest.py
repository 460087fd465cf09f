import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run on the GPU box with -m gpu)")
    import __graft_entry__ as ge

    ge.build()  # no-op when the in-tree libraries are newer than their sources


@pytest.fixture(scope="session")
def oracle():
    from oracle_backend import oracle_lib

    return oracle_lib()


@pytest.fixture(scope="session")
def renderer():
    """One GPU context for the session. Fails loudly (no skip, no fallback) when there is no B200."""
    from mass_raytrace_b200 import Renderer

    r = Renderer(0)
    yield r
    r.close()


@pytest.fixture(scope="session")
def tmp_mesh_dir(tmp_path_factory):
    return tmp_path_factory.mktemp("meshes")
