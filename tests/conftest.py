import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def lib_built():
    from sonata_b200 import _native, build
    if not os.path.exists(_native.LIB_PATH):
        build.build()
    return _native.lib()


@pytest.fixture(scope="session")
def voice_paths(lib_built):
    from sonata_b200 import voicegen
    d = voicegen.default_voice_dir()
    return {q: voicegen.write_voice(d, q) for q in ("medium", "high")}


@pytest.fixture(scope="session")
def oracle_weights():
    from oracle import vits_oracle as vo
    from sonata_b200 import voicegen
    cache = {}

    def get(q):
        if q not in cache:
            cache[q] = vo.to_torch(voicegen.make_tensors(q))
        return cache[q]
    return get
