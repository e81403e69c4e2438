import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)

import _pkgload  # noqa: E402

_pkgload.load()


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def cache_dir(tmp_path_factory):
    d = os.environ.get("MSX_CACHE")
    if d:
        os.makedirs(d, exist_ok=True)
        return d
    return str(tmp_path_factory.mktemp("msx_cache"))


@pytest.fixture(scope="session")
def gguf_for(cache_dir):
    from moshi_cpp_b200 import configs, synth

    made = {}

    def make(preset: str, quant: str = "q4_k", seed: int = 1234):
        key = (preset, quant, seed)
        if key not in made:
            path = os.path.join(cache_dir, f"{preset}-{quant}-s{seed}.gguf")
            if not os.path.exists(path):
                synth.write_gguf(path, configs.get(preset), quant, seed)
            made[key] = path
        return made[key], configs.get(preset)

    return make
