"""pytest configuration: the `gpu` marker, repo paths and the checker (oracle) loaders.

The oracle (oracle/liboracle.so, our C restatement) and, when present, the compiled
unmodified reference (oracle/_ref/libref*.so) are TEST INFRASTRUCTURE: they are only
loaded here, in __graft_entry__.smoke() and in bench.py's CPU-baseline legs.
"""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on the B200 box)")


@pytest.fixture(scope="session", autouse=True)
def _built_oracle():
    """make sure oracle/liboracle.so exists (gcc only; seconds)"""
    if not os.path.exists(os.path.join(ROOT, "oracle", "liboracle.so")):
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "liboracle.so"])


@pytest.fixture(autouse=True)
def _fresh_error_latch(request):
    """uaes_last_error() keeps the most recent failure of the calling thread, and several tests provoke
    failures on purpose: every test starts with an empty latch"""
    if "gpu" in request.keywords:
        try:
            import importlib
            importlib.import_module("micro-aes_b200").core().uaes_clear_error()
        except Exception:
            pass
    yield
