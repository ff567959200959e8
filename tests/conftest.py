import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests", "emul"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def _has_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    if _has_gpu():
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for it in items:
        if "gpu" in it.keywords:
            it.add_marker(skip)


@pytest.fixture(scope="session")
def po():
    from oracle import pyoracle
    pyoracle.build()
    return pyoracle


@pytest.fixture(scope="session")
def synth():
    from rmcl_b200 import synth as s
    return s


_MESHES = {}


def mesh(name):
    """Session cache of synthetic meshes: 'cube29', 'building:N', 'indoor:N', 'uvsphere:A:B'."""
    from rmcl_b200 import synth as s
    if name not in _MESHES:
        if name.startswith("cube"):
            _MESHES[name] = s.cube(int(name[4:]))
        elif name.startswith("building:"):
            _MESHES[name] = s.building(int(name.split(":")[1]))
        elif name.startswith("indoor:"):
            _MESHES[name] = s.indoor(int(name.split(":")[1]))
        elif name.startswith("uvsphere:"):
            _, a, b = name.split(":")
            _MESHES[name] = s.uvsphere(int(a), int(b))
        else:
            raise KeyError(name)
    return _MESHES[name]


_ORACLE_SCENES = {}


def oracle_scene(name):
    from oracle import pyoracle
    if name not in _ORACLE_SCENES:
        V, F = mesh(name)
        _ORACLE_SCENES[name] = pyoracle.Scene(V, F)
    return _ORACLE_SCENES[name]


_GPU_MAPS = {}


def gpu_map(name):
    import rmcl_b200
    if name not in _GPU_MAPS:
        V, F = mesh(name)
        _GPU_MAPS[name] = rmcl_b200.Map(V, F, device=0)
    return _GPU_MAPS[name]


def quat_close(a, b, tol):
    """quaternions equal up to sign"""
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return min(np.abs(a - b).max(), np.abs(a + b).max()) <= tol


def random_rays(n, lo, hi, seed=0):
    rng = np.random.default_rng(seed)
    o = rng.uniform(lo, hi, (n, 3)).astype(np.float32)
    d = rng.normal(size=(n, 3))
    d = (d / np.linalg.norm(d, axis=1, keepdims=True)).astype(np.float32)
    return o, d


@pytest.fixture(autouse=True)
def _no_pending_cuda_error(request):
    """GPU hygiene: no entry point of the library may leave an unconsumed CUDA runtime error behind (the next caller's
    cudaGetLastError() would trip over it) -- neither during a test nor when its handles are garbage-collected afterwards."""
    gpu = "gpu" in request.keywords and _has_gpu()
    if gpu:
        import gc
        import rmcl_b200
        gc.collect()
        msg = rmcl_b200.load_library().b2_peek_cuda_error().decode()
        assert msg == "", f"pending CUDA error before the test (left by the destruction of earlier handles): {msg}"
    yield
    if gpu:
        msg = rmcl_b200.load_library().b2_peek_cuda_error().decode()
        assert msg == "", f"pending CUDA error after the test: {msg}"
