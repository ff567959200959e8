"""CPU tests: the C-ABI library loads and exports every symbol include/rmcl_b200.h declares; failure without a GPU is loud."""
import ctypes
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "rmcl_b200.h")).read()
    return sorted(set(re.findall(r"B2_API\s+[\w\s\*]+?\b(b2_\w+)\s*\(", src)))


def test_header_declares_expected_surface():
    names = _declared()
    for must in ("b2_mesh_create", "b2_rcc_find", "b2_rcc_cross_statistics", "b2_rcc_correct_once", "b2_rcc_correct_batch", "b2_umeyama_batch",
                 "b2_pf_sensor_update", "b2_pf_sensor_update_host", "b2_last_error"):
        assert must in names
    assert len(names) >= 35


def test_library_exports_every_declared_symbol():
    import rmcl_b200
    path = rmcl_b200.lib_path()
    assert os.path.exists(path), "librmcl_b200.so not built (python -c 'import __graft_entry__ as g; g.build()')"
    lib = ctypes.CDLL(path)
    for name in _declared():
        assert hasattr(lib, name), name
    assert sorted(rmcl_b200.api.EXPORTS) == _declared()


def test_pod_sizes_match_header():
    import rmcl_b200
    from rmcl_b200 import api, synth
    assert ctypes.sizeof(api.PFParams) == 32 and ctypes.sizeof(api.GladiatorConfig) == 32 and ctypes.sizeof(api._SphericalModel) == 32 and ctypes.sizeof(api._PinholeModel) == 32
    assert synth.TRANSFORM_DTYPE.itemsize == 32 and synth.CROSS_STATS_DTYPE.itemsize == 64


def test_no_cpu_fallback():
    """Without a CUDA device every compute entry point must fail loudly, never compute on the host."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    import rmcl_b200
    with pytest.raises(rmcl_b200.B2Error):
        rmcl_b200.Map(np.zeros((3, 3), np.float32), np.array([[0, 1, 2]], np.uint32))


def test_product_does_not_reference_oracle():
    """The product tree must never import / include / link the oracle."""
    for dirpath, _, files in os.walk(os.path.join(ROOT, "rmcl_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h", "Makefile")):
                txt = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "pyoracle" not in txt and "oracle.h" not in txt and "liboracle" not in txt, os.path.join(dirpath, f)
    for f in os.listdir(os.path.join(ROOT, "include")):
        p = os.path.join(ROOT, "include", f)
        if os.path.isfile(p):
            assert "oracle" not in open(p).read().lower()
