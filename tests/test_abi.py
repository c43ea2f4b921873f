"""CPU tests: the C-ABI library loads and exports every symbol include/upsp_gpu.h declares;
without a GPU every compute entry point fails loudly (no CPU fallback)."""
import ctypes
import os
import re

import pytest

from conftest import ROOT


def _declared_symbols():
    hdr = open(os.path.join(ROOT, "include", "upsp_gpu.h")).read()
    return sorted(set(re.findall(r"UPSP_API\s+[\w\s\*]+?\b(upsp_\w+)\s*\(", hdr)))


def test_header_declares_expected_entry_points():
    syms = _declared_symbols()
    for s in ("upsp_gpu_create", "upsp_gpu_process_frames", "upsp_gpu_transpose", "upsp_gpu_phase2",
              "upsp_gpu_push_frames", "upsp_gpu_finish_phase1", "upsp_op_warp_affine"):
        assert s in syms
    assert len(syms) >= 35


def test_library_exports_every_declared_symbol(up):
    lib = ctypes.CDLL(up.LIB_PATH)
    missing = [s for s in _declared_symbols() if not hasattr(lib, s)]
    assert not missing, missing


def test_library_has_sm100a_code_only(up):
    import subprocess
    out = subprocess.run(["cuobjdump", "-lelf", up.LIB_PATH], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_\d+a?", out))
    assert archs == {"sm_100a"}, archs


def test_no_cpu_fallback_without_gpu(up):
    if up.device_count() > 0:
        pytest.skip("a GPU is visible")
    with pytest.raises(up.UpspGpuError) as e:
        up.PspGpu(1, 10, 10)
    assert "no CPU fallback" in str(e.value)
    import numpy as np
    with pytest.raises(up.UpspGpuError):
        up.op_transpose(np.zeros((4, 4), np.float32))


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "upsp-processing_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".hpp", ".h", ".cpp")):
                txt = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "oracle" not in txt.lower() or f == "__init__.py" and "oracle" not in txt.lower(), \
                    f"{f} mentions the oracle"
