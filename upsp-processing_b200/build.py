"""Build recipe for libupsp_gpu.so (sm_100a only, in-tree).

    python upsp-processing_b200/build.py            # build if stale
    python upsp-processing_b200/build.py --force
"""
from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libupsp_gpu.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-std=c++17", "-lineinfo", "-Xptxas=-v",
    "-Xcompiler", "-fPIC,-fvisibility=hidden,-ffp-contract=off,-O2",
]
LINK_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-cudart", "static"]
# translation units of the library (each compiled to its own object, in parallel)
UNITS = ["upsp_gpu.cu", "proj_tma.cu"]
OBJ_DIR = os.path.join(HERE, "build")


def sources():
    return [os.path.join(CSRC, f) for f in sorted(os.listdir(CSRC)) if f.endswith((".cu", ".cuh", ".hpp", ".inc", ".h"))] + [
        os.path.join(HERE, "..", "include", "upsp_gpu.h")]


def stale() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(s) > t for s in sources())


def _deps(unit: str):
    """Headers a translation unit depends on (coarse: every header of csrc/ + the public one)."""
    hdrs = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".hpp", ".inc", ".h"))]
    return [os.path.join(CSRC, unit), os.path.join(HERE, "..", "include", "upsp_gpu.h")] + hdrs


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not stale():
        return LIB
    os.makedirs(OBJ_DIR, exist_ok=True)
    ccbin = ["-ccbin", "/usr/bin/g++"] if os.path.exists("/usr/bin/g++") else []   # the image's $CC wrapper lacks libgomp specs
    procs, objs, log = [], [], []
    for u in UNITS:
        obj = os.path.join(OBJ_DIR, u.replace(".cu", ".o"))
        objs.append(obj)
        if not force and os.path.exists(obj) and os.path.getmtime(obj) >= max(map(os.path.getmtime, _deps(u))):
            continue
        cmd = [NVCC] + NVCC_FLAGS + ccbin + ["-c", "-o", obj, os.path.join(CSRC, u)]
        procs.append((cmd, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    failed = False
    for cmd, pr in procs:
        out, _ = pr.communicate()
        log.append(" ".join(cmd) + "\n" + out)
        if verbose or pr.returncode:
            sys.stderr.write(out)
        failed = failed or pr.returncode != 0
    if failed:
        raise RuntimeError("nvcc failed building libupsp_gpu.so")
    cmd = [NVCC] + LINK_FLAGS + ccbin + ["-o", LIB] + objs
    r = subprocess.run(cmd, capture_output=True, text=True)
    log.append(" ".join(cmd) + "\n" + r.stdout + r.stderr)
    if verbose or r.returncode:
        sys.stderr.write(r.stdout + r.stderr)
    if r.returncode:
        raise RuntimeError("nvcc failed linking libupsp_gpu.so")
    with open(os.path.join(HERE, "build.log"), "w") as f:
        f.write("\n".join(log))
    return LIB


HOST_BIN = os.path.join(HERE, "psp_process_b200")


def build_host(force: bool = False) -> str:
    """The C++ host driver (g++, links libupsp_gpu.so through its C ABI only)."""
    src = os.path.join(HERE, "host", "psp_process_b200.cpp")
    hdrs = [os.path.join(HERE, "host", h) for h in os.listdir(os.path.join(HERE, "host")) if h.endswith((".hpp", ".inc"))]
    build()
    if not force and os.path.exists(HOST_BIN) and os.path.getmtime(HOST_BIN) >= max(
            [os.path.getmtime(src), os.path.getmtime(LIB)] + list(map(os.path.getmtime, hdrs))):
        return HOST_BIN
    gxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
    cmd = [gxx, "-O2", "-std=c++17", "-ffp-contract=off", "-Wall", "-Wextra", "-o", HOST_BIN, src, "-L" + HERE, "-lupsp_gpu",
           "-Wl,-rpath,$ORIGIN", "-ldl", "-lpthread", "-lrt"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode:
        sys.stderr.write(r.stdout + r.stderr)
        raise RuntimeError("g++ failed building psp_process_b200")
    return HOST_BIN


XPOSE_BIN = os.path.join(HERE, "upsp_matrix_transpose_b200")


def build_transpose_tool(force: bool = False) -> str:
    """host/upsp_matrix_transpose_b200.cpp: the reference's stand-alone transpose tool on the C ABI."""
    src = os.path.join(HERE, "host", "upsp_matrix_transpose_b200.cpp")
    build()
    if not force and os.path.exists(XPOSE_BIN) and os.path.getmtime(XPOSE_BIN) >= max(os.path.getmtime(src), os.path.getmtime(LIB)):
        return XPOSE_BIN
    gxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
    cmd = [gxx, "-O2", "-std=c++17", "-Wall", "-Wextra", "-o", XPOSE_BIN, src, "-L" + HERE, "-lupsp_gpu",
           "-Wl,-rpath,$ORIGIN", "-ldl", "-lpthread", "-lrt"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode:
        sys.stderr.write(r.stdout + r.stderr)
        raise RuntimeError("g++ failed building upsp_matrix_transpose_b200")
    return XPOSE_BIN


H5_PROBE_BIN = os.path.join(HERE, "psp_hdf5_probe")


def build_h5_probe(force: bool = False) -> str:
    """host/psp_hdf5_probe.cpp: writes a sample PSP HDF5 file with host/psp_hdf5.hpp (no library dependency at all)."""
    src = os.path.join(HERE, "host", "psp_hdf5_probe.cpp")
    hdr = os.path.join(HERE, "host", "psp_hdf5.hpp")
    if not force and os.path.exists(H5_PROBE_BIN) and os.path.getmtime(H5_PROBE_BIN) >= max(os.path.getmtime(src), os.path.getmtime(hdr)):
        return H5_PROBE_BIN
    gxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
    r = subprocess.run([gxx, "-O2", "-std=c++17", "-Wall", "-Wextra", "-o", H5_PROBE_BIN, src], capture_output=True, text=True)
    if r.returncode:
        sys.stderr.write(r.stdout + r.stderr)
        raise RuntimeError("g++ failed building psp_hdf5_probe")
    return H5_PROBE_BIN


ADD_FIELD_BIN = os.path.join(HERE, "add_field_b200")


def build_add_field_tool(force: bool = False) -> str:
    """host/add_field_b200.cpp: the reference's `add_field` (flat [rows x extent] float file -> chunked HDF5 dataset) on
    host/h5_append.hpp, no HDF5 library."""
    src = os.path.join(HERE, "host", "add_field_b200.cpp")
    deps = [src, os.path.join(HERE, "host", "h5_append.hpp")]
    if not force and os.path.exists(ADD_FIELD_BIN) and os.path.getmtime(ADD_FIELD_BIN) >= max(map(os.path.getmtime, deps)):
        return ADD_FIELD_BIN
    gxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
    r = subprocess.run([gxx, "-O2", "-std=c++17", "-Wall", "-Wextra", "-o", ADD_FIELD_BIN, src], capture_output=True, text=True)
    if r.returncode:
        sys.stderr.write(r.stdout + r.stderr)
        raise RuntimeError("g++ failed building add_field_b200")
    return ADD_FIELD_BIN


PATCH_PROBE_BIN = os.path.join(HERE, "patch_geometry_probe")


def build_patch_probe(force: bool = False) -> str:
    """host/patch_geometry_probe.cpp: phase-0 patch geometry (host C++) exercised from the tests."""
    src = os.path.join(HERE, "host", "patch_geometry_probe.cpp")
    deps = [src, os.path.join(HERE, "host", "patch_geometry.hpp")]
    if not force and os.path.exists(PATCH_PROBE_BIN) and os.path.getmtime(PATCH_PROBE_BIN) >= max(map(os.path.getmtime, deps)):
        return PATCH_PROBE_BIN
    gxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
    r = subprocess.run([gxx, "-O2", "-std=c++17", "-Wall", "-Wextra", "-o", PATCH_PROBE_BIN, src], capture_output=True, text=True)
    if r.returncode:
        sys.stderr.write(r.stdout + r.stderr)
        raise RuntimeError("g++ failed building patch_geometry_probe")
    return PATCH_PROBE_BIN


GRID_PROBE_BIN = os.path.join(HERE, "grid_probe")


def build_grid_probe(force: bool = False) -> str:
    """host/grid_probe.cpp: .tri grid reader + node normals (host C++) exercised from the tests."""
    src = os.path.join(HERE, "host", "grid_probe.cpp")
    deps = [src, os.path.join(HERE, "host", "grid_readers.hpp")]
    if not force and os.path.exists(GRID_PROBE_BIN) and os.path.getmtime(GRID_PROBE_BIN) >= max(map(os.path.getmtime, deps)):
        return GRID_PROBE_BIN
    gxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
    r = subprocess.run([gxx, "-O2", "-std=c++17", "-ffp-contract=off", "-Wall", "-Wextra", "-o", GRID_PROBE_BIN, src],
                       capture_output=True, text=True)
    if r.returncode:
        sys.stderr.write(r.stdout + r.stderr)
        raise RuntimeError("g++ failed building grid_probe")
    return GRID_PROBE_BIN


INPUTS_PROBE_BIN = os.path.join(HERE, "inputs_probe")


def build_inputs_probe(force: bool = False) -> str:
    """host/inputs_probe.cpp: deck / paint calibration / wtd / targets / p3d function / histogram readers and
    the structured model's seam detection (host C++) exercised from the tests."""
    src = os.path.join(HERE, "host", "inputs_probe.cpp")
    deps = [src] + [os.path.join(HERE, "host", h) for h in ("upsp_inputs.hpp", "run_inputs.hpp", "p3d_model.hpp", "grid_readers.hpp", "interpolation.hpp")]
    if not force and os.path.exists(INPUTS_PROBE_BIN) and os.path.getmtime(INPUTS_PROBE_BIN) >= max(map(os.path.getmtime, deps)):
        return INPUTS_PROBE_BIN
    gxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
    r = subprocess.run([gxx, "-O2", "-std=c++17", "-ffp-contract=off", "-Wall", "-Wextra", "-o", INPUTS_PROBE_BIN, src],
                       capture_output=True, text=True)
    if r.returncode:
        sys.stderr.write(r.stdout + r.stderr)
        raise RuntimeError("g++ failed building inputs_probe")
    return INPUTS_PROBE_BIN


TBL_BIN = os.path.join(HERE, "xyz_scalar_to_tbl_b200")


def build_tbl_tool(force: bool = False) -> str:
    """host/xyz_scalar_to_tbl_b200.cpp: flat files -> Tecplot table (the reference's xyz_scalar_to_tbl / _delta tools)."""
    src = os.path.join(HERE, "host", "xyz_scalar_to_tbl_b200.cpp")
    if not force and os.path.exists(TBL_BIN) and os.path.getmtime(TBL_BIN) >= os.path.getmtime(src):
        return TBL_BIN
    gxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
    r = subprocess.run([gxx, "-O2", "-std=c++17", "-Wall", "-Wextra", "-o", TBL_BIN, src], capture_output=True, text=True)
    if r.returncode:
        sys.stderr.write(r.stdout + r.stderr)
        raise RuntimeError("g++ failed building xyz_scalar_to_tbl_b200")
    return TBL_BIN


SETUP_BIN = os.path.join(HERE, "psp_setup_b200")


def build_setup_tool(force: bool = False) -> str:
    """host/psp_setup_b200.cpp: grid + calibration -> projection matrix (phase 0 on the GPU)."""
    src = os.path.join(HERE, "host", "psp_setup_b200.cpp")
    deps = [src] + [os.path.join(HERE, "host", h) for h in (
        "camera_cal.hpp", "grid_readers.hpp", "p3d_model.hpp", "projection_weights.hpp", "run_inputs.hpp", "upsp_inputs.hpp",
        "video_readers.hpp", "targets.hpp", "patch_geometry.hpp", "interpolation.hpp", "deck_job.hpp")]
    build()
    if not force and os.path.exists(SETUP_BIN) and os.path.getmtime(SETUP_BIN) >= max([os.path.getmtime(LIB)] + list(map(os.path.getmtime, deps))):
        return SETUP_BIN
    gxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
    cmd = [gxx, "-O2", "-std=c++17", "-ffp-contract=off", "-Wall", "-Wextra", "-o", SETUP_BIN, src, "-L" + HERE, "-lupsp_gpu",
           "-Wl,-rpath,$ORIGIN", "-ldl", "-lpthread", "-lrt"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode:
        sys.stderr.write(r.stdout + r.stderr)
        raise RuntimeError("g++ failed building psp_setup_b200")
    return SETUP_BIN


WEIGHTS_PROBE_BIN = os.path.join(HERE, "projection_weights_probe")


def build_weights_probe(force: bool = False) -> str:
    """host/projection_weights_probe.cpp: multi-camera blending weights (host C++) exercised from the tests."""
    src = os.path.join(HERE, "host", "projection_weights_probe.cpp")
    deps = [src, os.path.join(HERE, "host", "projection_weights.hpp")]
    if not force and os.path.exists(WEIGHTS_PROBE_BIN) and os.path.getmtime(WEIGHTS_PROBE_BIN) >= max(map(os.path.getmtime, deps)):
        return WEIGHTS_PROBE_BIN
    gxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
    r = subprocess.run([gxx, "-O2", "-std=c++17", "-ffp-contract=off", "-Wall", "-Wextra", "-o", WEIGHTS_PROBE_BIN, src],
                       capture_output=True, text=True)
    if r.returncode:
        sys.stderr.write(r.stdout + r.stderr)
        raise RuntimeError("g++ failed building projection_weights_probe")
    return WEIGHTS_PROBE_BIN


PROBE_BIN = os.path.join(HERE, "video_probe")


def build_probe(force: bool = False) -> str:
    """host/video_probe.cpp: container readers (cine / mraw) exercised from the tests; plain g++."""
    src = os.path.join(HERE, "host", "video_probe.cpp")
    deps = [src, os.path.join(HERE, "host", "video_readers.hpp"), os.path.join(HERE, "host", "cine_lut.inc")]
    if not force and os.path.exists(PROBE_BIN) and os.path.getmtime(PROBE_BIN) >= max(map(os.path.getmtime, deps)):
        return PROBE_BIN
    gxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
    r = subprocess.run([gxx, "-O2", "-std=c++17", "-Wall", "-Wextra", "-o", PROBE_BIN, src], capture_output=True, text=True)
    if r.returncode:
        sys.stderr.write(r.stdout + r.stderr)
        raise RuntimeError("g++ failed building video_probe")
    return PROBE_BIN


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
    print(build_host(force="--force" in sys.argv))
    print(build_probe(force="--force" in sys.argv))
    print(build_transpose_tool(force="--force" in sys.argv))
    print(build_patch_probe(force="--force" in sys.argv))
    print(build_grid_probe(force="--force" in sys.argv))
    print(build_setup_tool(force="--force" in sys.argv))
    print(build_weights_probe(force="--force" in sys.argv))
    print(build_inputs_probe(force="--force" in sys.argv))
    print(build_tbl_tool(force="--force" in sys.argv))
