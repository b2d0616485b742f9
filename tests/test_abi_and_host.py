"""CPU tests: the C-ABI library loads without a GPU, exports every symbol include/*.h declares, fails loudly
without a device (no CPU fallback), and the product never touches oracle/."""
import ctypes
import os
import re

import numpy as np
import pytest

import orb_slam_b200 as fe

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    names = set()
    for h in ("orbfe.h", "orbfe_match.h", "orbfe_bow.h", "orbfe_comm.h"):
        txt = open(os.path.join(ROOT, "include", h)).read()
        txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
        names |= set(re.findall(r"\b(orbfe_[a-z0-9_]+)\s*\(", txt))
    return names


def test_library_exports_every_declared_symbol():
    L = fe.lib()
    declared = _declared_symbols()
    assert declared == set(fe.ABI_SYMBOLS), declared ^ set(fe.ABI_SYMBOLS)
    for s in sorted(declared):
        assert hasattr(L, s), "liborbfe.so does not export %s" % s
    assert L.orbfe_version() == 1


def test_no_cpu_fallback_without_device():
    L = fe.lib()
    if L.orbfe_device_count() > 0:
        pytest.skip("a CUDA device is present")
    with pytest.raises(fe.OrbfeError) as e:
        fe.ORBextractor(1000, 1.2, 8)
    assert e.value.code == fe.ORBFE_ERR_NO_DEVICE
    with pytest.raises(fe.OrbfeError) as e:
        fe.ORBmatcher(0.9, True)
    assert e.value.code == fe.ORBFE_ERR_NO_DEVICE


def test_argument_validation_happens_before_device_use():
    L = fe.lib()
    h = ctypes.c_void_p()
    assert L.orbfe_extractor_create(1000, 1.2, 0, 1, 20, 0, ctypes.byref(h)) == fe.ORBFE_ERR_ARG      # nlevels < 1
    assert L.orbfe_extractor_create(1000, 1.0, 8, 1, 20, 0, ctypes.byref(h)) == fe.ORBFE_ERR_ARG      # scale <= 1
    assert L.orbfe_extractor_create(1000, 1.2, 8, 2, 20, 0, ctypes.byref(h)) == fe.ORBFE_ERR_ARG      # bad score type
    assert L.orbfe_extractor_create(1000, 1.2, 8, 1, 20, 0, None) == fe.ORBFE_ERR_ARG
    assert b"" != L.orbfe_last_error()
    assert L.orbfe_extractor_destroy(None) == 0 and L.orbfe_matcher_destroy(None) == 0


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "orb_slam_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cpp", ".cc", ".h", ".cuh")):
                txt = open(os.path.join(dp, f), errors="replace").read()
                assert not re.search(r"^\s*(import|from)\s+oracle\b", txt, flags=re.M), f
                assert "orb_oracle" not in txt, f
    # and the shared library has no dependency on the oracle .so
    import subprocess
    out = subprocess.run(["ldd", fe.library_path()], stdout=subprocess.PIPE, text=True).stdout
    assert "orb_oracle" not in out and "libtorch" not in out and "libcuda.so" not in out


def test_keypoint_layout_matches_cv_keypoint():
    assert fe.KP_DTYPE.itemsize == 28
    assert [fe.KP_DTYPE.fields[n][1] for n in ("x", "y", "size", "angle", "response", "octave", "class_id")] == \
        [0, 4, 8, 12, 16, 20, 24]


def test_bench_roofline_bytes_match_survey():
    import bench
    ab = bench.algorithmic_bytes()
    assert ab["P"] == 6419321
    assert abs(ab["total"] / 1e6 - 32.50) < 0.02    # 32.50 MB per 1080p frame
    assert bench.level_sizes()[-1] == (536, 301)


def test_synth_is_deterministic():
    from orb_slam_b200.synth import textured_frame
    a, b = textured_frame(160, 120, seed=3), textured_frame(160, 120, seed=3)
    assert np.array_equal(a, b) and a.dtype == np.uint8 and a.std() > 20


def test_headers_are_plain_c(tmp_path):
    """The boundary is a C ABI: the three headers compile as C99 (-pedantic) and a C program links against liborbfe.so."""
    import subprocess
    src = tmp_path / "abi.c"
    src.write_text('#include "orbfe.h"\n#include "orbfe_match.h"\n#include "orbfe_bow.h"\n#include "orbfe_comm.h"\n'
                   'int main(void) { return (sizeof(OrbfeKeyPoint) == 28 && orbfe_version() == ORBFE_VERSION) ? 0 : 1; }\n')
    exe = tmp_path / "abi.bin"
    so_dir = os.path.dirname(fe.library_path())
    r = subprocess.run(["gcc", "-std=c99", "-pedantic", "-Wall", "-Wextra", "-Werror", "-I", os.path.join(ROOT, "include"), str(src),
                        "-L", so_dir, "-lorbfe", "-Wl,-rpath," + so_dir, "-o", str(exe)],
                       stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    assert r.returncode == 0, r.stdout
    assert subprocess.run([str(exe)]).returncode == 0


def test_stream_driver_builds_and_fails_loudly_without_a_device():
    """bench.py's end-to-end pipeline driver (tools/e2e_driver.cpp) is plain C++ over the public C-ABI: it builds with g++, exports
    its five entry points, links nothing from oracle/, and creating it without a CUDA device fails with the library's message."""
    import ctypes as C
    import subprocess
    from orb_slam_b200.build import build_e2e_driver
    so = build_e2e_driver()
    dl = C.CDLL(so)
    for sym in ("e2e_create", "e2e_run", "e2e_last_matches", "e2e_error", "e2e_destroy"):
        assert hasattr(dl, sym), sym
    needed = subprocess.run(["readelf", "-d", so], stdout=subprocess.PIPE, text=True).stdout
    assert "oracle" not in needed and "liborbfe" not in needed     # entry points of liborbfe.so arrive as function pointers
    src = open(os.path.join(ROOT, "tools", "e2e_driver.cpp")).read()
    assert "#include <cuda" not in src and "cudaMemcpy" not in src and "cudaMalloc" not in src   # the driver makes no CUDA call of its own
    if fe.lib().orbfe_device_count() == 0:
        from orb_slam_b200.stream_driver import StreamDriver
        frames = np.zeros((2, 48, 64), np.uint8)
        bufs = [(np.zeros((1, 50, 28), np.uint8), np.zeros((1, 50, 32), np.uint8), np.zeros(1, np.int32)) for _ in range(3)]
        with pytest.raises(RuntimeError):
            StreamDriver(64, 48, 50, 2, 1.2, 20, 1, 2, 1, 1, 0, 50.0, 50.0, 32.0, 24.0, 3.0, 15.0, frames.ctypes.data, np.zeros((2, 12), np.float32),
                         [b[0].ctypes.data for b in bufs], [b[1].ctypes.data for b in bufs], [b[2].ctypes.data for b in bufs])
