"""CPU-side checks of the C-ABI boundary: libpm.so builds, loads and exports every symbol
include/pm.h declares; pm_dmatch is layout-identical to cv::DMatch; without a GPU the product
fails loudly (no CPU fallback)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def L():
    import points_matching_b200 as pm
    pm.build()
    from points_matching_b200 import _lib
    return _lib.lib()


def _header_symbols():
    src = open(os.path.join(ROOT, "include", "pm.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(pm_[a-z0-9_]+)\s*\(", src)))


def test_every_header_symbol_is_exported(L):
    from points_matching_b200 import _lib
    syms = _header_symbols()
    assert len(syms) >= 30
    for s in syms:
        assert hasattr(L, s), f"{s} declared in include/pm.h but not exported by libpm.so"
    assert sorted(_lib.EXPORTS) == syms


def test_dmatch_layout_matches_cv_dmatch():
    from points_matching_b200 import DMATCH
    assert DMATCH.itemsize == 16
    assert [DMATCH.fields[n][1] for n in ("queryIdx", "trainIdx", "imgIdx", "distance")] == [0, 4, 8, 12]


def test_version_and_no_cpu_fallback(L):
    import torch
    assert L.pm_version() == 200
    if torch.cuda.is_available():
        pytest.skip("GPU present: covered by the -m gpu tests")
    h = C.c_void_p()
    assert L.pm_create(C.byref(h), 0) == -4          # PM_NO_DEVICE
    import points_matching_b200 as pm
    with pytest.raises(pm.PMError):
        pm.Context(0)
    with pytest.raises(pm.PMError):
        pm.BFMatcher(pm.NORM_L2).match(np.zeros((2, 128), np.float32), np.zeros((2, 128), np.float32))


def test_sample_sets_host_generator(L):
    from points_matching_b200.api import make_sample_sets
    a = make_sample_sets(1000, 500, 8, seed=5)
    b = make_sample_sets(1000, 500, 8, seed=5)
    assert (a == b).all() and a.min() >= 0 and a.max() < 1000
    s = np.sort(a, axis=1)
    assert (s[:, 1:] != s[:, :-1]).all()             # distinct within a row
    assert (make_sample_sets(1000, 500, 8, seed=6) != a).any()
    c = make_sample_sets(8, 50, 8, seed=1)           # n == m: every row a permutation
    assert (np.sort(c, axis=1) == np.arange(8)).all()
    import points_matching_b200 as pm
    with pytest.raises(pm.PMError):
        make_sample_sets(5, 10, 8)


def test_sass_is_blackwell_native():
    """The L2 kernel must carry tcgen05 / TMA instructions (UTCHMMA, LDTM, UTMALDG)."""
    import shutil
    import subprocess
    if not shutil.which("cuobjdump"):
        pytest.skip("cuobjdump not on PATH")
    from points_matching_b200 import SO_PATH
    sass = subprocess.run(["cuobjdump", "-sass", SO_PATH], capture_output=True, text=True).stdout
    for mnemonic in ("UTCHMMA", "LDTM", "UTMALDG", "POPC"):
        assert mnemonic in sass, mnemonic
    assert "sm_100a" in subprocess.run(["cuobjdump", "-lelf", SO_PATH], capture_output=True, text=True).stdout


def test_cpp_host_layer_compiles_and_links(L):
    """include/pm.hpp (the OpenCV look-alikes the reference's main.cpp would call) and the host
    C++ example build with plain g++ against libpm.so -- no CUDA toolkit on the host side."""
    import subprocess
    subprocess.check_call(["make", "-C", os.path.join(ROOT, "examples"), "-s"])
    exe = os.path.join(ROOT, "examples", "match_and_estimate")
    assert os.path.exists(exe)
    r = subprocess.run([exe], capture_output=True, text=True)
    assert r.returncode == 2 and "usage" in r.stderr


def test_bench_reference_arm_prints_the_contract_line():
    """`bench.py --impl reference` runs the reference's CPU routine (OpenCV batchDistance, else the oracle port) and
    prints ONE JSON line with the contract's keys; without a GPU the product arm refuses to run (no CPU fallback)."""
    import json
    import subprocess
    import sys
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1"],
                         capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "sift_knn2_match_pairs_per_sec" and d["unit"] == "pairs/s"
    assert d["higher_is_better"] is True and d["n_gpus"] == 1 and d["steps"] == 1 and d["value"] > 0
    assert d["cpu_baseline"]["kind"] in ("reference", "port") and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"]["value"] == d["value"] and d["e2e"]["h2d_bytes_per_step"] == 0
    assert "workload" in d["config"]
    import torch
    if not torch.cuda.is_available():
        ours = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "1", "--warmup", "1"],
                              capture_output=True, text=True, timeout=600, cwd=ROOT)
        assert ours.returncode != 0 and "no CUDA device" in (ours.stderr + ours.stdout)
