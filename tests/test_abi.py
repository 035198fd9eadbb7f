"""CPU tests of the boundary: the C-ABI library loads, exports every symbol declared in
include/ppsfm_b200.h, fails loudly without a GPU, and never references the oracle."""
import ctypes
import os
import re
import subprocess

import pytest

import privacy_preserving_sfm_b200 as pp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    src = open(os.path.join(ROOT, "include", "ppsfm_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(ppsfm_[a-z0-9_]+)\s*\(", src)))


def test_library_builds_and_exports_every_declared_symbol():
    pp.build_library()
    lib = ctypes.CDLL(pp.library_path())
    syms = _declared_symbols()
    assert len(syms) >= 15
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in include/ppsfm_b200.h but not exported"


def test_library_is_sm100a_only():
    out = subprocess.run(["cuobjdump", "-lelf", pp.library_path()], capture_output=True,
                         text=True).stdout
    archs = set(re.findall(r"sm_\d+a?", out))
    assert archs == {"sm_100a"}, archs


def test_host_only_entry_points_work_without_gpu():
    L = pp.load_library()
    assert b"sm_100a" in L.ppsfm_version()
    assert pp.ComputeNumTrials(100, 100, 0.99, 3.0) == 1
    o = pp.RANSACOptions()
    L.ppsfm_ransac_options_default(ctypes.byref(o))
    assert (o.min_inlier_ratio, o.confidence, o.dyn_num_trials_multiplier) == (0.1, 0.99, 3.0)
    assert o.min_num_trials == 0 and o.max_num_trials == 2 ** 64 - 1


def test_event_driven_replay_equals_the_literal_loop():
    """RansacResident replays the reference's trial loop (src/optim/ransac.h:213-249) over the
    counts of a wave by jumping between the visits that change state; the library's host-only
    self-test runs it against the literal model-by-model loop on random waves (empty trials,
    ties, carried best models; about half of the waves abort, a third of them mid-wave)."""
    L = pp.load_library()
    L.ppsfm_selftest_replay.argtypes = [ctypes.c_uint32, ctypes.c_int]
    L.ppsfm_selftest_replay.restype = ctypes.c_int
    for seed in (1, 2, 3, 20201017):
        assert L.ppsfm_selftest_replay(seed, 20000) == 0


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(pp.PpsfmError):
        pp.Context(0)


def test_product_never_touches_the_oracle():
    pkg = os.path.join(ROOT, "privacy_preserving_sfm_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cc", ".hpp")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in txt and "liboracle" not in txt, f
    out = subprocess.run(["ldd", pp.library_path()], capture_output=True, text=True).stdout
    assert "oracle" not in out
