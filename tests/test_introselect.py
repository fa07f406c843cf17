"""The selection ORDER of cv::KeyPointsFilter::retainBest is defined by libstdc++'s introselect
(/root/reference/src/ORBextractor.cc:586-588, 601-604).  The CUDA kernels run a hand-written
twin (sdslam_b200/csrc/introselect.cuh); here the same header is compiled for the host and compared move for
move against the real std::nth_element, and against the oracle's retainBest.  CPU only."""
import ctypes
import os
import subprocess

import numpy as np
import pytest

from oracle import binding as orc

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def harness(tmp_path_factory):
    so = str(tmp_path_factory.mktemp("twin") / "twin.so")
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-I", os.path.join(ROOT, "sdslam_b200", "csrc"),
                           "-o", so, os.path.join(ROOT, "tests", "helpers", "introselect_harness.cc")])
    lib = ctypes.CDLL(so)
    lib.twin_vs_std.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p]
    lib.exhaustive.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_int]
    lib.exhaustive.restype = ctypes.c_long
    lib.heap_select_calls.restype = ctypes.c_long
    lib.adversary_responses.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_void_p]
    lib.twin_retain_best.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int]
    return lib


def _entries(rng, n, distinct):
    resp = rng.integers(1, distinct + 1, n).astype(np.uint32)
    return (np.arange(n, dtype=np.uint32) << 8) | resp  # unique payload, ascending = emission order


def _check(lib, a, nth):
    t, s = np.empty_like(a), np.empty_like(a)
    first = lib.twin_vs_std(a.ctypes.data_as(ctypes.c_void_p), len(a), nth, t.ctypes.data_as(ctypes.c_void_p),
                            s.ctypes.data_as(ctypes.c_void_p))
    assert first == -1, "twin differs from std::nth_element at %d (n=%d nth=%d)" % (first, len(a), nth)


@pytest.mark.parametrize("distinct", [1, 2, 5, 66, 255])
def test_twin_matches_std_random(harness, distinct):
    rng = np.random.default_rng(distinct)
    for n in list(range(1, 40)) + [64, 100, 257, 1000, 4096, 20000]:
        a = _entries(rng, n, distinct)
        for nth in sorted({0, n // 3, n // 2, max(n - 2, 0), n - 1}):
            _check(harness, a, nth)


def test_twin_matches_std_adversarial(harness):
    """Sorted, reverse-sorted, organ-pipe and median-of-3 killer inputs drive introselect into its heap-select
    fallback (depth limit 2*floor(log2 n))."""
    for n in (16, 100, 1000, 5000):
        base = np.arange(n, dtype=np.uint32)
        pats = [base % 200 + 1, (n - base) % 200 + 1, np.minimum(base, n - base) % 250 + 1]
        killer = np.zeros(n, np.uint32)  # classic median-of-3 killer permutation, folded into 8-bit responses
        k = n // 2
        for i in range(k):
            killer[i] = (i + 1 if i % 2 == 0 else k + i + (1 if i % 2 else 0)) % 250 + 1
            killer[k + i] = (2 * (i + 1)) % 250 + 1
        pats.append(killer)
        for p in pats:
            a = (np.arange(n, dtype=np.uint32) << 8) | p.astype(np.uint32)
            for nth in (0, n // 4, n // 2, n - 1):
                _check(harness, a, nth)


def test_twin_heap_select_fallback(harness):
    """Inputs built by McIlroy's adversary exhaust introselect's depth limit, so the heap-select branch of the twin
    (never reached by random data) is compared against libstdc++ too -- and is proven to have run."""
    hits = 0
    for n in (64, 300, 1000, 4000):
        for nth in (n // 2, n // 5, 3):
            resp = np.zeros(n, np.uint32)
            harness.adversary_responses(n, nth, resp.ctypes.data_as(ctypes.c_void_p))
            a = (np.arange(n, dtype=np.uint32) << 8) | resp
            before = harness.heap_select_calls()
            _check(harness, a, nth)
            hits += harness.heap_select_calls() - before
    assert hits > 0, "the adversarial inputs never reached heap_select"


@pytest.mark.parametrize("n,distinct", [(9, 4), (12, 3), (15, 2), (17, 2), (14, 3)])
def test_twin_matches_std_exhaustive(harness, n, distinct):
    """All response strings of a given length: covers every branch of the median-of-3 / partition / insertion
    sort logic on tie-heavy inputs."""
    nths = np.array(sorted({0, 1, n // 2, n - 2, n - 1}), np.int32)
    assert harness.exhaustive(n, distinct, nths.ctypes.data_as(ctypes.c_void_p), len(nths)) == 0


def test_twin_retain_best_equals_oracle(harness):
    """retainBest + resize as the kernels do it (first n of the twin's nth_element) == the oracle's
    KeyPointsFilter::retainBest (real std::nth_element + std::partition) truncated to n."""
    rng = np.random.default_rng(9)
    for n in (5, 31, 200, 1500):
        for keep in (1, 2, n // 3, n - 1):
            a = _entries(rng, n, 40)
            order = orc.retain_best_order((a & 0xFF).astype(np.float32), keep)[:keep]
            b = a.copy()
            m = harness.twin_retain_best(b.ctypes.data_as(ctypes.c_void_p), n, keep)
            assert m == keep and np.array_equal(b[:keep] >> 8, order.astype(np.uint32))
