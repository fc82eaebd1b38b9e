"""The linear-sum-assignment core of the GPU Hungarian solver (dpft_b200/csrc/lsap_core.h) built for the HOST (tools/lsap_host.cpp,
LANES = 1: the same source the warp kernel compiles with LANES = 32) against scipy.optimize.linear_sum_assignment, which is
what the reference calls (src/dprt/training/assigner.py:136); plus the tensor-only re-ordering of the kernel's output into
scipy's (row_ind, col_ind) convention."""
import ctypes
import os
import subprocess

import numpy as np
import pytest
import torch
from scipy.optimize import linear_sum_assignment

from conftest import ROOT
from dpft_b200 import criterion


@pytest.fixture(scope="module")
def lsap_host(tmp_path_factory):
    so = str(tmp_path_factory.mktemp("lsap") / "liblsap_host.so")
    subprocess.run(["g++", "-O2", "-shared", "-fPIC", "-o", so, os.path.join(ROOT, "tools", "lsap_host.cpp")], check=True)
    lib = ctypes.CDLL(so)
    lib.lsap_host.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_void_p]
    lib.lsap_host.restype = ctypes.c_int

    lib.lsap_host_lanes.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_void_p]
    lib.lsap_host_lanes.restype = ctypes.c_int

    def solve(cost, m, lanes=None):          # cost (N, Mmax) float32, first m columns valid -> prediction per target
        cost = np.ascontiguousarray(cost, dtype=np.float32)
        out = np.full(cost.shape[1], -7, dtype=np.int64)
        if lanes is None:
            st = lib.lsap_host(cost.ctypes.data, cost.shape[1], m, cost.shape[0], out.ctypes.data)
        else:
            st = lib.lsap_host_lanes(cost.ctypes.data, cost.shape[1], m, cost.shape[0], lanes, out.ctypes.data)
        return st, out
    return solve


def test_host_build_of_the_core_matches_scipy(lsap_host):
    rng = np.random.default_rng(0)
    for t in range(300):
        N = int(rng.integers(1, 450))
        M = int(rng.integers(1, min(N, 64) + 1))
        Mmax = min(64, M + int(rng.integers(0, 4)))
        c = (rng.standard_normal((N, Mmax)) * float(rng.choice([1, 10, 0.01]))).astype(np.float32)
        ties = t % 5 == 0
        if ties:
            c = np.round(c)
        st, out = lsap_host(c, M)
        assert st == 0 and len(set(out[:M].tolist())) == M and out[:M].min() >= 0 and out[:M].max() < N
        rows, cols = linear_sum_assignment(c[:, :M].astype(np.float64))
        mine, want = c[out[:M], np.arange(M)].astype(np.float64).sum(), c[rows, cols].astype(np.float64).sum()
        assert abs(mine - want) <= 1e-9 * max(1.0, abs(want)), (t, mine, want)          # optimal
        if not ties:                                                                     # unique optimum: the same matching
            match = dict(zip(cols.tolist(), rows.tolist()))
            assert all(match[i] == int(out[i]) for i in range(M)), t


def test_core_reports_problems_without_a_finite_assignment(lsap_host):
    c = np.full((3, 2), 1e30, dtype=np.float32) * 1e30                                  # inf everywhere
    st, out = lsap_host(c, 2)
    assert st == -1 and (out[:2] == -1).all()


def test_output_reordering_equals_scipy_convention(lsap_host):
    rng = np.random.default_rng(1)
    B, N, Mmax = 5, 40, 9
    counts = [3, 0, 9, 1, 6]
    cost = rng.standard_normal((B, N, Mmax)).astype(np.float32)
    col4row = np.full((B, Mmax), -1, dtype=np.int64)
    for b, m in enumerate(counts):
        if m:
            col4row[b, :m] = lsap_host(cost[b], m)[1][:m]
    i, j, valid = criterion.order_like_scipy(torch.from_numpy(col4row), torch.tensor(counts, dtype=torch.int32), N)
    for b, m in enumerate(counts):
        assert int(valid[b].sum()) == m and bool(valid[b, :m].all())
        if m:
            rows, cols = linear_sum_assignment(cost[b, :, :m].astype(np.float64))
            assert i[b, :m].tolist() == rows.tolist() and j[b, :m].tolist() == cols.tolist()
        assert i[b, m:].abs().sum() == 0 and j[b, m:].abs().sum() == 0


def test_warp_decomposition_emulated_on_the_host_matches_the_sequential_build(lsap_host):
    """The kernel's 32-lane decomposition (lane-strided scans + xor-butterfly reduction with the shared ordering rule),
    emulated lane by lane on the host: all lanes agree on every winner and the matching equals the sequential build's —
    including cost matrices full of ties, where the tie rule decides."""
    rng = np.random.default_rng(2)
    for t in range(200):
        N = int(rng.integers(1, 450))
        M = int(rng.integers(1, min(N, 64) + 1))
        c = rng.standard_normal((N, M)).astype(np.float32)
        if t % 2 == 0:
            c = np.round(c * 2)
        st1, seq = lsap_host(c, M)
        for lanes in (32, 4):
            st2, par = lsap_host(c, M, lanes=lanes)
            assert st1 == 0 and st2 == 0
            assert (seq[:M] == par[:M]).all(), (t, lanes)
