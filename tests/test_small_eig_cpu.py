"""The per-thread eigenvalue routine of the Einstein kernel (csrc/small_eig.cuh) is plain C++: build it for the host
and check it against LAPACK (np.linalg.eigvals, what the reference calls per grid point, einstein.py:293-298)."""
import ctypes as C
import os
import shutil
import subprocess

import numpy as np
import pytest

from conftest import ROOT, load_golden
from oracle import oracle_np
from test_oracle_golden import einstein_inputs


@pytest.fixture(scope="module")
def lib(tmp_path_factory):
    if shutil.which("g++") is None:
        pytest.skip("g++ not available")
    so = str(tmp_path_factory.mktemp("small_eig") / "libsmall_eig.so")
    src = os.path.join(ROOT, "tests", "helpers", "small_eig_host.cpp")
    subprocess.run(["g++", "-O2", "-ffp-contract=off", "-shared", "-fPIC", "-o", so, src], check=True)
    return C.CDLL(so)


def _eig(lib, mats):
    n = mats.shape[-1]
    m = np.ascontiguousarray(mats, dtype=np.float64)
    cnt = m.shape[0]
    wr, wi, ok = np.zeros((cnt, n)), np.zeros((cnt, n)), np.zeros(cnt, dtype=np.int32)
    ptr = lambda a: a.ctypes.data_as(C.c_void_p)
    getattr(lib, f"small_eig{n}")(ptr(m), cnt, ptr(wr), ptr(wi), ptr(ok))
    return wr + 1j * wi, ok


@pytest.mark.parametrize("n", [3, 4])
def test_random_matrices_match_lapack(lib, n):
    rng = np.random.default_rng(7)
    a = rng.standard_normal((5000, n, n)) * 10.0 ** rng.integers(-3, 4, size=(5000, 1, 1))
    a[:200] *= np.array([1e-6, 1.0, 1e6, 1.0][:n])[None, :, None]          # badly scaled rows: balancing matters
    w, ok = _eig(lib, a)
    ref = np.linalg.eigvals(a)
    assert ok.all()
    err = np.abs(np.sort_complex(w) - np.sort_complex(ref)).max(axis=1) / np.abs(ref).max(axis=1)
    assert err.max() < 1e-11


def test_special_matrices(lib):
    mats = np.array([
        np.zeros((3, 3)),                                            # flat space: G = 0
        np.diag([1.0, 1.0, 2.0]),                                    # repeated, diagonalisable
        [[0.0, 1.0, 0.0], [-1.0, 0.0, 0.0], [0.0, 0.0, 3.0]],        # +-i and 3
        [[2.0, 1.0, 0.0], [0.0, 2.0, 1.0], [0.0, 0.0, 2.0]],         # Jordan block
    ])
    w, ok = _eig(lib, mats)
    assert ok.all()
    assert np.all(w[0] == 0.0)
    assert sorted(w[1].real) == [1.0, 1.0, 2.0] and np.all(w[1].imag == 0.0)
    assert np.allclose(np.sort_complex(w[2]), [-1j, 1j, 3.0], atol=1e-15)
    assert np.allclose(w[3], 2.0, atol=1e-15)


def test_einstein_tensor_blocks_match_lapack(lib):
    """On the G^mu_nu of a real slice: eigenvalues within 1e-10 of the matrix scale, identical Type-I mask."""
    rec = load_golden("golden_einstein.json")["ein_plot_n101"]
    bx, by, dx, dy = einstein_inputs(rec)
    o = oracle_np.einstein_eigenvalues(bx, by, dx, dy)
    g = np.moveaxis(o["G_mixed"], (0, 1), (2, 3)).reshape(-1, 4, 4)
    assert np.all(g[:, 3, :3] == 0.0) and np.all(g[:, :3, 3] == 0.0)        # the (t,x,y) block decouples exactly
    w3, ok = _eig(lib, g[:, :3, :3])
    assert ok.all()
    w = np.concatenate([w3, g[:, 3, 3][:, None].astype(complex)], axis=1)
    ref = np.moveaxis(o["eig_all"], 0, -1).reshape(-1, 4)
    scale = np.maximum(np.abs(g).max(axis=(1, 2)), 1e-300)
    assert (np.abs(np.sort_complex(w) - np.sort_complex(ref)).max(axis=1) / scale).max() < 1e-10
    mine = np.all(np.abs(w.imag) < 1e-8, axis=1)
    assert np.array_equal(mine, o["is_type_i"].reshape(-1))
