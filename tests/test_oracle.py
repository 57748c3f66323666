"""CPU tier: pin the oracle (oracle/bbfft_oracle.c) against the reference's own known answers:
host-side KATs of test/codegen.cpp, the analytic single-mode tests of test/c2c.cpp, numpy, and
the golden vectors produced by the unmodified reference (tests/golden/make_golden.py)."""
import os

import numpy as np
import pytest

from common import TOL, analytic_c2c_input, cdtype, random_complex, rdtype, reference_tol, rel_l2

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "reference_vectors.npz")


def test_factor_golden(oracle):
    # reference test/codegen.cpp:100-125
    assert oracle.factor(4096, 1) == [4096]
    assert oracle.factor(4096, 2) == [64, 64]
    assert oracle.factor(4096, 3) == [16, 16, 16]
    assert oracle.factor(2080, 2) == [40, 52]
    assert oracle.factor(2080, 3) == [10, 13, 16]
    assert oracle.factor(2080, 4) == [4, 5, 8, 13]
    assert oracle.factor(3465, 2) == [55, 63]
    assert oracle.factor(3465, 3) == [11, 15, 21]
    assert oracle.factor(3465, 4) == [5, 7, 9, 11]
    assert oracle.factor(9216, 2) == [96, 96]
    assert oracle.factor(9216, 3) == [16, 24, 24]
    assert oracle.factor(9216, 4) == [8, 8, 12, 12]
    assert oracle.factor(64536, 2) == [24, 2689]
    assert oracle.factor(65536, 3) == [32, 32, 64]
    assert oracle.factor(65536, 4) == [16, 16, 16, 16]


def test_scrambler_golden(oracle):
    # reference test/codegen.cpp:15-69
    f = [3, 2, 2]
    expect = [0, 4, 8, 2, 6, 10, 1, 5, 9, 3, 7, 11]
    for i, e in enumerate(expect):
        assert oracle.scramble(i, f) == e
        assert oracle.unscramble(e, f) == i
    assert oracle.scramble(12, f) == 12 and oracle.scramble(13, f) == 16
    f = [13, 5, 2, 7]
    for i in range(2 * 13 * 5 * 2 * 7):
        assert oracle.scramble(oracle.unscramble(i, f), f) == i
        assert oracle.unscramble(oracle.scramble(i, f), f) == i


def test_trial_division(oracle):
    assert oracle.trial_division(360) == [2, 2, 2, 3, 3, 5]
    assert oracle.trial_division(343) == [7, 7, 7]
    assert oracle.trial_division(13) == [13]


@pytest.mark.parametrize("fp", [4, 8])
@pytest.mark.parametrize("M,N,K", [(1, 64, 5), (16, 8, 3), (3, 105, 2), (2, 512, 2), (5, 11, 4), (17, 27, 2)])
def test_oracle_dft_vs_numpy(oracle, fp, M, N, K):
    rng = np.random.default_rng(7)
    x = random_complex(rng, (K, N, M), fp)
    for d in (-1, 1):
        cfg = oracle.make_config(1, [M, N, K], fp, d, 0, inplace=False)
        y = np.empty_like(x)
        oracle.dft(cfg, x, y)
        ref = np.fft.fft(x.astype(np.complex128), axis=1) if d < 0 else np.fft.ifft(x.astype(np.complex128), axis=1) * N
        assert rel_l2(y, ref) < (2e-7 if fp == 4 else 1e-15)
        y2 = np.empty_like(x)
        oracle.bbfft(cfg, x, y2)
        assert rel_l2(y2, ref) < TOL[fp] * 0.1


@pytest.mark.parametrize("fp", [4, 8])
@pytest.mark.parametrize("M,N,K", [(1, 2, 1), (2, 3, 32), (3, 5, 1), (16, 7, 32), (17, 11, 1), (1, 13, 32), (3, 4, 32),
                                   (16, 16, 1), (2, 32, 32), (1, 128, 1), (3, 256, 2), (2, 512, 2), (3, 27, 32),
                                   (2, 63, 1), (1, 105, 32), (2, 363, 1)])
def test_oracle_analytic_c2c(oracle, fp, M, N, K):
    # reference test/c2c.cpp:17-66 (analytic single mode -> scaled Kronecker delta)
    x, X = analytic_c2c_input(M, N, K, fp)
    cfg = oracle.make_config(1, [M, N, K], fp, -1, 0, inplace=True)
    y = x.copy()
    oracle.bbfft(cfg, y)
    tol = reference_tol(N, fp)
    assert np.max(np.abs(y.real - X.real)) <= tol * 2 and np.max(np.abs(y.imag - X.imag)) <= tol * 2


def _golden_cases():
    if not os.path.exists(GOLDEN):
        return []
    z = np.load(GOLDEN)
    return [str(n) for n in z["names"]]


@pytest.mark.parametrize("name", _golden_cases())
def test_oracle_matches_reference_golden(oracle, name):
    """The oracle restatement reproduces the unmodified reference (emulated) on the same inputs."""
    z = np.load(GOLDEN)
    ttype, fp, d, M, N, K, inplace = [int(v) for v in z[name + "__meta"][:7]]
    istride = [int(v) for v in z[name + "__meta"][7:10]]
    ostride = [int(v) for v in z[name + "__meta"][10:13]]
    x = z[name + "__in"]
    want = z[name + "__out"]
    cfg = oracle.make_config(1, [M, N, K], fp, d, ttype, istride=istride, ostride=ostride)
    if inplace:
        nbytes = max(x.nbytes, want.nbytes)
        raw = np.zeros(nbytes, dtype=np.uint8)
        raw[: x.nbytes] = x.view(np.uint8)
        oracle.dft(cfg, raw)
        got = raw.view(want.dtype)[: want.size]
    else:
        got = np.zeros_like(want)
        oracle.dft(cfg, x, got)
    # compare only the addressed elements (padding is untouched / undefined)
    mask = _addressed_mask(ttype, M, N, K, ostride, want.size, out=True)
    assert rel_l2(got[mask], want[mask]) < TOL[fp] * 0.1


def _addressed_mask(ttype, M, N, K, stride, size, out):
    n_out = N if ttype in (0, 2) else N // 2 + 1
    if not out:
        n_out = N if ttype in (0, 1) else N // 2 + 1
    mask = np.zeros(size, dtype=bool)
    for k in range(K):
        for n in range(n_out):
            s = k * stride[2] + n * stride[1]
            mask[s: s + M] = True
    return mask
