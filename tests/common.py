"""Shared helpers for the parity tests (test infrastructure)."""
import numpy as np

C2C, R2C, C2R = 0, 1, 2

TOL = {4: 1e-5, 8: 1e-12}  # relative L2 gates of BASELINE.json (fp32 / fp64)


def rel_l2(a, b):
    a = np.asarray(a)
    b = np.asarray(b)
    ca = a.astype(np.complex128 if np.iscomplexobj(a) else np.float64).ravel()
    cb = b.astype(np.complex128 if np.iscomplexobj(b) else np.float64).ravel()
    nb = np.linalg.norm(cb)
    return float(np.linalg.norm(ca - cb) / (nb if nb > 0 else 1.0))


def cdtype(fp):
    return np.complex64 if fp == 4 else np.complex128


def rdtype(fp):
    return np.float32 if fp == 4 else np.float64


def random_complex(rng, shape, fp):
    return (rng.standard_normal(shape) + 1j * rng.standard_normal(shape)).astype(cdtype(fp))


def reference_tol(n, fp):
    """The reference's own per-component tolerance: 1e2 * eps * sqrt(N) (test/fft.hpp:17-19)."""
    eps = np.finfo(rdtype(fp)).eps
    return 1e2 * eps * np.sqrt(n)


def analytic_c2c_input(M, N, K, fp, scale_K=None):
    """Single Fourier mode per (m,k) column, reference test/c2c.cpp:28-46:
    x[m,n,k] = (1 + k/K) * exp(+2 pi i ((m+k) mod N) n / N) / N  ->  X[m,n,k] = (1+k/K) delta(n - (m+k) mod N)."""
    k = np.arange(K).reshape(K, 1, 1)
    n = np.arange(N).reshape(1, N, 1)
    m = np.arange(M).reshape(1, 1, M)
    scale = 1.0 + k / float(K)
    mode = (m + k) % N
    x = scale * np.exp(2j * np.pi * mode * n / N) / N
    X = scale * (n == mode)
    return x.astype(cdtype(fp)), X.astype(cdtype(fp))
