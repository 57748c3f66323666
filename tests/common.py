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


def real_problem(rng, pkg_or_oracle, ttype, M, N, K, fp, inplace, valid_nyquist=True, pollute=False):
    """Seeded buffers for a 1d r2c (ttype 1) / c2r (ttype 2) problem with default strides.
    Returns (istride, ostride, input buffer, output dtype, number of output elements)."""
    ist, ost = pkg_or_oracle.default_strides(1, [M, N, K], ttype, inplace)
    nin, nout = K * ist[2], K * ost[2]
    if ttype == R2C:
        x = rng.uniform(-1.0, 1.0, nin).astype(rdtype(fp))
        return ist, ost, x, cdtype(fp), nout
    x = (rng.uniform(-1.0, 1.0, nin) + 1j * rng.uniform(-1.0, 1.0, nin)).astype(cdtype(fp))
    for k in range(K):
        if valid_nyquist and N % 2 == 0:
            s = k * ist[2] + (N // 2) * ist[1]
            x[s: s + M] = x[s: s + M].real
        s = k * ist[2]
        # imag(X[0]) must be ignored by c2r (reference test/r2c.cpp:310-324)
        x[s: s + M] = x[s: s + M].real + (1j * (1.0 + np.arange(M) + k) if pollute else 0)
    return ist, ost, x, rdtype(fp), nout


def addressed_mask(M, nrow, K, stride, size):
    mask = np.zeros(size, dtype=bool)
    for k in range(K):
        for n in range(nrow):
            s = k * stride[2] + n * stride[1]
            mask[s: s + M] = True
    return mask


def analytic_r2c(M, N, K, fp):
    """Reference test/r2c.cpp:40-131 in 1d: x = scale(k) cos(2 pi b n / N) / N with b = (m+k) mod N
    <-> X[n] = scale(k) (delta(n-b) + delta(n+b)) / 2 on n = 0..N/2."""
    k = np.arange(K).reshape(K, 1, 1)
    n = np.arange(N).reshape(1, N, 1)
    m = np.arange(M).reshape(1, 1, M)
    b = (m + k) % N
    scale = 1.0 + k / float(K)
    x = scale * np.cos(2 * np.pi * b * n / N) / N
    nh = np.arange(N // 2 + 1).reshape(1, N // 2 + 1, 1)
    X = scale * (((nh - b) % N == 0).astype(float) + ((nh + b) % N == 0).astype(float)) / 2.0
    return x.astype(rdtype(fp)), X.astype(cdtype(fp))
