"""Generate tests/golden/reference_vectors.npz from the UNMODIFIED reference.

Run in the build container (needs /root/reference, compiled by `make -C oracle ref` into
oracle/_ref/libbbfft_refemu.so): every case is a seeded random input pushed through the
reference's own planner + generated OpenCL-C kernel under the host work-item emulator.  The
.npz stores inputs' seeds, shapes and the reference outputs, so the GPU box (where the
reference is absent) can check parity against "the reference on the same inputs".
    python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import refemu  # noqa: E402

# (name, type, fp, dir, M, N, K, inplace)
CASES = [
    ("c2c_f32_M1_N64_K8", 0, 4, -1, 1, 64, 8, False),       # C1 shape (small K)
    ("c2c_f32_M16_N8_K4", 0, 4, -1, 16, 8, 4, False),
    ("c2c_f32_M16_N64_K3", 0, 4, 1, 16, 64, 3, False),
    ("c2c_f32_M16_N105_K2", 0, 4, -1, 16, 105, 2, False),
    ("c2c_f32_M16_N512_K2", 0, 4, -1, 16, 512, 2, False),
    ("c2c_f32_M3_N30_K5", 0, 4, -1, 3, 30, 5, True),
    ("c2c_f64_M16_N343_K2", 0, 8, -1, 16, 343, 2, False),
    ("c2c_f64_M16_N16_K4", 0, 8, 1, 16, 16, 4, False),
    ("c2c_f64_M2_N500_K2", 0, 8, -1, 2, 500, 2, True),
    ("c2c_f64_M1_N11_K7", 0, 8, -1, 1, 11, 7, False),
    ("r2c_f32_M1_N256_K4", 1, 4, -1, 1, 256, 4, False),     # C3 shape (small K)
    ("r2c_f32_M1_N256_K4_ip", 1, 4, -1, 1, 256, 4, True),
    ("c2r_f32_M1_N256_K4", 2, 4, 1, 1, 256, 4, False),
    ("c2r_f32_M1_N256_K4_ip", 2, 4, 1, 1, 256, 4, True),
    ("r2c_f64_M4_N105_K3", 1, 8, -1, 4, 105, 3, False),
    ("c2r_f64_M5_N11_K5", 2, 8, 1, 5, 11, 5, False),
    ("r2c_f32_M16_N30_K4", 1, 4, -1, 16, 30, 4, False),
    ("c2r_f64_M16_N48_K3_ip", 2, 8, 1, 16, 48, 3, True),
    ("r2c_f64_M2_N15_K5", 1, 8, -1, 2, 15, 5, False),       # odd N, odd K (unpaired last row)
]


def make_input(rng, ttype, fp, M, N, K, istride):
    """Buffers are allocated to the full strided extent K * stride[2] (in elements)."""
    rdt = np.float32 if fp == 4 else np.float64
    cdt = np.complex64 if fp == 4 else np.complex128
    if ttype == 1:  # real input
        buf = np.zeros(K * istride[2], dtype=rdt)
        vals = rng.uniform(0.0, 1.0, size=(K, N, M)).astype(rdt)
        for k in range(K):
            for n in range(N):
                buf[k * istride[2] + n * istride[1]: k * istride[2] + n * istride[1] + M] = vals[k, n]
        return buf
    nin = N if ttype == 0 else N // 2 + 1
    buf = np.zeros(K * istride[2], dtype=cdt)
    vals = (rng.uniform(0.0, 1.0, size=(K, nin, M)) + 1j * rng.uniform(0.0, 1.0, size=(K, nin, M))).astype(cdt)
    if ttype == 2:
        vals[:, 0, :] = vals[:, 0, :].real + 1j * 0.25  # "polluted" imag(X[0]) must be ignored
        if N % 2 == 0:
            vals[:, N // 2, :] = vals[:, N // 2, :].real
    for k in range(K):
        for n in range(nin):
            buf[k * istride[2] + n * istride[1]: k * istride[2] + n * istride[1] + M] = vals[k, n]
    return buf


def out_buffer(ttype, fp, K, ostride):
    rdt = np.float32 if fp == 4 else np.float64
    cdt = np.complex64 if fp == 4 else np.complex128
    return np.zeros(K * ostride[2], dtype=rdt if ttype == 2 else cdt)


def main():
    out = {}
    names = []
    for i, (name, ttype, fp, d, M, N, K, inplace) in enumerate(CASES):
        rng = np.random.default_rng(1000 + i)
        cfg = refemu.make_config(1, [M, N, K], fp, d, ttype, inplace=inplace)
        istride, ostride = list(cfg.istride), list(cfg.ostride)
        x = make_input(rng, ttype, fp, M, N, K, istride)
        plan = refemu.Plan(cfg)
        if inplace:
            # one byte buffer viewed as input and output type
            rdt = np.float32 if fp == 4 else np.float64
            nbytes = max(x.nbytes, out_buffer(ttype, fp, K, ostride).nbytes)
            raw = np.zeros(nbytes, dtype=np.uint8)
            raw[: x.nbytes] = x.view(np.uint8)
            plan.execute(raw)
            y = raw.view(out_buffer(ttype, fp, K, ostride).dtype)[: K * ostride[2]].copy()
        else:
            y = out_buffer(ttype, fp, K, ostride)
            plan.execute(x, y)
        out[name + "__in"] = x
        out[name + "__out"] = y
        out[name + "__meta"] = np.array([ttype, fp, d, M, N, K, int(inplace)] + istride[:3] + ostride[:3], dtype=np.int64)
        names.append(name)
        print(name, plan.kernel_names)
    out["names"] = np.array(names)
    np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "reference_vectors.npz"), **out)


if __name__ == "__main__":
    main()
