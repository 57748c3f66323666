"""CPU tier: a bounded, seeded run of tools/fuzz_emu.py -- random (type, precision, direction, M, N, K,
strides, in/out-of-place), the planner picks the kernel, the CPU emulator runs the real kernel source,
the long-double oracle judges.  Larger runs (and BBFFT_EMU_RACECHECK=1) are a command line away."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_seeded_fuzz_of_planned_kernels_against_oracle(pkg):
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "fuzz_emu.py"), "--n", "30", "--seed", "7", "--maxn", "96"],
                       capture_output=True, text=True, timeout=1500)
    tail = "\n".join(r.stdout.splitlines()[-40:])
    assert r.returncode == 0 and "done: 0 problems" in r.stdout, tail + r.stderr[-2000:]


def test_seeded_fuzz_of_fused_real_2d_kernels(pkg):
    """Random r2c / c2r 2d configurations (precision, M, N1 x N2, K, placement) through the fused real tile kernels
    under the emulator's race checker, judged by numpy's float64 rfft2 / irfft2."""
    env = dict(os.environ, BBFFT_EMU_RACECHECK="1")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "fuzz_emu.py"), "--n", "0", "--real2d", "45", "--seed", "5"],
                       capture_output=True, text=True, timeout=1500, env=env)
    tail = "\n".join(r.stdout.splitlines()[-40:])
    assert r.returncode == 0 and "done: 0 problems" in r.stdout, tail + r.stderr[-2000:]
    assert r.stdout.count("ok   real2d") >= 8, tail
