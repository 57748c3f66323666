"""Diagnostic: the load-callback bit-identity check of tests/cpp/test_cuda_api.cpp (K = 4, outputs
pre-filled with a marker) for every (fp, M, N), reporting where a callback plan differs from the plain one."""
import importlib, os, sys
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
pkg = importlib.import_module("double-batched-fft-library_b200")
from common import cdtype, rdtype
stream = torch.cuda.current_stream().cuda_stream
only = sys.argv[1:]  # e.g. "4,32,212"
for fp in (4, 8):
    for M in (1, 32):
        for N in (8, 64, 212):
            if only and "%d,%d,%d" % (fp, M, N) not in only:
                continue
            K = 4
            real = "float" if fp == 4 else "double"
            rng = np.random.default_rng(N + M)
            Next = 2 * N; ns_ref, ns = Next // 2 + 1, N // 2 + 1
            X = (rng.uniform(0, 1, (K, ns, M)) + 1j * rng.uniform(0, 1, (K, ns, M))).astype(cdtype(fp))
            Xref = np.zeros((K, ns_ref, M), dtype=cdtype(fp)); Xref[:, :ns, :] = X
            src = ("%s2 load(global %s2* in, size_t offset) {\n    size_t m = offset %% %d;\n    size_t n = offset / %d %% %d;\n"
                   "    size_t k = offset / %d;\n    if (n < %d) { return in[m + %d * (n + %d * k)]; }\n    return 0;\n}\n"
                   % (real, real, M, M, ns_ref, M * ns_ref, ns, M, ns))
            strides = dict(istride=[1, M, M * ns_ref], ostride=[1, M, M * Next])
            cref = pkg.make_config(1, [M, Next, K], fp, pkg.BACKWARD, pkg.C2R, **strides)
            ccb = pkg.make_config(1, [M, Next, K], fp, pkg.BACKWARD, pkg.C2R, callbacks=(src, "load", None, "opencl"), **strides)
            outs = []
            for cfg, inp in ((cref, Xref), (ccb, X)):
                plan = pkg.Plan(cfg, stream=stream)
                res = []
                for rep in range(3):
                    xd = torch.from_numpy(inp.reshape(-1)).cuda()
                    yd = torch.full((K * Next * M,), float("nan"), dtype=torch.float32 if fp == 4 else torch.float64, device="cuda")
                    plan.execute(xd, yd)
                    torch.cuda.synchronize()
                    res.append(yd.cpu().numpy())
                outs.append((res, plan.kernel_names[0]))
                plan.close()
            (r0, n0), (r1, n1) = outs
            stable0 = all(np.array_equal(r0[0], r, equal_nan=True) for r in r0)
            stable1 = all(np.array_equal(r1[0], r, equal_nan=True) for r in r1)
            same = np.array_equal(r0[0], r1[0], equal_nan=True)
            nan0, nan1 = int(np.isnan(r0[0]).sum()), int(np.isnan(r1[0]).sum())
            print("fp%d M=%d N=%d same=%s stable(plain,cb)=%s,%s unwritten(plain,cb)=%d,%d" % (fp, M, N, same, stable0, stable1, nan0, nan1), flush=True)
            if not same or not stable0 or not stable1 or nan0 or nan1:
                d = np.nonzero(~((r0[0] == r1[0]) | (np.isnan(r0[0]) & np.isnan(r1[0]))))[0]
                print("   differing elements:", len(d), "first:", [(int(i % M), int(i // M % Next), int(i // (M * Next))) for i in d[:8]])
                print("   values:", [(float(r0[0][i]), float(r1[0][i])) for i in d[:4]])
                print("   ", n0[:110]); print("   ", n1[:110])
