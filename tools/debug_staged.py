"""Debug aid (GPU): repeated executes of the staged tile kernel against the plain one; prints which tiles differ."""
import importlib, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pkg = importlib.import_module("double-batched-fft-library_b200")

def run(tune, K, reps, fp=4, n=128, fill=None):
    cfg = pkg.make_config(2, [1, n, n, K], fp, pkg.FORWARD, pkg.C2C, inplace=False)
    rdt = torch.float32 if fp == 4 else torch.float64
    x = torch.view_as_complex(torch.rand(K * n * n, 2, dtype=rdt, device="cuda"))
    ref = torch.empty_like(x)
    p0 = pkg.Plan(cfg, stream=torch.cuda.current_stream().cuda_stream, tune="SG=0")
    p0.execute(x, ref)
    torch.cuda.synchronize()
    p = pkg.Plan(cfg, stream=torch.cuda.current_stream().cuda_stream, tune=tune)
    bad_runs = 0
    for r in range(reps):
        y = torch.full_like(x, float("nan")) if fill is None else torch.zeros_like(x)
        p.execute(x, y)
        torch.cuda.synchronize()
        same = (torch.view_as_real(y) == torch.view_as_real(ref)).all(dim=-1).view(K, n * n)
        bad = (~same.all(dim=1)).nonzero().flatten().tolist()
        if bad:
            bad_runs += 1
            t = bad[0]
            nb = int((~same[t]).sum())
            nan = int(torch.isnan(torch.view_as_real(y).view(K, -1)[t]).any())
            print("  %s K=%d run %d: %d bad tiles %s; tile %d has %d bad elements, nan=%d" % (tune, K, r, len(bad), bad[:12], t, nb, nan))
    print("%s K=%d: %d of %d runs bad (%s, PDL=%s)" % (tune, K, bad_runs, reps, p.kernel_names[0][-24:], os.environ.get("BBFFT_CUDA_PDL", "default")))
    p.close(); p0.close()

if __name__ == "__main__":
    for K in (700, 149, 296, 8192):
        for tune in ("SG=-1,BK=1", "SG=-1,BK=0", "PS=1,SG=0"):
            run(tune, K, 12 if K < 8000 else 4)
