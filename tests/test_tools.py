"""CPU tier: the tuning tools' host logic (candidate enumeration, wisdom merge)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))


def test_tuner_candidates_respect_the_real_transform_rules(pkg):
    import tune_gpu
    for n, fp in ((500, 4), (441, 8), (256, 4)):
        c2c = tune_gpu.candidates(n, fp, 16, "c2c")
        r2c = tune_gpu.candidates(n, fp, 16, "r2c")
        c2r = tune_gpu.candidates(n, fp, 16, "c2r")
        assert c2c[0] == r2c[0] == c2r[0] == ""  # the heuristic default always competes
        full = 128 // (2 * fp)
        for t in r2c + c2r:
            if "ML=" in t:
                assert "ML=%d" % full in t  # real candidates keep full lanes (in-place legality)
        # the fused (mirrored) stage gets the smallest radix: last for r2c, first for c2r
        for t in r2c:
            if t.startswith("R=") and "x" in t.split(",")[0]:
                rs = [int(v) for v in t.split(",")[0][2:].split("x")]
                assert rs[-1] == min(rs)
        for t in c2r:
            if t.startswith("R=") and "x" in t.split(",")[0]:
                rs = [int(v) for v in t.split(",")[0][2:].split("x")]
                assert rs[0] == min(rs)
        # every candidate plans (or is rejected cleanly) for the sweep shape
        cfg, K = tune_gpu.make_cfg(pkg, "r2c", fp, n, 16, 1 << 30)
        planned = 0
        for t in r2c[:12]:
            try:
                pkg.describe(cfg, t)
                planned += 1
            except pkg.BbfftError:
                pass
        assert planned >= 6


def test_make_wisdom_merges_without_regressing(tmp_path):
    inc = tmp_path / "wisdom.inc"
    rinc = tmp_path / "wisdom_real.inc"
    inc.write_text('// header\n{4, 100, "R=10x10,T=10,ML=16,BH=1,MB=3"}, // 6000 GB/s (heuristic 5500)\n'
                   '{8, 200, "R=10x20,T=20,ML=8,BH=1,MB=2"}, // 5800 GB/s (heuristic 5000)\n')
    rinc.write_text('// header\n{1, 4, 100, 0, "R=10x5,T=5,ML=16,BH=3,MB=4"}, // 6100 GB/s (heuristic 5000)\n')
    new = {
        # current entry re-measured within 1.5 % of the winner: keep it
        "4,100": {"best": "R=4x25,T=4,ML=16,BH=2,MB=3", "gbs": 6030.0, "default_gbs": 5500.0, "mode": "interleaved",
                  "top": [["R=4x25,T=4,ML=16,BH=2,MB=3", 6030], ["R=10x10,T=10,ML=16,BH=1,MB=3", 6000]]},
        # clear winner: replace
        "8,200": {"best": "R=8x25,T=25,ML=8,BH=1,MB=2", "gbs": 6300.0, "default_gbs": 5000.0, "mode": "interleaved",
                  "top": [["R=8x25,T=25,ML=8,BH=1,MB=2", 6300], ["R=10x20,T=20,ML=8,BH=1,MB=2", 5800]]},
        # new real entry, and one whose best is the heuristic default (no entry)
        "c2r,4,100": {"best": "R=5x10,T=5,ML=16,BH=3,MB=4", "gbs": 6000.0, "default_gbs": 5200.0, "mode": "interleaved", "top": []},
        "r2c,8,64": {"best": "", "gbs": 6400.0, "default_gbs": 6400.0, "mode": "interleaved", "top": []},
    }
    src = tmp_path / "tune.json"
    src.write_text(json.dumps(new))
    subprocess.check_call([sys.executable, os.path.join(ROOT, "tools", "make_wisdom.py"), str(src), "--keep-existing", "--no-pin",
                           "--out", str(inc), "--real-out", str(rinc)])
    w = inc.read_text()
    assert '{4, 100, "R=10x10,T=10,ML=16,BH=1,MB=3"}' in w
    assert '{8, 200, "R=8x25,T=25,ML=8,BH=1,MB=2"}' in w and "R=10x20" not in w
    r = rinc.read_text()
    assert '{1, 4, 100, 0, "R=10x5,T=5,ML=16,BH=3,MB=4"}' in r and '{2, 4, 100, 0, "R=5x10,T=5,ML=16,BH=3,MB=4"}' in r
    assert "{1, 8, 64" not in r
