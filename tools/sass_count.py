"""Instruction histogram of planned kernels (NVRTC, no GPU): python tools/sass_count.py "<descriptor>[:tune]" ..."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import importlib,subprocess,tempfile,collections,re,sys
pkg=importlib.import_module("double-batched-fft-library_b200")
for arg in sys.argv[1:]:
    desc,_,tune=arg.partition(":")
    cfg=pkg.parse_descriptor(desc)
    d=pkg.describe(cfg,tune)
    cub=pkg.compile_to_cubin(d["source"])
    with tempfile.NamedTemporaryFile(suffix=".cubin") as f:
        f.write(cub); f.flush()
        out=subprocess.run(["cuobjdump","-sass",f.name],capture_output=True,text=True).stdout
        res=subprocess.run(["cuobjdump","-res-usage",f.name],capture_output=True,text=True).stdout
    ops=collections.Counter(re.findall(r"^\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d\s+)?([A-Z0-9_]+)",out,re.M))
    tot=sum(ops.values())
    m=re.search(r"REG:(\d+) STACK:(\d+)",res)
    print(d["identifier"][:70],"thr",d["threads"],"regs",m.group(1),"stack",m.group(2),"total",tot,ops.most_common(14))
