#!/usr/bin/env python3
"""Random-scene STOCHASTIC-mode check (CPU only, test infrastructure; needs oracle/_ref).

For each seed: the scene of fuzz_parity.py (with an area light now and then), rendered at 48x32 and
128 spp twice by the reference (two RNG seeds) and once by our device code compiled for the CPU.
Prints the reference's own seed-to-seed RMSE, ours against the reference pair's mean (scaled to the
same variance), and the difference of the image means; lines that look off are flagged `<<<<`.
GI, area-light sampling, glossy refraction and the light-selection rule only run in this mode.

usage: fuzz_stochastic.py FIRST_SEED END_SEED
"""
import os
import random
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'tests'))
sys.path.insert(0, os.path.join(ROOT, 'tools'))
import fuzz_parity as fp
from aovio import read_aov
WORK = '/tmp/bhrt_fuzz_stochastic'
os.makedirs(WORK, exist_ok=True)
def run(binary, path, out, seed, spp):
    args=[binary, path, out, "--spp", str(spp), "--gi", "3", "--seed", str(seed)]
    if 'bhrt_ref' in binary: args += ["--threads","4"]
    subprocess.check_output(args, cwd=ROOT, stderr=subprocess.DEVNULL)
    return read_aov(out+".aov")["linear"]
lo,hi=int(sys.argv[1]),int(sys.argv[2]); spp=128
for seed in range(lo,hi):
    path=WORK + "/f%d.xml"%seed
    xml=fp.scene_xml(seed, 48, 32)
    # give point lights a size now and then (area lights exist only in the stochastic mode)
    rng=random.Random(seed*7+1)
    if rng.random()<0.6: xml=xml.replace("<position ", '<size value="%g"/>\n      <position '%rng.choice([0.5,1,3]),1)
    open(path,"w").write(xml)
    try:
        r1=run(fp.REF_BIN,path,WORK + "/r1",1,spp); r2=run(fp.REF_BIN,path,WORK + "/r2",2,spp); e=run(fp.EMU_BIN,path,WORK + "/e",3,spp)
    except subprocess.CalledProcessError as ex:
        print(seed,"crash",os.path.basename(ex.cmd[0])); continue
    if not (np.isfinite(r1).all() and np.isfinite(r2).all() and np.isfinite(e).all()):
        print(seed, "nonfinite", np.isfinite(r1).all(), np.isfinite(e).all()); continue
    rm=(r1+r2)/2
    own=np.sqrt(((r1-r2)**2).mean()); ours=np.sqrt(((e-rm)**2).mean())*np.sqrt(2/1.5)   # var(e-rm)=1.5 var -> compare with var(r1-r2)=2 var
    dm=e.mean()-rm.mean(); sd=abs(r1.mean()-r2.mean())
    flag = "  <<<<" if (ours>1.35*own+0.004 or abs(dm)>3*sd+0.004) else ""
    print("seed %d rmse own %.4f ours %.4f  mean ref %.4f diff %+.4f (ref pair %.4f)%s"%(seed,own,ours,rm.mean(),dm,sd,flag))
