#!/usr/bin/env python3
"""GPU probe of the TMA-staged blur: runs the ORB extractor on a few frames in child processes - per-level kernel (PSLAM_NO_TMA=1), tensor maps read from
global memory (default), tensor maps read from the kernel parameter block (PSLAM_TMA_MAPS=param) - and checks each against the CPU oracle."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHILD = r'''
import sys, numpy as np
sys.path.insert(0, %r); sys.path.insert(0, %r + "/tests")
from planarslam_b200 import synth
from planarslam_b200.orb import ORBextractor
import oracle_lib
imgs = np.stack([synth.render_frame(2, f)[0] for f in (0, 9)])
ext = ORBextractor(1000, 1.2, 8, 20, 7, max_batch=2)
kps, desc = ext.extract_batch(imgs)
ok = True
for f in range(2):
    ok_f = True
    okps, odesc = oracle_lib.orb_extract(imgs[f])
    ok_f = len(kps[f]) == len(okps) and kps[f].tobytes() == okps.tobytes() and np.array_equal(desc[f], odesc)
    ok = ok and ok_f
print("RESULT", "bit-exact" if ok else "MISMATCH")
''' % (ROOT, ROOT)

for name, env in (("no_tma", {"PSLAM_NO_TMA": "1"}), ("maps_global", {}), ("maps_param", {"PSLAM_TMA_MAPS": "param"})):
    r = subprocess.run([sys.executable, "-c", CHILD], capture_output=True, text=True, env=dict(os.environ, **env), timeout=300)
    tail = (r.stdout.strip().splitlines() or [""])[-1]
    err = [l for l in r.stderr.splitlines() if "rror" in l][-2:]
    print(f"{name}: rc={r.returncode} {tail} {err}")
    sys.stdout.flush()
if "--sanitize" in sys.argv:
    for env in ({}, {"PSLAM_TMA_MAPS": "param"}):
        r = subprocess.run(["compute-sanitizer", "--tool", "memcheck", "--print-limit", "5", sys.executable, "-c", CHILD], capture_output=True, text=True,
                           env=dict(os.environ, **env), timeout=600)
        print("---- compute-sanitizer", env)
        print("\n".join((r.stdout + r.stderr).splitlines()[:40]))
