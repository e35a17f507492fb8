import sys, os
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "tests")); sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import numpy as np
import oracle_lib
from planarslam_b200 import synth
from planarslam_b200.lines import LineSegment
g = np.stack([synth.render_frame(seed=s, frame=3 * s)[0] for s in range(4)])
ls = LineSegment(max_batch=4)
res = ls.detect(g, 2)
segs, width, prec, nfa = res[1]
osegs, owidth, oprec, onfa = oracle_lib.lsd_detect(g[1], 2)
d = np.abs(nfa - onfa)
print("n differing > 1e-6:", (d > 1e-6).sum(), "of", len(d))
for i in np.argsort(-d)[:12]:
    print(i, segs[i], "w", width[i], "p", prec[i], "gpu nfa", nfa[i], "orc nfa", onfa[i])
