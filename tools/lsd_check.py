"""Debug helper (GPU box): LSD through the ABI vs the oracle, prints the first differences and timings."""
import sys, time, os
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "tests")); sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import numpy as np
import oracle_lib
from planarslam_b200 import synth
from planarslam_b200.lines import LineSegment
g = np.stack([synth.render_frame(seed=s, frame=3 * s)[0] for s in range(4)])
ls = LineSegment(max_batch=3108)
for refine in (2,):
    res = ls.detect(g, refine)
    for f in range(4):
        segs, width, prec, nfa = res[f]
        osegs, owidth, oprec, onfa = oracle_lib.lsd_detect(g[f], refine)
        m = min(len(segs), len(osegs))
        eq = (segs[:m] == osegs[:m]).all(1)
        print("refine", refine, "frame", f, "n", len(segs), len(osegs), "first mismatch", None if eq.all() else int(np.argmin(eq)), "equal", int(eq.sum()))
        if not eq.all():
            i = int(np.argmin(eq))
            print("   gpu", segs[i], width[i], prec[i], nfa[i]); print("   orc", osegs[i], owidth[i], oprec[i], onfa[i])
        else:
            print("   nfa maxdiff", np.abs(nfa[:m] - onfa[:m]).max(), "prec eq", np.array_equal(prec[:m], oprec[:m]), "width rel", np.abs(width[:m] / owidth[:m] - 1).max())
big = np.concatenate([g] * 64)
for name, fn in (("detect NONE", lambda: ls.detect(big, 0)), ("detect STD", lambda: ls.detect(big, 1)), ("detect ADV", lambda: ls.detect(big, 2)), ("extract 40", lambda: ls.ExtractLineSegment(big, 40))):
    fn(); t = time.time(); fn(); dt = time.time() - t
    print(f"{name}: {len(big)} frames in {dt*1e3:.1f} ms (host-pointer call) -> {len(big)/dt:.0f} fps")
ls.ctx.profile(True); ls.ExtractLineSegment(big, 40); print(ls.ctx.profile_report())
huge = np.concatenate([g] * 777)
ls.ctx.profile(False); ls.ExtractLineSegment(huge, 40)
ls.ctx.profile(True); ls.ExtractLineSegment(huge, 40); rep = ls.ctx.profile_report(); print(len(huge), rep, "-> %.0f fps (kernels only)" % (len(huge) / (sum(v[1] for v in rep.values()) * 1e-3)))
t = time.time(); oracle_lib.lsd_detect(g[0], 2); print("oracle one frame ms", (time.time() - t) * 1e3)
