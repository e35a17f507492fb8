"""ncu target: one ExtractLineSegment call on 148 frames (one warp per SM)."""
import sys, os
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import numpy as np
from planarslam_b200 import synth
from planarslam_b200.lines import LineSegment
n = int(sys.argv[1]) if len(sys.argv) > 1 else 148
g = np.stack([synth.render_frame(seed=s % 8, frame=3 * (s % 8))[0] for s in range(n)])
ls = LineSegment(max_batch=n)
ls.ExtractLineSegment(g, 40)
