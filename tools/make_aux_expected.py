"""CPU: expected result signatures for tools/aux_new_kernels.py, from the oracle (run at commit time, here in the build container).
The key lines are the oracle's (published rectangle iterator = what the GPU detector produces, tests/test_lsd_gpu.py)."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, os.path.join(ROOT, "tools"))
import oracle_lib  # noqa: E402
from aux_new_kernels import N_BASE, lines3d_signature, manhattan_signature  # noqa: E402
from planarslam_b200 import synth  # noqa: E402
from planarslam_b200.lines import LINE3D_DTYPE  # noqa: E402
from planarslam_b200.manhattan import MANHATTAN_RESULT_DTYPE  # noqa: E402
from planarslam_b200.synth_manhattan import make_manhattan  # noqa: E402

out = np.zeros((N_BASE, 40), LINE3D_DTYPE)
drawn = np.zeros(N_BASE, np.int64)
for s in range(N_BASE):
    gray, d16, _, _ = synth.render_frame(seed=s, frame=3 * s)
    kl, _ = oracle_lib.extract_line_segments(gray, 40)
    r = oracle_lib.lines3d_frame(kl, d16.astype(np.float32) * np.float32(1.0 / synth.DEPTH_FACTOR), synth.TUM3_K, seed=1)
    n = len(kl)
    out["valid"][s, :n] = r["valid"]; out["n_inliers"][s, :n] = r["n_inliers"]; out["n_points"][s, :n] = r["n_points"]
    out["A"][s, :n] = r["lines3d"][:, :3]; out["B"][s, :n] = r["lines3d"][:, 3:]
    drawn[s] = r["n_drawn"]
res = np.zeros(N_BASE, MANHATTAN_RESULT_DTYPE)
for s in range(N_BASE):
    R_last, normals, dirs, _ = make_manhattan(s)
    r = oracle_lib.track_manhattan_frame(R_last, normals, dirs)
    res["found"][s] = r["found"]; res["n_cone"][s] = r["n_cone"]; res["n_selected"][s] = r["n_selected"]; res["R"][s] = r["R"]
exp = {"lines3d": lines3d_signature(out, drawn), "manhattan": manhattan_signature(res)}
path = os.path.join(ROOT, "tests", "golden", "aux_new_kernels_expected.json")
json.dump(exp, open(path, "w"), indent=1)
print(path, exp["lines3d"])
