"""Debug helper (GPU box): LBA through the ABI vs the oracle, prints the largest differences and a timing."""
import sys, time, os
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "tests")); sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import numpy as np
import oracle_lib
from planarslam_b200 import synth_lba, synth_pose
from planarslam_b200.lba import LocalBundleAdjuster
ba = LocalBundleAdjuster()
probs = [synth_lba.make_lba_problem(s) for s in range(3)]
res = ba.LocalBundleAdjustmentBatch(probs)
for p, r in zip(probs, res):
    o = oracle_lib.local_bundle_adjustment(p)
    pe = np.array([synth_pose.pose_error(r["kf_Tcw_d"][k], o["kf_Tcw_d"][k]) for k in range(len(r["kf_Tcw_d"]))])
    d = np.abs(r["pt_Xw_d"] - o["pt_Xw_d"]).max(1)
    print("pose", pe.max(0), "pt max", d.max(), "pt median", np.median(d), "n>1e-6", (d > 1e-6).sum(), "line", np.abs(r["line_Xw_d"] - o["line_Xw_d"]).max(),
          "plane", np.abs(r["plane_Xw_d"] - o["plane_Xw_d"]).max())
    print("  it", r["iterations"], o["iterations"], "tr", r["trials"], o["trials"], "chi", r["chi2"], o["chi2"], "lam", r["lambda"], o["lambda"])
    print("  erase diff", (r["erase_pt"] != o["erase_pt"]).sum(), (r["erase_line"] != o["erase_line"]).sum(), [(a != b).sum() for a, b in zip(r["erase_plane"], o["erase_plane"])])
    w = np.argmax(d); print("  worst pt", w, r["pt_Xw_d"][w], o["pt_Xw_d"][w], "obs", (p["pt_obs_pt"] == w).sum())
# points-only problem: everything except se3_exp's libm calls is order-identical
p = synth_lba.make_lba_problem(5, n_kf=6, n_fixed=1, n_points=300, n_pt_obs=1500, n_lines=0, n_line_obs=0, n_plane_obs=(0, 0, 0), outlier_frac=0.0, mono_frac=0.0)
r = ba.LocalBundleAdjustment(p); o = oracle_lib.local_bundle_adjustment(p)
print("points-only: pose", np.abs(r["kf_Tcw_d"] - o["kf_Tcw_d"]).max(), "pt", np.abs(r["pt_Xw_d"] - o["pt_Xw_d"]).max(), r["chi2"], o["chi2"])
for nb in (1, 148, 296):
    pr = [probs[i % 3] for i in range(nb)]
    ba.pack(pr); ba.run_packed(); ba.ctx.synchronize()
    t = time.time(); ba.run_packed(); ba.ctx.synchronize(); dt = time.time() - t
    print(f"batch {nb}: {dt*1e3:.2f} ms -> {nb/dt:.1f} LBA/s")
t = time.time(); oracle_lib.local_bundle_adjustment(probs[0]); print("oracle one problem", (time.time() - t) * 1e3, "ms")
