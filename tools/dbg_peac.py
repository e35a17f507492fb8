import sys, numpy as np
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
from planarslam_b200 import synth
from planarslam_b200.planes import PlaneDetection
import oracle_lib
K = np.array([[535.4, 0, 320.1], [0, 539.2, 247.6], [0, 0, 1]], np.float32)
pd = PlaneDetection(max_batch=4)
depth = np.stack([synth.render_frame(2, f)[1] for f in (0, 17, 40, 55)])
res = pd.run_batch(depth, K, np.float32(1/5000.0))
for f in range(4):
    orc = oracle_lib.PeacOracle(depth[f])
    lab = res[f][0]
    print(f, "equal", np.array_equal(lab, orc.labels), "ndiff", (lab != orc.labels).sum(), "gpu hist", np.unique(lab, return_counts=True), "orc hist", np.unique(orc.labels, return_counts=True)[1])
    np.save(f'gpurun_out/peac_gpu_{f}.npy', lab); np.save(f'gpurun_out/peac_orc_{f}.npy', orc.labels)
    np.save(f'gpurun_out/peac_bm_{f}.npy', pd.debug_coarse(f)[0])
