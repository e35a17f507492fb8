import sys, numpy as np
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
from planarslam_b200 import synth
from planarslam_b200.orb import ORBextractor
import oracle_lib
g = synth.render_frame(2, 0)[0]
ext = ORBextractor(1000, 1.2, 8, 20, 7)
k, d = ext(g)
orc = oracle_lib.OrbOracle(); ok, od = orc.extract(g)
for l in range(8):
    a = ext.debug_candidates(0, l); b = orc.candidates(l)
    np.save(f'gpurun_out/cand_gpu_{l}.npy', a); np.save(f'gpurun_out/cand_orc_{l}.npy', b)
    print(l, len(a), len(b), np.array_equal(a, b))
np.save('gpurun_out/kps_gpu.npy', k); np.save('gpurun_out/kps_orc.npy', ok)
print(len(k), len(ok))
