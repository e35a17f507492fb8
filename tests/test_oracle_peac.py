"""CPU: sanity of the PEAC oracle's own building blocks (no GPU)."""
import numpy as np

import oracle_lib
from planarslam_b200 import synth


def test_jacobi_eigensolver_against_numpy():
    L = oracle_lib.lib()
    rng = np.random.default_rng(0)
    for _ in range(300):
        A = rng.normal(size=(3, 3)) * 10 ** rng.uniform(-6, 3)
        K = np.ascontiguousarray(A @ A.T)
        s = np.zeros(3)
        V = np.zeros((3, 3))
        L.orc_eig33(K.ctypes.data, s.ctypes.data, V.ctypes.data)
        w = np.linalg.eigvalsh(K)
        assert s[0] <= s[1] <= s[2]
        assert np.allclose(s, w, rtol=1e-9, atol=1e-12 * abs(w).max())
        assert np.allclose(K @ V, V * s, atol=1e-9 * abs(w).max())
        assert np.allclose(V.T @ V, np.eye(3), atol=1e-12)


def test_heap_is_libstdcxx_priority_queue():
    """The explicit binary heap used by the oracle (and restated on the GPU) pops in exactly the order of
    std::priority_queue with the reference comparator, including tied keys."""
    L = oracle_lib.lib()
    rng = np.random.default_rng(1)
    for trial in range(20):
        n = 2000
        keys = np.round(rng.random(n) * (5 if trial % 2 else 1e6)) / 4.0     # many exact ties on odd trials
        ops = (rng.random(6000) < 0.55).astype(np.int32)
        assert L.orc_heap_selftest(keys.ctypes.data, n, ops.ctypes.data, len(ops)) == 0


def test_peac_oracle_on_room_corner():
    g, d, z, _ = synth.render_frame(2, 17)
    orc = oracle_lib.PeacOracle(d)
    assert len(orc.planes) == 3 and orc.n_coarse >= 3
    lab = orc.labels
    assert (lab >= 0).mean() > 0.9                       # almost everything is on one of the three planes
    for i, (d8, i2) in enumerate(orc.planes):
        n, c = d8[0:3], d8[3:6]
        assert abs(np.linalg.norm(n) - 1) < 1e-9 and n @ c <= 0
        assert len(orc.membership[i]) == (lab == i).sum()
        assert np.all(np.diff(orc.membership[i]) > 0)
    ns = np.array([p[0][0:3] for p in orc.planes])
    assert abs(ns @ ns.T - np.eye(3)).max() < 0.05       # the three walls are mutually orthogonal
