"""GPU: size-independent properties at the batch size the benchmark runs (two frames per SM: 296 frames in one call).

The per-frame kernels keep one warp / CTA per frame resident and walk the batch in waves; a frame's result must not depend on its
position in the batch, on its neighbours, or on the wave it lands in: every copy of a frame must give the bytes the single-frame
call gives (which the other GPU tests prove identical to the oracle), and a second run must repeat them (determinism)."""
import numpy as np
import pytest

from planarslam_b200 import synth

pytestmark = pytest.mark.gpu

N_DISTINCT, BATCH = 4, 296


def _tile(a):
    return np.ascontiguousarray(np.tile(a, (BATCH // N_DISTINCT,) + (1,) * (a.ndim - 1)))


def test_orb_full_batch_equals_single_frame_results():
    from planarslam_b200.orb import ORBextractor
    imgs = np.stack([synth.render_frame(4, f)[0] for f in (0, 9, 21, 33)])
    single = ORBextractor(1000, 1.2, 8, 20, 7, max_batch=1)
    ref = [single(im) for im in imgs]
    ext = ORBextractor(1000, 1.2, 8, 20, 7, max_batch=BATCH)
    for run in range(2):
        kps, desc = ext.extract_batch(_tile(imgs))
        for i in range(BATCH):
            k, d = ref[i % N_DISTINCT]
            assert kps[i].tobytes() == k.tobytes() and np.array_equal(desc[i], d), (run, i)


def test_peac_full_batch_equals_single_frame_results():
    from planarslam_b200.planes import PlaneDetection
    K = np.array([[535.4, 0, 320.1], [0, 539.2, 247.6], [0, 0, 1]], np.float32)
    scale = np.float32(1.0 / 5000.0)
    depth = np.stack([synth.render_frame(4, f)[1] for f in (0, 9, 21, 33)])
    ref = PlaneDetection(max_batch=N_DISTINCT).run_batch(depth, K, scale)
    res = PlaneDetection(max_batch=BATCH).run_batch(_tile(depth), K, scale)
    for i in range(BATCH):
        labels, planes, members = res[i]
        rl, rp, rm = ref[i % N_DISTINCT]
        assert np.array_equal(labels, rl), i
        assert len(planes) == len(rp) and all(np.array_equal(planes[name], rp[name]) for name in planes.dtype.names), i
        assert len(members) == len(rm) and all(np.array_equal(a, b) for a, b in zip(members, rm)), i


def test_lines_full_batch_equals_single_frame_results():
    from planarslam_b200.lines import LineSegment
    imgs = np.stack([synth.render_frame(4, f)[0] for f in (0, 9, 21, 33)])
    ref = LineSegment(max_batch=N_DISTINCT).ExtractLineSegment(imgs, 40)
    res = LineSegment(max_batch=BATCH).ExtractLineSegment(_tile(imgs), 40)
    for i in range(BATCH):
        kl, lf = res[i]
        rk, rf = ref[i % N_DISTINCT]
        assert all(np.array_equal(kl[name], rk[name]) for name in kl.dtype.names) and np.array_equal(lf, rf), i
        assert len(kl) == 40
