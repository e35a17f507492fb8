"""GPU: ORB extraction through the C ABI vs the CPU oracle — bit-exact at every stage.

Bar (BASELINE.json north_star): FAST keypoint indices and rBRIEF bits bit-exact; here the whole 28-byte
keypoint records (coordinates, angle, response, octave) and the 32-byte descriptors must be identical.
"""
import numpy as np
import pytest

import oracle_lib
from planarslam_b200 import synth

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ext():
    from planarslam_b200.orb import ORBextractor
    return ORBextractor(1000, 1.2, 8, 20, 7, max_batch=4)


def _frames():
    return [synth.render_frame(2, 0)[0], synth.render_frame(2, 17)[0], synth.polygon_image(1), synth.polygon_image(5)]


def test_stages_match_oracle(ext):
    imgs = np.stack(_frames())
    ext.extract_batch(imgs)
    for f in range(len(imgs)):
        orc = oracle_lib.OrbOracle()
        orc.extract(imgs[f])
        for l in range(8):
            assert np.array_equal(ext.debug_level(f, l), orc.level(l)), f"pyramid level {l} frame {f}"
            assert np.array_equal(ext.debug_level(f, l, blurred=True), oracle_lib.blur(orc.level(l))), f"blur level {l} frame {f}"
            assert np.array_equal(ext.debug_candidates(f, l), orc.candidates(l)), f"FAST candidates level {l} frame {f}"


def test_keypoints_and_descriptors_bit_exact(ext):
    imgs = np.stack(_frames())
    kps, desc = ext.extract_batch(imgs)
    for f in range(len(imgs)):
        ok, od = oracle_lib.orb_extract(imgs[f])
        assert len(kps[f]) == len(ok) and len(ok) >= 900
        assert kps[f].tobytes() == ok.tobytes(), f"keypoint records differ, frame {f}"
        assert np.array_equal(desc[f], od), f"descriptors differ, frame {f}"


def test_single_call_equals_batch(ext):
    img = synth.render_frame(3, 5)[0]
    k1, d1 = ext(img)
    kb, db = ext.extract_batch(np.stack([img, img, img]))
    for i in range(3):
        assert k1.tobytes() == kb[i].tobytes() and np.array_equal(d1, db[i])


def test_edge_cases(ext):
    from planarslam_b200.orb import ORBextractor
    # flat image: no corners anywhere -> zero keypoints, like the reference's empty result
    flat = np.full((480, 640), 128, np.uint8)
    k, d = ext(flat)
    ok, _ = oracle_lib.orb_extract(flat)
    assert len(k) == 0 and len(ok) == 0 and d.shape == (0, 32)
    # pure noise: far more candidates than the quota, exercises the "largest node first" phase of the quadtree
    noise = np.random.default_rng(0).integers(0, 256, (480, 640), dtype=np.uint8)
    k, d = ext(noise)
    ok, od = oracle_lib.orb_extract(noise)
    assert k.tobytes() == ok.tobytes() and np.array_equal(d, od)
    # sparse texture: fewer candidates than the quota on most levels (threshold fallback + early quadtree stop)
    sparse = np.full((480, 640), 90, np.uint8)
    sparse[100:140, 200:260] = 200
    sparse[300:310, 400:500] = 10
    k, d = ext(sparse)
    ok, od = oracle_lib.orb_extract(sparse)
    assert len(ok) > 0 and k.tobytes() == ok.tobytes() and np.array_equal(d, od)
    # other sizes / parameters
    for (w, h, nf, nl) in [(1280, 960, 2000, 8), (320, 240, 500, 4), (752, 480, 1200, 8)]:
        e2 = ORBextractor(nf, 1.2, nl, 20, 7)
        img = synth.render_frame(11, 1, width=w, height=h)[0]
        k, d = e2(img)
        ok, od = oracle_lib.orb_extract(img, nf, 1.2, nl, 20, 7)
        assert len(ok) > 0.8 * nf
        assert k.tobytes() == ok.tobytes() and np.array_equal(d, od), (w, h)
    # empty input -> silent empty return (reference :1046)
    k, d = ext(np.zeros((0, 0), np.uint8))
    assert len(k) == 0


def test_scale_tables(ext):
    ext(synth.polygon_image(1))
    t = oracle_lib.OrbOracle().tables()
    assert np.array_equal(ext.GetScaleFactors(), t[0]) and np.array_equal(ext.GetInverseScaleFactors(), t[1])
    assert np.array_equal(ext.GetScaleSigmaSquares(), t[2]) and np.array_equal(ext.GetInverseScaleSigmaSquares(), t[3])
    assert np.array_equal(ext.GetFeaturesPerLevel(), t[4])
