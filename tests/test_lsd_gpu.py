"""GPU: the line-segment detector through the C ABI vs the CPU oracle (itself pinned to cv2 4.13, tests/test_oracle_lsd.py).

Bar: bit-exact segments (float32 end points, detection order), widths and precisions for all three refinement modes; the
log-NFA values go through log / pow / sinh / exp of two different libms and are compared to 1e-9; stage products (scaled
image, gradient norm, level-line angle, seed order) bit-exact; KeyLine records and line functions of the 40 longest segments
bit-exact except KeyLine.angle (atan2f vs a correctly rounded double atan2: 1 ulp)."""
import numpy as np
import pytest

import oracle_lib
from planarslam_b200 import synth

pytestmark = pytest.mark.gpu


def _frames(n=4):
    return np.stack([synth.render_frame(seed=s, frame=3 * s)[0] for s in range(n)])


def test_lsd_stages_bit_exact():
    from planarslam_b200.lines import LineSegment
    g = _frames(2)
    ls = LineSegment(max_batch=2)
    ls.detect(g, 1)
    for f in range(2):
        st, o = ls.debug_stage(f), oracle_lib.lsd_stages(g[f], 1)
        assert np.array_equal(st["scaled"], o["scaled"])
        assert np.array_equal(st["modgrad"][:-1, :-1], o["modgrad"][:-1, :-1])
        assert np.array_equal(st["angles"], o["angles"])
        defined = o["angles"].ravel()[o["order"]] != -1024.0                   # the GPU orders only the pixels that can seed a region
        assert np.array_equal(st["order"], o["order"][defined])


@pytest.mark.parametrize("refine", [0, 1, 2])
def test_lsd_segments_match_oracle(refine):
    from planarslam_b200.lines import LineSegment
    g = np.concatenate([_frames(3), synth.polygon_image(11)[None]])
    ls = LineSegment(max_batch=4)
    res = ls.detect(g, refine)
    for f in range(4):
        segs, width, prec, nfa = res[f]
        osegs, owidth, oprec, onfa = oracle_lib.lsd_detect(g[f], refine)
        assert len(segs) == len(osegs) > 50, (f, len(segs), len(osegs))
        assert np.array_equal(segs, osegs), f
        assert np.array_equal(width, owidth) and np.array_equal(prec, oprec), f          # same deterministic sincos on both sides
        assert np.allclose(nfa, onfa, rtol=1e-9, atol=1e-9), f


def test_extract_line_segments_match_oracle():
    from planarslam_b200.lines import LineSegment
    g = _frames(3)
    ls = LineSegment(max_batch=3)
    res = ls.ExtractLineSegment(g, 40)
    for f in range(3):
        kl, lf = res[f]
        okl, olf = oracle_lib.extract_line_segments(g[f], 40)
        assert len(kl) == len(okl) == 40
        for name in kl.dtype.names:
            if name == "angle":
                assert np.allclose(kl[name], okl[name], rtol=3e-7, atol=1e-7)
            else:
                assert np.array_equal(kl[name], okl[name]), (f, name)
        assert np.array_equal(lf, olf), f


def test_lsd_flat_image_and_box():
    from planarslam_b200.lines import LineSegment
    ls = LineSegment(max_batch=1)
    flat = np.full((480, 640), 77, np.uint8)
    assert len(ls.detect(flat, 2)[0]) == 0
    box = np.zeros((480, 640), np.uint8)
    box[130:390, 140:520] = 200
    segs = ls.detect(box, 1)[0]
    osegs = oracle_lib.lsd_detect(box, 1)[0]
    assert len(segs) == 4 and np.array_equal(segs, osegs)


def test_lsd_published_enumeration_matches_oracle():
    """pslam_lsd_set_rect_enumeration(ctx, 0): the published LSD rectangle iterator (non-default) against oracle variant 0; the default
    (1 = cv2 4.x rect_nfa, oracle variant 1, same deterministic sincos) is what every other test in this file runs."""
    from planarslam_b200.lines import LineSegment
    g = np.concatenate([_frames(3), synth.polygon_image(11)[None]])
    ls = LineSegment(max_batch=4)
    ls.set_rect_enumeration(0)
    res = ls.detect(g, 2)
    for f in range(4):
        segs, width, prec, nfa = res[f]
        osegs, owidth, oprec, onfa = oracle_lib.lsd_detect(g[f], 2, rect_enum=0)
        assert len(segs) == len(osegs) > 50, (f, len(segs), len(osegs))
        assert np.array_equal(segs, osegs), f
        assert np.array_equal(width, owidth) and np.array_equal(prec, oprec), f
        assert np.allclose(nfa, onfa, rtol=1e-9, atol=1e-9), f
    ls.set_rect_enumeration(1)
    segs1 = ls.detect(g[:1], 2)[0][0]
    assert np.array_equal(segs1, oracle_lib.lsd_detect(g[0], 2, rect_enum=1)[0])


@pytest.mark.parametrize("size", [(1280, 960), (320, 240), (752, 480)])
def test_lsd_other_image_sizes(size):
    """BASELINE.json config 5 (1280x960) and two more sizes: segments, KeyLines and LBD descriptors against the oracle (capacities scale with the image area)."""
    from planarslam_b200.lines import LineSegment
    w, h = size
    g = np.stack([synth.render_frame(seed=5, frame=2, width=w, height=h)[0], synth.render_frame(seed=6, frame=9, width=w, height=h)[0]])
    from planarslam_b200._lib import Context
    ls = LineSegment(ctx=Context(w, h, 2, nlevels=4 if w < 640 else 8))          # (the context also builds the ORB pyramid geometry: 8 levels need >= 640 columns)
    res = ls.detect(g, 2)
    for f in range(2):
        segs, width, prec, nfa = res[f]
        osegs, owidth, oprec, onfa = oracle_lib.lsd_detect(g[f], 2, cap=65536)
        assert len(segs) == len(osegs) > 20, (f, len(segs), len(osegs))
        assert np.array_equal(segs, osegs), f
        assert np.array_equal(width, owidth) and np.array_equal(prec, oprec), f
    kres = ls.ExtractLineSegment(g, 40)
    for f in range(2):
        kl, lf = kres[f]
        okl, olf = oracle_lib.extract_line_segments(g[f], 40)
        assert len(kl) == len(okl)
        for name in ("startPointX", "startPointY", "endPointX", "endPointY", "lineLength", "response", "class_id"):
            assert np.array_equal(kl[name], okl[name]), (f, name)
        assert np.array_equal(lf, olf)
