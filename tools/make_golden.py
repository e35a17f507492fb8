#!/usr/bin/env python3
"""Generates tests/golden/*.npz: outputs of the in-container cv2 4.13 (the upstream implementation behind the reference's OpenCV
calls) on seeded synthetic inputs, so that the oracle and the CUDA path stay pinned to upstream results even where cv2 is not
importable.  Run in the build container:  python tools/make_golden.py"""
import os
import sys

import cv2
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from planarslam_b200 import synth  # noqa: E402

out = os.path.join(ROOT, "tests", "golden")
os.makedirs(out, exist_ok=True)

# ---- line-segment detector: cv2.createLineSegmentDetector on rendered frames (seed, frame) ----
lsd = {}
for seed in (0, 7):
    g = synth.render_frame(seed=seed, frame=3 * seed)[0]
    for refine, flag in ((0, cv2.LSD_REFINE_NONE), (1, cv2.LSD_REFINE_STD), (2, cv2.LSD_REFINE_ADV)):
        r = cv2.createLineSegmentDetector(flag).detect(g)
        segs = r[0].reshape(-1, 4)
        if refine == 2:                      # only the 40 longest are pinned for ADV
            length = np.hypot(segs[:, 2] - segs[:, 0], segs[:, 3] - segs[:, 1])
            segs = segs[np.argsort(-length, kind="stable")[:40]]
        lsd[f"seed{seed}_refine{refine}"] = segs.astype(np.float32)
# what ExtractLineSegment keeps of LSD_REFINE_ADV on the 17 LSD test frames: LSDDetector clamps the end points into the image (x >= cols -> cols - 1, < 0 -> 0),
# KeyLine.response = clamped length / max(cols, rows), the 40 largest responses are kept (ties: detection order)
adv_frames = [synth.render_frame(seed=s, frame=3 * s)[0] for s in range(16)] + [synth.polygon_image(11)]
for k, g in enumerate(adv_frames):
    segs = cv2.createLineSegmentDetector(cv2.LSD_REFINE_ADV).detect(g)[0].reshape(-1, 4).astype(np.float32)
    h, w = g.shape
    segs[:, [0, 2]] = np.where(segs[:, [0, 2]] < 0, np.float32(0), np.where(segs[:, [0, 2]] >= w, np.float32(w - 1), segs[:, [0, 2]]))
    segs[:, [1, 3]] = np.where(segs[:, [1, 3]] < 0, np.float32(0), np.where(segs[:, [1, 3]] >= h, np.float32(h - 1), segs[:, [1, 3]]))
    length = np.sqrt((segs[:, 0] - segs[:, 2]).astype(np.float64) ** 2 + (segs[:, 1] - segs[:, 3]).astype(np.float64) ** 2).astype(np.float32)
    lsd[f"adv_top40_{k}"] = segs[np.argsort(-(length / np.float32(max(w, h))), kind="stable")[:40]]
np.savez_compressed(os.path.join(out, "lsd_cv2_4_13.npz"), **lsd)

# ---- 8-bit primitives on a seeded random image ----
rng = np.random.default_rng(20260923)
img = rng.integers(0, 256, (96, 128), dtype=np.uint8)
prims = dict(img=img,
             blur7_s2=cv2.GaussianBlur(img, (7, 7), 2, 2, borderType=cv2.BORDER_REFLECT_101),
             blur7_s075=cv2.GaussianBlur(img, (7, 7), 0.75),
             resize_linear_107x80=cv2.resize(img, (107, 80), interpolation=cv2.INTER_LINEAR),
             border19=cv2.copyMakeBorder(img, 19, 19, 19, 19, cv2.BORDER_REFLECT_101))
big = synth.render_frame(seed=3, frame=9)[0]
prims["frame_seed3"] = np.array([3, 9])
prims["resize_exact_0p8_crc"] = np.array([int(cv2.resize(cv2.GaussianBlur(big, (7, 7), 0.75), None, fx=0.8, fy=0.8, interpolation=cv2.INTER_LINEAR_EXACT)
                                              .astype(np.uint64).sum())])
np.savez_compressed(os.path.join(out, "cvprims_cv2_4_13.npz"), **prims)

# ---- brute-force Hamming matcher ----
q = rng.integers(0, 256, (300, 32), dtype=np.uint8)
t = rng.integers(0, 256, (280, 32), dtype=np.uint8)
t[:40] = q[:40] ^ np.packbits(rng.random((40, 256)) < 0.05, axis=1)      # near-duplicates
t[40:60] = t[20:40]                                                        # exact ties between train rows
knn = cv2.BFMatcher(cv2.NORM_HAMMING).knnMatch(q, t, k=2)
idx = np.array([[m.trainIdx for m in row] for row in knn], np.int32)
dist = np.array([[int(m.distance) for m in row] for row in knn], np.int32)
np.savez_compressed(os.path.join(out, "bfmatcher_cv2_4_13.npz"), q=q, t=t, idx=idx, dist=dist)
print("wrote", sorted(os.listdir(out)))
