#!/usr/bin/env python
"""Fixture of the makeSearch drop-in test: REAL SIFT keypoints + descriptors of the image the reference ships
(hnsw_sifts_retrieval/data/201505310117china2.jpg, the image makeSearch.cpp:29 reads), computed by cv2.SIFT_create(128) --
what makeSearch.cpp:28-34 computes with cv::xfeatures2d::SIFT::create(128).  Runs only in the build container (needs
/root/reference and cv2); the output travels: tests/golden/makesearch_china2.sift (format: tests/stubs/opencv2/opencv.hpp).

    python tests/gen_makesearch_fixture.py
"""
import os
import struct

import cv2
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
IMG = "/root/reference/hnsw_sifts_retrieval/data/201505310117china2.jpg"
OUT = os.path.join(ROOT, "tests", "golden", "makesearch_china2.sift")

im = cv2.imread(IMG, 1)
assert im is not None
det = cv2.SIFT_create(128)
kps = det.detect(im, None)
kps, desc = det.compute(im, kps)
desc = np.ascontiguousarray(desc, dtype=np.float32)
assert desc.shape[1] == 128 and len(kps) == desc.shape[0] >= 5
with open(OUT, "wb") as f:
    f.write(struct.pack("<i", len(kps)))
    for k in kps:
        f.write(struct.pack("<fffffii", k.pt[0], k.pt[1], k.angle, k.size, k.response, k.class_id, k.octave))
    desc.astype("<f4").tofile(f)
print(f"{OUT}: {len(kps)} keypoints, descriptors {desc.shape}, value range {desc.min():.0f}..{desc.max():.0f}")
