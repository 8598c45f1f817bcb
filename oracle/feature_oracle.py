"""TEST INFRASTRUCTURE (not a product path): CPU restatement of the reference's per-sample feature assembly and padding, the
checker for vln-hamt_b200/feature_store.py.

Follows pretrain_src/data/r2r_data.py: angle_feature :14-17, get_point_angle_feature :19-32, get_history_feature :264-308 (image /
pano / angle / class-probability parts), get_ob_pano_view :187-189 (image and angle parts), get_image_feature :310-323, and
data/common.py:5-20 (pad_tensors).

Pinned: tests/test_oracle.py::test_feature_oracle_matches_reference runs it against the UNMODIFIED reference methods (imported with
stub modules for the absent jsonlines / h5py, which those methods do not use once the feature cache is filled) and against the golden
file tests/golden/feature_assembly.pt produced from the reference by oracle/make_golden_features.py.
"""
from __future__ import annotations

import math
from typing import Dict, List, Sequence

import numpy as np


def angle_feature(heading: float, elevation: float, size: int) -> np.ndarray:
    return np.array([math.sin(heading), math.cos(heading), math.sin(elevation), math.cos(elevation)] * (size // 4), dtype=np.float32)


def point_angle_feature(size: int, base_view: int) -> np.ndarray:
    feat = np.empty((36, size), np.float32)
    base_heading = (base_view % 12) * math.radians(30)
    heading = elevation = 0.0
    for ix in range(36):
        if ix == 0:
            heading, elevation = 0, math.radians(-30)
        elif ix % 12 == 0:
            heading = 0
            elevation += math.radians(30)
        else:
            heading += math.radians(30)
        feat[ix] = angle_feature(heading - base_heading, elevation, size)
    return feat


def softmax(x: np.ndarray) -> np.ndarray:
    e = np.exp(x)                                    # r2r_data.py:93-96 (no max subtraction)
    return e / np.sum(e, axis=1, keepdims=True)


def history(fts: Dict[str, np.ndarray], scan: str, path: Sequence[str], path_view: Sequence[int], t_cur: int, D: int, A: int):
    """(hist_img [t,D], hist_pano_img [t,36,D], hist_pano_ang [t,36,A], hist_probs [t,P]) for the steps before t_cur."""
    img, pano, pang, probs = [], [], [], []
    for t in range(t_cur):
        v = fts["%s_%s" % (scan, path[t])]
        img.append(v[path_view[t], :D])
        pano.append(v[:, :D])
        pang.append(point_angle_feature(A, path_view[t]))
        probs.append(v[path_view[t], D:])
    if t_cur > 0:
        return np.stack(img, 0), np.stack(pano, 0), np.stack(pang, 0), softmax(np.stack(probs, 0))
    P = next(iter(fts.values())).shape[1] - D
    return (np.zeros((0, D), np.float32), np.zeros((0, 36, D), np.float32), np.zeros((0, 36, A), np.float32), np.zeros((0, P), np.float32))


def observation(fts: Dict[str, np.ndarray], scan: str, vp: str, view: int, D: int, A: int):
    """(ob_img [37,D], ob_ang [37,A]): 36 views + the all-zero STOP row."""
    v = fts["%s_%s" % (scan, vp)]
    img = np.vstack([v, np.zeros((1, v.shape[-1]), v.dtype)])[:, :D]
    ang = np.vstack([point_angle_feature(A, view), np.zeros((1, A), np.float32)])
    return img, ang


def pad(arrs: List[np.ndarray]) -> np.ndarray:
    """pad_tensors: B x [T_i, ...] -> [B, max T, ...] zero padded."""
    T = max(a.shape[0] for a in arrs)
    out = np.zeros((len(arrs), T) + arrs[0].shape[1:], arrs[0].dtype)
    for i, a in enumerate(arrs):
        out[i, :a.shape[0]] = a
    return out
