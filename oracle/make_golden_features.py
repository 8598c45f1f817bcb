"""Generate tests/golden/feature_assembly.pt from the UNMODIFIED reference data code (pretrain_src/data/r2r_data.py,
data/common.py).  The reference module imports jsonlines and h5py at import time (absent here); the methods exercised
(get_history_feature, get_ob_pano_view, get_image_feature with the in-memory cache filled, pad_tensors) never touch them, so empty
stub modules stand in.  Run in the build container:  PYTHONDONTWRITEBYTECODE=1 python oracle/make_golden_features.py
TEST INFRASTRUCTURE ONLY."""
import os
import sys
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
D, A, P, V = 64, 4, 8, 7
SCAN = "scanA"
SAMPLES = [dict(path=["p3", "p1", "p6", "p0", "p2"], views=[5, 17, 30, 0, 12], t_cur=4),
           dict(path=["p2", "p5"], views=[35, 11], t_cur=0),
           dict(path=["p4", "p4", "p0"], views=[24, 1, 13], t_cur=2)]


def scenario():
    g = np.random.RandomState(7)
    keys = ["%s_p%d" % (SCAN, i) for i in range(V)]
    feats = g.randn(V, 36, D + P).astype(np.float32)
    return keys, feats


def reference_db():
    for name in ("jsonlines", "h5py"):
        sys.modules.setdefault(name, types.ModuleType(name))
    sys.path.insert(0, "/root/reference/pretrain_src")
    from data import r2r_data, common
    keys, feats = scenario()
    db = r2r_data.MultiStepNavData.__new__(r2r_data.MultiStepNavData)
    db.image_feat_size, db.image_prob_size, db.angle_feat_size, db.hist_enc_pano = D, P, A, True
    db.in_memory, db._feature_store = True, {k: feats[i] for i, k in enumerate(keys)}
    db.angle_features = r2r_data.get_all_point_angle_feature(A)
    db.scanvp_cands = {k: {"x": [3, 0, 0, 0]} for k in keys}
    return db, common


def main():
    db, common = reference_db()
    hist, ob = [], []
    for s in SAMPLES:
        n = len(s["path"])
        rel = [np.zeros(2, np.float32)] * n
        h = db.get_history_feature(SCAN, s["path"], s["views"], rel, s["t_cur"], return_img_probs=True)
        o = db.get_ob_pano_view(SCAN, s["path"], s["views"], [-1] * n, rel, s["t_cur"])
        hist.append(h)
        ob.append(o)
    T = lambda xs: [torch.from_numpy(np.ascontiguousarray(x)) for x in xs]
    rec = dict(hist_img_fts=common.pad_tensors(T([h[0] for h in hist])), hist_pano_img_fts=common.pad_tensors(T([h[2] for h in hist])),
               hist_pano_ang_fts=common.pad_tensors(T([h[3] for h in hist])), hist_img_probs=common.pad_tensors(T([h[4] for h in hist])),
               ob_img_fts=torch.stack(T([o[0] for o in ob])), ob_ang_fts=torch.stack(T([o[1] for o in ob])),
               angle_features=torch.from_numpy(np.stack(db.angle_features, 0)))
    out = os.path.join(ROOT, "tests", "golden", "feature_assembly.pt")
    torch.save(rec, out)
    print("wrote", out, {k: tuple(v.shape) for k, v in rec.items()})


if __name__ == "__main__":
    main()
