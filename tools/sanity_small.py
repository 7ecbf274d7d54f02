"""Tiny invocation of every kernel of the path (for compute-sanitizer): SGBM 5/8 paths through the fused strip
sweep (one, two and four lanes per pixel, cluster and global-memory border hand-off) and through the independent passes, padded D, BM, rectification, XYZ, ROI means, min/max."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import cases  # noqa: E402
from mvstereovision3_b200 import api, synth  # noqa: E402

H, W = 40, 150
l, r = synth.random_pair(H, W, seed=1)
L, R = np.stack([l, r]), np.stack([r, l])
for mode, nd, flags in ((0, 64, 0), (1, 24, 0), (1, 16, 0x200), (0, 32, 0xff00), (0, 8, 0), (1, 128, 0x400), (0, 256, 0x900)):
    with api.Engine(W, H, max_batch=2) as e:
        e.set_sgbm_params(minDisp=1, numDisp=nd, blockSize=5, P1=8, P2=32, disp12MaxDiff=1, preFilterCap=31,
                          uniquenessRatio=10, speckleWindowSize=20, speckleRange=2, disparityMode=mode)
        e.debug_set_flags(flags)
        e.set_Q(cases.Q_REFERENCE)
        e.set_mean_rois(api.subimage_rois(W - 16, H, 16))
        e.compute(L, R, api.STAGE_SGBM | api.STAGE_XYZ | api.STAGE_MEANS)
        out = e.download(2, xyz=True, means=True)
        mm = e.download_minmax(2)
        print("sgbm mode", mode, "D", nd, "cluster", e.info.sgbm_td_cluster, "sum", int(out["disp"].astype(np.int64).sum()), mm.tolist())
with api.Engine(W, H, max_batch=2) as e:
    e.set_bm_params(numDisp=16, blockSize=9, preFilterCap=31, textureThreshold=10, uniquenessRatio=15)
    e.compute(L, R, api.STAGE_BM)
    print("bm sum", int(e.download(2)["disp"].astype(np.int64).sum()))
mx, my = cases.warp_maps(H, W, 0)
with api.Engine(W, H, max_batch=2) as e:
    e.upload_rectify_maps(0, mx, my, (3, 2, W - 7, H - 5))
    e.upload_rectify_maps(1, mx, my, (3, 2, W - 7, H - 5))
    e.set_sgbm_params(minDisp=0, numDisp=16, blockSize=3, disparityMode=0)
    e.compute(L, R, api.STAGE_RECTIFY | api.STAGE_SGBM)
    o = e.download(2, rect=True)
    print("rect sum", int(o["rectL"].astype(np.int64).sum()), "disp sum", int(o["disp"].astype(np.int64).sum()))
print("sanity ok")
