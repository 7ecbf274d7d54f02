"""One warm-up + timed StereoBM steps at bm.yml (cfg 1) -- the command profiled by ncu.  usage: python tools/prof_bm.py [batch]"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from mvstereovision3_b200 import api, synth  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 148
l, r, _ = synth.stereogram(480, 752, 0, 80, seed=42)
L, R = np.stack([l] * B), np.stack([r] * B)
with api.Engine(752, 480, max_batch=B) as e:
    e.set_bm_params(numDisp=80, blockSize=21, preFilterCap=2, textureThreshold=30, uniquenessRatio=0)
    for _ in range(2):
        e.compute(L, R, api.STAGE_BM)
        e.sync()
    e.profile_enable(True)
    e.compute(L, R, api.STAGE_BM)
    e.sync()
    for k, v in sorted(e.profile_read().items(), key=lambda kv: -kv[1][0]):
        print("  %-16s %8.3f ms x%d" % (k, v[0], v[1]))
