import os, sys
import numpy as np
sys.path.insert(0, "/root/repo")
from mvstereovision3_b200 import api, synth
H, W, B = 1080, 1920, 8
p = dict(minDisp=0, numDisp=256, blockSize=5, disp12MaxDiff=1, preFilterCap=0, uniquenessRatio=10,
         speckleWindowSize=150, speckleRange=2, disparityMode=1, P1=200, P2=800)
l, r, _ = synth.stereogram(H, W, 0, 256, seed=0)
L = np.stack([l] * B); R = np.stack([r] * B)
ref = None
for flags in (0, 0x1000):
    with api.Engine(W, H, max_batch=B) as e:
        e.set_sgbm_params(**p)
        e.debug_set_flags(flags)
        for _ in range(2):
            e.compute(L, R, api.STAGE_SGBM); e.sync()
        e.profile_enable(True)
        e.compute(L, R, api.STAGE_SGBM)
        d = e.download(B)["disp"]
        prof = e.profile_read()
        print("flags", hex(flags), "cluster", e.info.sgbm_td_cluster, {k: round(v[0], 2) for k, v in prof.items() if k.startswith("sgbm")})
        if ref is None: ref = d
        else: print("same result:", bool(np.array_equal(ref, d)))
