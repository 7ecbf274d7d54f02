"""Step timings of the other BASELINE.json configurations (cfg 1 BM, cfg 4 MODE_HH 1080p D=256, cfg 5 4K D=256)."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from mvstereovision3_b200 import api, synth  # noqa: E402


def run(name, W, H, B, stage, setup, reps=3):
    l, r, _ = synth.stereogram(H, W, 0, 64, seed=0)
    L = np.stack([l] * B)
    R = np.stack([r] * B)
    with api.Engine(W, H, max_batch=B) as e:
        setup(e)
        i = e.info
        for _ in range(2):
            e.compute(L, R, stage)
            e.sync()
        e.profile_enable(True)
        ms = []
        for _ in range(reps):
            e.timer_start()
            e.compute(L, R, stage)
            ms.append(e.timer_stop())
        prof = e.profile_read()
        best = min(ms)
        cells = i.sgbm_W1 * H * i.sgbm_D if stage == api.STAGE_SGBM else (W - 80 + 1) * H * 80
        kern = sum(v[0] for v in prof.values()) / reps
        print("%-28s B=%-3d step %8.2f ms (kernels %.2f ms) -> %8.1f fps, %6.1f Gcells/s, td_cluster=%d" % (
            name, B, best, kern, B / best * 1e3, cells * B / kern / 1e6, i.sgbm_td_cluster))
        for k, v in sorted(prof.items(), key=lambda kv: -kv[1][0])[:5]:
            print("     %-16s %8.3f ms x%d" % (k, v[0] / reps, v[1] // reps))


cfg4 = dict(minDisp=0, numDisp=256, blockSize=5, disp12MaxDiff=1, preFilterCap=0, uniquenessRatio=10,
            speckleWindowSize=150, speckleRange=2, disparityMode=1, P1=200, P2=800)
cfg5 = dict(cfg4, disparityMode=0, uniquenessRatio=0, disp12MaxDiff=0, speckleWindowSize=0, speckleRange=0)
shipped = dict(minDisp=1, numDisp=128, blockSize=13, disp12MaxDiff=0, preFilterCap=0, uniquenessRatio=0,
               speckleWindowSize=150, speckleRange=2, disparityMode=0, P1=0, P2=0)
run("cfg1 BM bm.yml 752x480", 752, 480, 64, api.STAGE_BM,
    lambda e: e.set_bm_params(numDisp=80, blockSize=21, preFilterCap=2, textureThreshold=30, uniquenessRatio=0))
run("sgbm.yml as shipped D=128", 752, 480, 74, api.STAGE_SGBM, lambda e: e.set_sgbm_params(**shipped))
run("cfg4 HH 1920x1080 D=256", 1920, 1080, 8, api.STAGE_SGBM, lambda e: e.set_sgbm_params(**cfg4))
run("cfg5 SGBM 3840x2160 D=256", 3840, 2160, 4, api.STAGE_SGBM, lambda e: e.set_sgbm_params(**cfg5))
