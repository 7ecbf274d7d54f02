"""One warm-up + one timed SGBM step of the bench workload (cfg 2) at a small batch -- the command profiled by ncu.
usage: python tools/prof_step.py [batch] [cfg]   (cfg: cfg2 | shipped | cfg4)"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from mvstereovision3_b200 import api, synth  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
cfg = sys.argv[2] if len(sys.argv) > 2 else "cfg2"
if cfg == "cfg2":
    H, W = 480, 752
    p = dict(minDisp=1, numDisp=64, blockSize=13, disp12MaxDiff=0, preFilterCap=0, uniquenessRatio=0,
             speckleWindowSize=150, speckleRange=2, disparityMode=0, P1=0, P2=0)
elif cfg == "shipped":
    H, W = 480, 752
    p = dict(minDisp=1, numDisp=128, blockSize=13, disp12MaxDiff=0, preFilterCap=0, uniquenessRatio=0,
             speckleWindowSize=150, speckleRange=2, disparityMode=0, P1=0, P2=0)
else:
    H, W = 1080, 1920
    p = dict(minDisp=0, numDisp=256, blockSize=5, disp12MaxDiff=1, preFilterCap=0, uniquenessRatio=10,
             speckleWindowSize=150, speckleRange=2, disparityMode=1, P1=200, P2=800)
if os.environ.get("MVSV_W"):
    W = int(os.environ["MVSV_W"])
l, r, _ = synth.stereogram(H, W, p["minDisp"], p["numDisp"], seed=0)
L = np.stack([l] * B)
R = np.stack([r] * B)
with api.Engine(W, H, max_batch=B) as e:
    e.set_sgbm_params(**p)
    if os.environ.get("MVSV_DBG"):
        e.debug_set_flags(int(os.environ["MVSV_DBG"], 0))
        print("debug flags", os.environ["MVSV_DBG"], "td cluster", e.info.sgbm_td_cluster)
    for _ in range(2):
        e.compute(L, R, api.STAGE_SGBM)
        e.sync()
    e.profile_enable(True)
    e.timer_start()
    e.compute(L, R, api.STAGE_SGBM)
    ms = e.timer_stop()
    print("step %.3f ms for %d frames -> %.1f fps" % (ms, B, B / ms * 1e3))
    for k, v in sorted(e.profile_read().items(), key=lambda kv: -kv[1][0]):
        print("  %-16s %8.3f ms x%d" % (k, v[0], v[1]))
