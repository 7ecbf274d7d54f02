"""Single-pair latency through the host-facing call (pageable numpy images in, CV_16S map out), the reference's
one-pair-per-call use (src/disparity.cpp:6-10): median of 200 calls at cfg 2 and at configs/sgbm.yml as shipped."""
import sys
import time

import numpy as np

sys.path.insert(0, ".")
from mvstereovision3_b200 import api, synth  # noqa: E402

for name, D in (("cfg2 (D=64)", 64), ("sgbm.yml as shipped (D=128)", 128)):
    l, r, _ = synth.stereogram(480, 752, 1, D, seed=3)
    with api.Engine(752, 480) as e:
        e.set_sgbm_params(minDisp=1, numDisp=D, blockSize=13, speckleWindowSize=150, speckleRange=2)
        ts = []
        for i in range(220):
            t = time.perf_counter()
            e.compute(l, r, api.STAGE_SGBM)
            d = e.download(1)["disp"]
            ts.append(time.perf_counter() - t)
        ts = np.array(ts[20:]) * 1e3
        print("%-30s median %.3f ms  p95 %.3f ms  -> %.0f pairs/s" % (name, np.median(ts), np.percentile(ts, 95), 1e3 / np.median(ts)))
