"""Repeats one BASELINE configuration at its bench batch and checks that every step returns the same bits and that
frame 0 equals cv2 (bench.py's parity gate, run as a soak).  usage: python tools/determinism_check.py KEY [reps]"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from mvstereovision3_b200 import api  # noqa: E402

key = sys.argv[1]
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 10
spec = bench.CONFIGS[key]
p, H, W, B = spec["params"], spec["H"], spec["W"], spec["batch"]
gen, uniq = bench.make_inputs(spec, list(range(B)), B)
L = np.stack([gen[i % uniq][0] for i in range(B)])
R = np.stack([gen[i % uniq][1] for i in range(B)])
with api.Engine(W, H, max_batch=B) as e:
    if spec["kind"] == "bm":
        e.set_bm_params(**p); stages = api.STAGE_BM
    else:
        e.set_sgbm_params(**p); stages = api.STAGE_SGBM
    print(key, "strips per frame", e.info.sgbm_td_cluster)
    first, bad = None, 0
    for it in range(reps):
        e.compute(L, R, stages)
        d = e.download(B)["disp"]
        if first is None:
            first = d.copy()
            for b in range(uniq, B):
                if not np.array_equal(d[b], d[b % uniq]):
                    bad += 1
                    print("  step 0: frame %d differs from its twin %d in %d pixels" % (b, b % uniq, int((d[b] != d[b % uniq]).sum())))
        elif not np.array_equal(d, first):
            bad += 1
            fr = sorted(set(np.argwhere(d != first)[:, 0].tolist()))
            print("  step %d differs from step 0 in frames %s" % (it, fr[:10]))
    import cv2
    want = bench.cv_reference_frame(cv2, spec, gen[0][0], gen[0][1])
    print("  frame 0 == cv2:", bool(np.array_equal(first[0], want)), " mismatching steps/frames:", bad)
