"""Device-resident step time of one BASELINE configuration (no per-kernel events): python tools/time_step.py KEY [steps]"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import bench  # noqa: E402
from mvstereovision3_b200 import api  # noqa: E402

key = sys.argv[1] if len(sys.argv) > 1 else "cfg2"
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 20
spec = bench.CONFIGS[key]
p, H, W, B = spec["params"], spec["H"], spec["W"], spec["batch"]
gen, uniq = bench.make_inputs(spec, list(range(B)), B)
dl = torch.from_numpy(np.stack([gen[i % uniq][0] for i in range(B)])).cuda()
dr = torch.from_numpy(np.stack([gen[i % uniq][1] for i in range(B)])).cuda()
with api.Engine(W, H, max_batch=B) as e:
    if spec["kind"] == "bm":
        e.set_bm_params(**p); stages = api.STAGE_BM
    else:
        e.set_sgbm_params(**p); stages = api.STAGE_SGBM
    for _ in range(3):
        e.compute_device(dl.data_ptr(), W, dr.data_ptr(), W, H * W, B, stages)
    e.sync()
    e.timer_start()
    for _ in range(steps):
        e.compute_device(dl.data_ptr(), W, dr.data_ptr(), W, H * W, B, stages)
    ms = e.timer_stop() / steps
    print("%s: %.3f ms per step of %d frames -> %.1f frames/s" % (key, ms, B, B / ms * 1e3))
