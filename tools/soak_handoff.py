"""Soak test of the cluster hand-off: thousands of sweeps over small MODE_HH frames at every cluster size; every
result must equal the first one (and the oracle)."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import cases, oracle
from mvstereovision3_b200 import api, synth

p = cases.sgbm_params(minDisp=1, numDisp=32, blockSize=5, P1=20, P2=90, uniquenessRatio=5, disp12MaxDiff=1,
                      speckleWindowSize=30, speckleRange=2, mode=1)
gp = dict(p); gp["disparityMode"] = gp.pop("mode")
H, W, B = 37, 171, 24
ls, rs = zip(*[synth.random_pair(H, W, seed=s) for s in range(B)])
L, R = np.stack(ls), np.stack(rs)
want = np.stack([oracle.sgbm(ls[b], rs[b], p) for b in range(B)])
t0 = time.time()
total = 0
for nc in (1, 2, 4, 8):
    with api.Engine(W, H, max_batch=B) as e:
        e.set_sgbm_params(**gp)
        e.debug_set_flags(nc << 8)
        assert e.info.sgbm_td_cluster == nc
        for it in range(400):
            e.compute(L, R, api.STAGE_SGBM)
            got = e.download(B)["disp"]
            if not np.array_equal(got, want):
                raise SystemExit("MISMATCH nc=%d iteration %d: %d pixels" % (nc, it, int((got != want).sum())))
            total += 2 * B
    print("cluster %d: 400 x %d frames x 2 sweeps identical to the oracle" % (nc, B))
print("soak ok: %d cluster sweeps in %.1f s" % (total, time.time() - t0))
