"""Event timeline of the 2-engine end-to-end pipeline (cfg 2): when each step's upload, kernels and download run."""
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
from mvstereovision3_b200 import api, synth  # noqa: E402

P = dict(minDisp=1, numDisp=64, blockSize=13, speckleWindowSize=150, speckleRange=2)
H, W, B = 480, 752, 148
NL = int(sys.argv[1]) if len(sys.argv) > 1 else 2
ORDER = (sys.argv[2] != "0") if len(sys.argv) > 2 else True


def lane():
    e = api.Engine(W, H, max_batch=B)
    e.set_sgbm_params(**P)
    hl, hr, hd = api.pinned((B, H, W), np.uint8), api.pinned((B, H, W), np.uint8), api.pinned((B, H, W), np.int16)
    return e, hl, hr, hd, torch.cuda.ExternalStream(e.stream)


lanes = [lane() for _ in range(NL)]
gen = [synth.stereogram(H, W, 1, 64, seed=i)[:2] for i in range(8)]
for e, hl, hr, hd, _ in lanes:
    for i in range(B):
        hl.array[i], hr.array[i] = gen[i % 8]

ev = {}


def mark(k, name, s):
    e = torch.cuda.Event(enable_timing=True)
    e.record(s)
    ev[(k, name)] = e


def submit(k):
    e, a, b, _, s = lanes[k % NL]
    if k > 0 and ORDER:
        e.order_after(lanes[(k - 1) % NL][0])
    mark(k, "k0", s)
    e.compute(a.array, b.array, api.STAGE_SGBM)
    mark(k, "k1", s)


def collect(k):
    e, _, _, d, s = lanes[k % NL]
    e.download(B, out={"disp": d.array})
    mark(k, "d2h1", s)


n = 10
for rep in range(2):
    ev.clear()
    t0 = torch.cuda.Event(enable_timing=True)
    t0.record(lanes[0][4])
    submit(0)
    for k in range(1, n):
        submit(k)
        if k >= NL - 1:
            collect(k - NL + 1)
    for j in range(n - NL + 1, n):
        collect(j)
    torch.cuda.synchronize()
for k in range(n):
    g = lambda nm: t0.elapsed_time(ev[(k, nm)])
    print("step %d lane %d: submitted %7.2f  kernels done %7.2f  d2h done %7.2f" % (k, k % NL, g("k0"), g("k1"), g("d2h1")))
