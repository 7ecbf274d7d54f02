"""Experiment: do kernels bound by different resources overlap when two engines (two streams) run half batches
concurrently?  Device-resident inputs, aggregate frames/s vs one engine with the full batch."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from mvstereovision3_b200 import api, synth

H, W = 480, 752
p = dict(minDisp=1, numDisp=64, blockSize=13, disp12MaxDiff=0, preFilterCap=0, uniquenessRatio=0,
         speckleWindowSize=150, speckleRange=2, disparityMode=0, P1=0, P2=0)
l, r, _ = synth.stereogram(H, W, 1, 64, seed=0)


def run(nengines, B, steps=20):
    dl = torch.from_numpy(np.stack([l] * B)).cuda()
    dr = torch.from_numpy(np.stack([r] * B)).cuda()
    engs = []
    for _ in range(nengines):
        e = api.Engine(W, H, max_batch=B)
        e.set_sgbm_params(**p)
        engs.append(e)
    def step():
        for e in engs:
            e.compute_device(dl.data_ptr(), W, dr.data_ptr(), W, H * W, B, api.STAGE_SGBM)
    for _ in range(3):
        step()
    for e in engs:
        e.sync()
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    for e in engs:
        e.sync()
    dt = time.perf_counter() - t0
    fps = nengines * B * steps / dt
    print("%d engine(s) x batch %3d: %8.1f frames/s" % (nengines, B, fps))
    for e in engs:
        e.close()


run(1, 148)
run(2, 74)
run(2, 148)
run(3, 74)
