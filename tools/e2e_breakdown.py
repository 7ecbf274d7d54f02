"""Where the end-to-end step (host buffers in, disparity maps out) spends its time at cfg 2, batch 148:
copies alone, kernels alone, serial sum, and the software pipeline with 2 and 3 engines."""
import sys
import time

import numpy as np

sys.path.insert(0, ".")
from mvstereovision3_b200 import api, synth  # noqa: E402

P = dict(minDisp=1, numDisp=64, blockSize=13, speckleWindowSize=150, speckleRange=2)
H, W, B = 480, 752, int(sys.argv[1]) if len(sys.argv) > 1 else 148
N = 12


def lane():
    e = api.Engine(W, H, max_batch=B)
    e.set_sgbm_params(**P)
    hl, hr, hd = api.pinned((B, H, W), np.uint8), api.pinned((B, H, W), np.uint8), api.pinned((B, H, W), np.int16)
    return e, hl, hr, hd


lanes = [lane() for _ in range(3)]
gen = [synth.stereogram(H, W, 1, 64, seed=i)[:2] for i in range(8)]
for e, hl, hr, hd in lanes:
    for i in range(B):
        hl.array[i], hr.array[i] = gen[i % 8]


def timed(fn, n=N):
    fn()
    for l in lanes:
        l[0].sync()
    t = time.perf_counter()
    for _ in range(n):
        fn()
    for l in lanes:
        l[0].sync()
    return (time.perf_counter() - t) / n * 1e3


e, hl, hr, hd = lanes[0]
print("upload only (stages=0)      %.2f ms" % timed(lambda: (e.compute(hl.array, hr.array, 0), e.sync())))
e.compute(hl.array, hr.array, api.STAGE_SGBM)
e.sync()
print("download only               %.2f ms" % timed(lambda: e.download(B, out={"disp": hd.array})))
print("upload + kernels            %.2f ms" % timed(lambda: (e.compute(hl.array, hr.array, api.STAGE_SGBM), e.sync())))
print("serial e2e                  %.2f ms" % timed(lambda: (e.compute(hl.array, hr.array, api.STAGE_SGBM), e.download(B, out={"disp": hd.array}))))


def pipe(nl, n):
    def go():
        k = 0
        lanes[0][0].compute(lanes[0][1].array, lanes[0][2].array, api.STAGE_SGBM)
        for k in range(1, n):
            l = lanes[k % nl]
            l[0].order_after(lanes[(k - 1) % nl][0])
            l[0].compute(l[1].array, l[2].array, api.STAGE_SGBM)
            if k >= nl - 1:
                m = lanes[(k - nl + 1) % nl]
                m[0].download(B, out={"disp": m[3].array})
        for j in range(max(n - nl + 1, 0), n):
            m = lanes[j % nl]
            m[0].download(B, out={"disp": m[3].array})
    return go


for nl in (2, 3):
    n = 24
    ms = timed(pipe(nl, n), n=2) / n
    print("pipeline with %d engines     %.2f ms/step  (%.0f frames/s)" % (nl, ms, B / ms * 1e3))
