"""Times Disparity::tm on the engine at the reference's frame size (752x480, kernelSize 5), host images in and out."""
import sys
import time

import numpy as np

sys.path.insert(0, ".")
from mvstereovision3_b200 import api, synth  # noqa: E402

l, r = synth.random_pair(480, 752, seed=1)
with api.Engine(752, 480) as e:
    e.tm(l, r, 5)
    e.profile_enable(True)
    t = time.perf_counter()
    n = 5
    for _ in range(n):
        e.tm(l, r, 5)
    dt = (time.perf_counter() - t) / n
    print("tm 752x480 k=5: %.2f ms per pair end to end; kernels: %s" % (dt * 1e3, e.profile_read()))
