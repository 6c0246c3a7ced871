import sys, os
sys.path.insert(0, "/root/repo")
import numpy as np
import rip_b200 as rip
rng = np.random.default_rng(1)
n, h, w = 16, 1080, 1920
img = rng.integers(0, 256, (n, h, w, 4), dtype=np.uint8)
d_in = rip.DeviceBuffer(img.nbytes).upload(img)
d_out = rip.DeviceBuffer(img.nbytes)
k = int(sys.argv[1]) if len(sys.argv) > 1 else 5
wt = rip.gauss_weights(k, 1.0 if k == 5 else 6.0)
for i in range(4):
    e0, e1 = rip.Event(), rip.Event()
    e0.record(); rip.gauss_dev(d_in.ptr, d_out.ptr, w, h, n, 4, k, wt); e1.record(); e1.sync()
    print(k, e0.elapsed_ns(e1) / 1e3, "us")
