#!/usr/bin/env python
"""Write tests/golden/ref_weights.json: the float32 bit patterns produced by the REFERENCE's own Gaussian weight
generator (oracle/_ref/librip_ref_weights.so = /root/reference/src/GaussianBlur/src/Controller.cpp compiled in place,
see oracle/Makefile) for the kernel sizes / sigmas the tests and the bench use.  Run in the build container (the GPU
box has no /root/reference); the JSON is the committed fixture that travels.

    python tools/make_ref_weights.py
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import oracle as O  # noqa: E402

CASES = [(5, 1.0), (5, 1.5), (17, 6.0), (3, 0.8), (7, 2.0), (9, 2.5), (31, 9.5), (1, 1.0)]

if __name__ == "__main__":
    if O.build_ref() is None:
        sys.exit("oracle/_ref could not be built: no reference tree")
    out = {"source": "Controller::_GenerateGausianKernel, /root/reference/src/GaussianBlur/src/Controller.cpp:342-362,395-417 "
                     "(g++ -O0 -std=c++17, image support CL_FALSE)", "cases": {}}
    for k, s in CASES:
        w = O.ref_gauss_weights(k, s)
        out["cases"][f"k{k}_s{s}"] = {"ksize": k, "sigma": s, "bits": [int(x) for x in w.ravel().view(np.uint32)]}
    p = os.path.join(ROOT, "tests", "golden", "ref_weights.json")
    with open(p, "w") as f:
        json.dump(out, f, indent=0, sort_keys=True)
    print(p, os.path.getsize(p), "bytes")
