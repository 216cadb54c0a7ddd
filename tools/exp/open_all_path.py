#!/usr/bin/env python3
"""Two kb_open_all_fk calls at d = 2^12 for a launch list under ncu (the second call is the steady state: hat_s cached)."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from keaki_b200 import _ffi  # noqa: E402

c = _ffi.Context(0)
one = np.zeros(8, np.uint32); one[0] = 5
d = int(os.environ.get("D", "4096"))
c.srs_generate(one, d, download=False)
co = np.random.default_rng(1).integers(0, 2**30, size=(d, 8), dtype=np.uint32)
co[:, 7] &= 0x0FFFFFFF
for i in range(2):
    c.open_all_fk(co)
    print("open_all ms", c.last_kernel_ms(0))
