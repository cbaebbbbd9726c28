import sys, os; sys.path.insert(0,'.')
import workloads, numpy as np
from decompress_b200 import _lib, lzo
eng=_lib.Engine(0)
txt=workloads.text(1<<20,3)
for n in (100, 5000, 70000, 1<<20):
    b=workloads.lzo1x_1_compress(txt[:n])
    r=lzo.uncompress_batch([b],[n],engine=eng)[0]
    print(n, len(b), r[:3], r[3]==txt[:n])
