import os, sys, hashlib
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from helpers import PRODUCT_SO, QB3Lib, golden_cases, golden_image
P = QB3Lib(PRODUCT_SO, 256)
for case in golden_cases():
    if case["name"] != "K8": continue
    s = bytes.fromhex(case["stream"])
    for i in range(3):
        d = P.decode(s)
        print("decode", None if d is None else (d.shape, int(d.astype(np.int64).sum()), np.argwhere(d != 0)[:5].tolist()), flush=True)
import ctypes
L = ctypes.CDLL(PRODUCT_SO)
print("last cuda error", L.qb3cu_last_cuda_error())
