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
# timing of single image decodes through the QB3.h API: a BEST + RLE stream and an FTL stream of the same image
import time
from helpers import oracle, synth_tiles
img = synth_tiles(1, 512, 512, 3, np.uint8)[0]
for mode in (7, 8):
    s = oracle().encode(img, mode=mode)
    P.decode(s)
    t0 = time.perf_counter()
    for _ in range(3): d = P.decode(s)
    print("mode byte", s[10], "decode ms", (time.perf_counter() - t0) / 3 * 1e3, np.array_equal(d, img))
big = synth_tiles(1, 4096, 4096, 3, np.uint8)[0]
P.encode(big, mode=8)
t0 = time.perf_counter(); s = P.encode(big, mode=8); print("encode 4096x4096x3 through qb3_encode ms", (time.perf_counter() - t0) * 1e3, len(s))
