"""Where the time of a single image goes through the QB3.h API (the compatibility path)."""
import ctypes as C, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from helpers import PRODUCT_SO, QB3Lib, oracle, synth_tiles
P = QB3Lib(PRODUCT_SO, 256)
L = P.lib
img = synth_tiles(1, 512, 512, 3, np.uint8)[0]
s = oracle().encode(img, mode=8)
buf = np.frombuffer(s, dtype=np.uint8).copy()
out = np.zeros_like(img)
P.decode(s)
for rep in range(3):
    t = [time.perf_counter()]
    dims = (C.c_size_t * 3)()
    d = L.qb3_read_start(buf.ctypes.data, len(buf), dims); t.append(time.perf_counter())
    L.qb3_read_info(d); t.append(time.perf_counter())
    n = L.qb3_read_data(d, out.ctypes.data); t.append(time.perf_counter())
    n2 = L.qb3_read_data(d, out.ctypes.data); t.append(time.perf_counter())
    L.qb3_destroy_decoder(d); t.append(time.perf_counter())
    print("read_start %.2f  read_info %.2f  read_data %.2f  read_data again %.2f  destroy %.2f ms   ok %s" %
          tuple([1e3 * (t[i + 1] - t[i]) for i in range(5)] + [bool(n) and np.array_equal(out, img)]))
