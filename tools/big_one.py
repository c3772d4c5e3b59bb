"""One encode of a few large tiles, for a launch list: python tools/big_one.py MODE W H BANDS DTYPE NTILES"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import qb3_b200 as q
from bench import device_synth_tiles
mode, w, h, b, dt, n = (int(x) for x in sys.argv[1:7])
dev = torch.device("cuda", 0)
cfg = q.config(w, h, b, dt, mode=mode)
src = device_synth_tiles(n, w, h, b, dt, dev)
for it in range(2):
    dst, sizes, st = q.encode_batch(cfg, src, n)
    torch.cuda.synchronize()
print("ok", int(sizes.sum().item()))
