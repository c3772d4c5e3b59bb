"""Repeats the host pipeline round trip on pageable and pinned buffers, new pipes every time, and says what differs."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import qb3_b200 as q
from helpers import oracle, synth_tiles
rng = np.random.default_rng(int(sys.argv[1]) if len(sys.argv) > 1 else 1)
n, w, h, b = 37, 64, 48, 3
tiles = synth_tiles(n, w, h, b, np.uint8)
want = [oracle().encode(tiles[t], mode=8) for t in range(n)]
cfg = q.config(w, h, b, q.U8, mode=8)
bad = 0
for it in range(int(sys.argv[2]) if len(sys.argv) > 2 else 150):
    chunk, depth = int(rng.integers(1, 12)), int(rng.integers(2, 5))
    pipe = q.Pipe(cfg, chunk, depth)
    packed = np.zeros(n * q.slot_bytes(cfg), np.uint8)
    offsets, sizes = np.zeros(n, np.uint64), np.zeros(n, np.uint64)
    pipe.encode(tiles, n, packed, offsets, sizes)
    o = 0
    for t in range(n):
        if int(offsets[t]) != o or int(sizes[t]) != len(want[t]) or packed[o:o + len(want[t])].tobytes() != want[t]:
            print("iter", it, "chunk", chunk, "depth", depth, "ENCODE tile", t, "offset", int(offsets[t]), "want", o, "size", int(sizes[t]), "want", len(want[t]),
                  "bytes equal", packed[int(offsets[t]):int(offsets[t]) + len(want[t])].tobytes() == want[t]); bad += 1; break
        o += (len(want[t]) + 15) // 16 * 16
    out = np.zeros_like(tiles); status = np.full(n, 99, np.uint32)
    pipe.decode(packed, offsets, sizes, n, out, status)
    if status.any() or not np.array_equal(out, tiles):
        wrong = [t for t in range(n) if not np.array_equal(out[t], tiles[t])]
        rows = sorted({int(r) for t in wrong[:3] for r in np.argwhere((out[t] != tiles[t]).any(axis=(1, 2)))[:, 0]})
        print("iter", it, "chunk", chunk, "depth", depth, "DECODE status", status.tolist()[:12], "wrong tiles", wrong[:12], "rows", rows[:16]); bad += 1
    pipe.close()
print("done, failures:", bad)
