"""Encode timing for a few large tiles (one CTA per tile against tiles in parts)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import qb3_b200 as q
from bench import device_synth_tiles
dev = torch.device("cuda", 0)
MODE = int(sys.argv[1]) if len(sys.argv) > 1 else 8   # 8 FTL, 4 BASE, 7 BEST
for (w, h, b, dt, n) in [(4096, 4096, 3, 0, 19), (2048, 2048, 3, 0, 79), (4096, 4096, 1, 2, 29), (1024, 1024, 3, 0, 317), (4096, 4096, 3, 0, 1), (8192, 8192, 1, 0, 1)]:
    cfg = q.config(w, h, b, dt, mode=MODE)
    src = device_synth_tiles(n, w, h, b, dt, dev)
    slot = q.slot_bytes(cfg)
    dst = torch.empty((n, slot), dtype=torch.uint8, device=dev)
    sizes = torch.empty(n, dtype=torch.int64, device=dev); est = torch.empty(n, dtype=torch.int32, device=dev)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    best = 1e9
    for it in range(4):
        ev[0].record(); q.encode_batch(cfg, src, n, dst=dst, sizes=sizes, status=est); ev[1].record()
        torch.cuda.synchronize()
        if it: best = min(best, ev[0].elapsed_time(ev[1]))
    raw = src.numel()
    print("mode %d: " % MODE + "%dx%dx%d type %d, %d tiles: encode %.2f ms, %.0f GB/s, ratio %.3f" % (w, h, b, dt, n, best, raw / best / 1e6, sizes.sum().item() / raw), flush=True)
    del src, dst
    torch.cuda.empty_cache()
