"""Throughput sweep (BASELINE config 5): tile size x bands x type x mode at about 1 GB of raw pixels per
case per GPU, kernels only (inputs resident in HBM). Prints a markdown table. Every case is round trip checked over
all tiles, and the streams of its first and last tile are compared byte for byte with the oracle's (a bug that is
symmetric in both directions passes a round trip). Under torchrun every rank takes the same cases on its own shard of
the tile sequence (weak scaling) and rank 0 prints the times of the slowest rank.

  python tools/sweep.py [--gb 1.0] [--quick] > profiles/r02_sweep.md
"""
import argparse, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import torch
import qb3_b200 as q
from bench import device_synth_tiles
from helpers import DTYPES, oracle

ap = argparse.ArgumentParser()
ap.add_argument("--gb", type=float, default=1.0)
ap.add_argument("--quick", action="store_true")
args = ap.parse_args()
rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", "0"), ("WORLD_SIZE", "1"), ("LOCAL_RANK", "0")))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    import torch.distributed as dist
    dist.init_process_group("nccl", device_id=dev)


def say(*a):
    if rank == 0:
        print(*a, flush=True)


NAMES = ["u8", "i8", "u16", "i16", "u32", "i32", "u64", "i64"]
MODES = [(8, "FTL"), (4, "BASE"), (7, "BEST")]
cases = []
# tile size sweep on 3 band u8 and 1 band u16
for side in (64, 128, 256, 512, 1024, 2048, 4096):
    cases.append((side, side, 3, 0))
    if not args.quick:
        cases.append((side, side, 1, 2))
# band sweep on 256 x 256, the large counts on 512 x 512 as well
for b in (1, 4, 8, 16, 64, 256):
    cases.append((256, 256, b, 0))
    if not args.quick:
        cases.append((256, 256, b, 2))
for b in (64, 256):
    cases.append((512, 512, b, 0))
    if not args.quick:
        cases.append((512, 512, b, 2))
# type sweep on 512 x 512 x 1
for dt in range(8):
    cases.append((512, 512, 1, dt))
seen = set()
say("| tile | bands | type | mode | tiles per GPU | ratio | encode ms | decode ms | encode GB/s | decode GB/s | enc+dec GB/s | oracle bytes |")
say("|---|---|---|---|---|---|---|---|---|---|---|---|")
for w, h, b, dt in cases:
    for mode, mname in MODES:
        if (w, h, b, dt, mode) in seen:
            continue
        seen.add((w, h, b, dt, mode))
        if args.quick and mode == 7 and (w > 1024 or b > 16):
            continue
        ts = q.TYPESIZE[dt]
        tile = w * h * b * ts
        n = max(1, int(args.gb * 1e9 / tile))
        cfg = q.config(w, h, b, dt, mode=mode)
        src = device_synth_tiles(n, w, h, b, dt, dev, t0=rank * n)
        slot = q.slot_bytes(cfg)
        dst = torch.empty((n, slot), dtype=torch.uint8, device=dev)
        sizes = torch.empty(n, dtype=torch.int64, device=dev); est = torch.empty(n, dtype=torch.int32, device=dev)
        out = torch.empty((n, tile), dtype=torch.uint8, device=dev); dstat = torch.empty(n, dtype=torch.int32, device=dev)
        offsets = torch.arange(n, device=dev, dtype=torch.int64) * slot
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
        best_e = best_d = 1e9
        for it in range(4):
            ev[0].record()
            q.encode_batch(cfg, src, n, dst=dst, sizes=sizes, status=est)
            ev[1].record()
            q.decode_batch(cfg, dst, offsets, sizes, n, out=out, status=dstat)
            ev[2].record()
            torch.cuda.synchronize()
            if it:
                best_e, best_d = min(best_e, ev[0].elapsed_time(ev[1])), min(best_d, ev[1].elapsed_time(ev[2]))
        ok = torch.equal(out, src) and not est.any().item() and not dstat.any().item()
        # first and last tile of the shard against the oracle, byte for byte (large tiles: the first only)
        same = True
        if tile <= (64 << 20):
            for t in sorted({0, n - 1} if tile <= (8 << 20) else {0}):
                img = src[t].cpu().numpy().view(np.dtype(DTYPES[dt])).reshape(h, w, b)
                want = oracle().encode(img, mode=mode)
                got = dst[t, :int(sizes[t].item())].cpu().numpy().tobytes()
                same = same and got == want
            verdict = "same" if same else "DIFFERENT"
        else:
            verdict = "not compared (tile over 64 MB)"
        if world > 1:
            tt = torch.tensor([best_e, best_d, 0.0 if (ok and same) else 1.0], device=dev, dtype=torch.float64)
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            best_e, best_d = tt[0].item(), tt[1].item()
            if tt[2].item():
                ok = False
        raw = n * tile * world
        ratio = sizes.sum().item() / (n * tile)
        say("| %dx%d | %d | %s | %s | %d | %.3f | %.2f | %.2f | %.0f | %.0f | %.0f | %s |%s" % (
            w, h, b, NAMES[dt], mname, n, ratio, best_e, best_d, raw / best_e / 1e6, raw / best_d / 1e6,
            raw / (best_e + best_d) / 1e6, verdict, "" if ok else " ROUND TRIP OR PARITY FAILED ON SOME RANK"))
        del src, dst, out
        torch.cuda.empty_cache()
if world > 1:
    dist.destroy_process_group()
