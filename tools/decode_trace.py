"""CUPTI timeline (torch.profiler) of one qb3cu_decode_batch call; kernels with start / end to gpurun_out/decode_trace_<tag>.txt"""
import os, sys, json, time
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import qb3_b200 as q
from bench import device_synth_tiles, WORKLOADS, DEFAULT_TILES, QUANTA
from torch.profiler import profile, ProfilerActivity
wl, tag = sys.argv[1], sys.argv[2]
w, h, bands, dcode, dname, mode, cband, desc = WORKLOADS[wl]
n = int(sys.argv[3]) if len(sys.argv) > 3 else DEFAULT_TILES[wl]
dev = torch.device("cuda", 0)
cfg = q.config(w, h, bands, dcode, mode=mode, cband=cband, quanta=QUANTA.get(wl, 1))
src = device_synth_tiles(n, w, h, bands, dcode, dev)
dst, sizes, st = q.encode_batch(cfg, src, n)
offsets = torch.arange(n, device=dev, dtype=torch.int64) * dst.stride(0)
out, dstat = q.decode_batch(cfg, dst, offsets, sizes, n)
torch.cuda.synchronize()
for _ in range(2):
    q.decode_batch(cfg, dst, offsets, sizes, n, out=out, status=dstat)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    t0 = time.perf_counter()
    q.decode_batch(cfg, dst, offsets, sizes, n, out=out, status=dstat)
    torch.cuda.synchronize()
    print("decode ms", (time.perf_counter() - t0) * 1e3)
path = os.path.join(ROOT, "gpurun_out", "decode_trace_%s.json" % tag)
prof.export_chrome_trace(path)
ev = [e for e in json.load(open(path))["traceEvents"] if e.get("ph") == "X" and e.get("cat") in ("kernel", "gpu_memcpy", "gpu_memset")]
ev.sort(key=lambda e: e["ts"])
z = ev[0]["ts"]
with open(os.path.join(ROOT, "gpurun_out", "decode_trace_%s.txt" % tag), "w") as o:
    for e in ev:
        a = e.get("args", {})
        o.write("%9.3f %9.3f %8.3f s%-4s %s grid %s block %s smem %s\n" % ((e["ts"] - z) / 1e3, (e["ts"] + e["dur"] - z) / 1e3, e["dur"] / 1e3,
                a.get("stream", "?"), e["name"][:40], a.get("grid"), a.get("block"), a.get("shared memory")))
os.remove(path)
