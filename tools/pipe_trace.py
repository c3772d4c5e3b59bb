"""CUPTI timeline (torch.profiler) of one side by side encode + decode step of the host buffer pipeline; writes a
compact table of copies and kernels with start / end times to gpurun_out/pipe_trace.txt."""
import os, sys, time, threading, json
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import qb3_b200 as q
from bench import device_synth_tiles
from torch.profiler import profile, ProfilerActivity
dev = torch.device("cuda", 0)
n, w, h, b = 2048, 512, 512, 3
tb = w * h * b
cfg = q.config(w, h, b, 0, mode=8)
slot = q.slot_bytes(cfg)
chunk, depth = int(sys.argv[1]) if len(sys.argv) > 1 else 512, int(sys.argv[2]) if len(sys.argv) > 2 else 4
mode = sys.argv[3] if len(sys.argv) > 3 else "both"
h_src = torch.empty((n, tb), dtype=torch.uint8).pin_memory(); h_src.copy_(device_synth_tiles(n, w, h, b, 0, dev))
h_out = torch.empty((n, tb), dtype=torch.uint8).pin_memory()
h_packed = [torch.empty(n * slot, dtype=torch.uint8).pin_memory() for _ in range(2)]
h_off = [torch.zeros(n, dtype=torch.int64) for _ in range(2)]
h_sz = [torch.zeros(n, dtype=torch.int64) for _ in range(2)]
h_st = torch.zeros(n, dtype=torch.int32)
ep, dp = q.Pipe(cfg, int(os.environ.get("EC", chunk)), int(os.environ.get("ED", depth))), q.Pipe(cfg, chunk, depth)
def enc(k): ep.encode(h_src, n, h_packed[k % 2], h_off[k % 2], h_sz[k % 2])
def dec(k): dp.decode(h_packed[k % 2], h_off[k % 2], h_sz[k % 2], n, h_out, h_st)
def both(k):
    a = threading.Thread(target=enc, args=(k,)); bt = threading.Thread(target=dec, args=(k + 1,))
    a.start(); bt.start(); a.join(); bt.join()
enc(0); enc(1); dec(0); dec(1); both(0); both(1)
f = {"both": both, "enc": enc, "dec": dec}[mode]
import pynvml as nv
nv.nvmlInit(); hnd = nv.nvmlDeviceGetHandleByIndex(0)
samples, stop = [], [False]
def sampler():
    while not stop[0]:
        samples.append((time.perf_counter(), nv.nvmlDeviceGetClockInfo(hnd, nv.NVML_CLOCK_SM), nv.nvmlDeviceGetPowerUsage(hnd) // 1000))
th = threading.Thread(target=sampler); th.start()
time.sleep(0.05)
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    t0 = time.perf_counter(); f(0); t1 = time.perf_counter()
time.sleep(0.02)
stop[0] = True; th.join()
print("step ms", (t1 - t0) * 1e3)
print("clocks (ms since step start, MHz, W):", " ".join("%.1f:%d:%d" % ((t - t0) * 1e3, c, w) for t, c, w in samples[::max(1, len(samples) // 120)]))
path = os.path.join(ROOT, "gpurun_out", "pipe_trace_%s.json" % mode)
prof.export_chrome_trace(path)
ev = [e for e in json.load(open(path))["traceEvents"] if e.get("ph") == "X" and e.get("cat") in ("kernel", "gpu_memcpy")]
ev.sort(key=lambda e: e["ts"])
z = ev[0]["ts"]
with open(os.path.join(ROOT, "gpurun_out", "pipe_trace_%s.txt" % mode), "w") as o:
    for e in ev:
        a = e.get("args", {})
        o.write("%9.3f %9.3f %8.3f s%-4s %s %s\n" % ((e["ts"] - z) / 1e3, (e["ts"] + e["dur"] - z) / 1e3, e["dur"] / 1e3,
                a.get("stream", "?"), e["name"][:60], a.get("bytes", "")))
os.remove(path)
