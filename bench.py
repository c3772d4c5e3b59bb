#!/usr/bin/env python
"""bench.py -- QB3 encode/decode raw-pixel GB/s on B200 (BASELINE.json metric).

A step is one pass of the hot path over one batch of synthetic tiles resident in HBM: encode every tile
(qb3cu_encode_batch), then decode every stream (qb3cu_decode_batch). Workload at any N: BASELINE config 2,
4096 tiles of 512x512x3 u8 per GPU, QB3M_FTL lossless (weak scaling: tiles are independent, every rank gets
its own 4096, no collective on the data path). value = raw pixel bytes of all ranks / max-over-ranks step time.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--tiles T] [--workload c2|c3base|c3best]

--impl reference times the reference's own CPU codec (oracle/_ref/libQB3ref.so, compiled from /root/reference)
tile-parallel on all host threads, on a bounded sample of the same workload.
"""
import argparse
import ctypes as C
import json
import os
import sys
import threading
import time

# many CUDA streams are in flight at once (the host pipeline's chunks): more hardware queues than the default 8,
# or streams share queues and wait for each other. Read when the CUDA context is created.
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

WORKLOADS = {
    # name: (w, h, bands, dtype code, numpy dtype name, mode, cband, description)
    "c2": (512, 512, 3, 0, "uint8", 8, None, "4096 tiles 512x512x3 u8, QB3M_FTL lossless, encode+decode"),
    "c2best": (512, 512, 3, 0, "uint8", 7, None, "tiles 512x512x3 u8, QB3M_BEST, encode+decode"),
    "c3base": (512, 512, 8, 2, "uint16", 4, [0] * 8, "tiles 512x512x8 u16, core band 0, QB3M_BASE, encode+decode"),
    "c3best": (512, 512, 8, 2, "uint16", 7, [0] * 8, "tiles 512x512x8 u16, core band 0, QB3M_BEST, encode+decode"),
    "c4i32": (512, 512, 1, 5, "int32", 8, None, "tiles 512x512x1 i32, QB3M_FTL lossless, encode+decode"),
    "c4u64q3": (512, 512, 1, 6, "uint64", 4, None, "tiles 512x512x1 u64, QB3M_BASE quanta 3, encode+decode"),
}
QUANTA = {"c4u64q3": 3}
DEFAULT_TILES = {"c2": 4096, "c2best": 1024, "c3base": 1024, "c3best": 1024, "c4i32": 2048, "c4u64q3": 1024}


def ncu_traffic(wl, ntiles, kernel):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of a kernel from the committed ncu --set full capture of
    this command (profiles/r02_traffic.json names the commit it was taken at). Bytes per tile are what is recorded, so
    the figure follows the tile count; None when nothing was captured for the workload / kernel."""
    try:
        t = json.load(open(os.path.join(ROOT, "profiles", "r02_traffic.json")))
        per_tile = t["bytes_per_tile"][wl][kernel]
        return {"bytes": per_tile * ntiles, "source": "profiles/r02_traffic.json", "captured_at_commit": t["commit"]}
    except (OSError, KeyError, ValueError):
        return None


def device_synth_tiles(ntiles, w, h, bands, dtype_code, device, t0=0, seed=12345, chunk=64):
    """BASELINE.md section 3 generator on the device, bit-identical to tests/helpers.synth_tiles.
    Returns a uint8 tensor [ntiles, tile_bytes] (little endian values)."""
    import torch
    ts = (1, 1, 2, 2, 4, 4, 8, 8)[dtype_code]
    bits = 8 * ts
    nb = {1: 3, 2: 6, 4: 8, 8: 10}[ts]
    A = (1 << 40) if bits == 64 else (1 << (bits - 1)) - 1
    out = torch.empty((ntiles, w * h * bands * ts), dtype=torch.uint8, device=device)
    i64 = torch.int64

    def lsr(x, k):  # logical shift right on int64
        return (x >> k) & ((1 << (64 - k)) - 1)

    def c64(v):  # python int -> wrapped int64 constant
        v &= (1 << 64) - 1
        return v - (1 << 64) if v >= (1 << 63) else v

    def tri(u, P, a):
        m = u % P
        return torch.minimum(m, P - m) * a // (P // 2)

    y = torch.arange(h, device=device, dtype=i64)[None, :, None, None]
    x = torch.arange(w, device=device, dtype=i64)[None, None, :, None]
    c = torch.arange(bands, device=device, dtype=i64)[None, None, None, :]
    for s in range(0, ntiles, chunk):
        n = min(chunk, ntiles - s)
        t = torch.arange(t0 + s, t0 + s + n, device=device, dtype=i64)[:, None, None, None]
        v = tri(x + 37 * t, 211, A // 2) + tri(y + 91 * t, 157, A // 2) + (c * A) // (8 * bands)
        idx = ((t * h + y) * w + x) * bands + c
        z = (idx ^ seed) + c64(0x9E3779B97F4A7C15)
        z = (z ^ lsr(z, 30)) * c64(0xBF58476D1CE4E5B9)
        z = (z ^ lsr(z, 27)) * c64(0x94D049BB133111EB)
        z = z ^ lsr(z, 31)
        v = v + (z & ((1 << nb) - 1))
        if ts == 1:
            out[s:s + n] = (v & 0xFF).to(torch.uint8).reshape(n, -1)
        elif ts == 8:
            out[s:s + n] = v.contiguous().view(torch.uint8).reshape(n, -1)
        else:  # wrap into the signed torch type of the same width, then reinterpret the bytes
            lo = v & ((1 << bits) - 1)
            lo = torch.where(lo >= (1 << (bits - 1)), lo - (1 << bits), lo)
            out[s:s + n] = lo.to(torch.int16 if ts == 2 else torch.int32).contiguous().view(torch.uint8).reshape(n, -1)
        del v, idx, z
    return out


class ClockSampler(threading.Thread):
    """Samples SM clock and throttle reasons of one GPU through NVML while the timed region runs."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.stop_flag, self.max_mhz = index, [], set(), False, None

    def run(self):
        try:
            import pynvml as nv
            nv.nvmlInit()
            h = nv.nvmlDeviceGetHandleByIndex(self.index)
            self.max_mhz = nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM)
            names = {nv.nvmlClocksThrottleReasonHwSlowdown: "hw_slowdown",
                     nv.nvmlClocksThrottleReasonHwThermalSlowdown: "hw_thermal_slowdown",
                     nv.nvmlClocksThrottleReasonSwThermalSlowdown: "sw_thermal_slowdown",
                     nv.nvmlClocksThrottleReasonSwPowerCap: "sw_power_cap"}
            while not self.stop_flag:
                self.samples.append(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                for bit, name in names.items():
                    if r & bit:
                        self.reasons.add(name)
                time.sleep(0.05)
        except Exception as e:  # NVML missing: report it instead of inventing numbers
            self.reasons.add("nvml_unavailable:%s" % type(e).__name__)

    def result(self):
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2] if s else None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(s)}


def bind_to_gpu_cores(index):
    """Several ranks on one host: keep this rank's threads, and with them the page locked buffers it allocates (first
    touch), on the CPU cores next to its GPU; traffic that crosses sockets halves the end to end rate. Best effort."""
    try:
        import pynvml as nv
        nv.nvmlInit()
        h = nv.nvmlDeviceGetHandleByIndex(index)
        ncpu = os.cpu_count() or 1
        words = nv.nvmlDeviceGetCpuAffinity(h, (ncpu + 63) // 64)
        cores = {64 * i + b for i, w in enumerate(words) for b in range(64) if (w >> b) & 1}
        cores &= set(os.sched_getaffinity(0))
        if cores:
            os.sched_setaffinity(0, cores)
            return len(cores)
    except Exception:  # noqa: BLE001 -- NVML or the affinity call missing: run unbound
        pass
    return None


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return json.load(open(p))["hbm_gbs"], "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def run_reference_arm(args, wl):
    """The reference CPU codec, tile-parallel on the host cores (rank 0 only)."""
    import numpy as np
    from helpers import REF_SO, REFBENCH_SO, synth_tiles
    w, h, bands, dcode, dname, mode, cband, desc = WORKLOADS[wl]
    if int(os.environ.get("RANK", "0")) != 0:
        return
    if not (os.path.exists(REF_SO) and os.path.exists(REFBENCH_SO)):
        print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref was not built in the build container"}))
        return
    ncores = len(os.sched_getaffinity(0))
    ntiles = args.ref_tiles or args.tiles or DEFAULT_TILES[wl]
    # the same tile sequence as the GPU arm's rank 0. The generator is integer only and bit identical on both sides
    # (tests/test_host_logic.py): on a box with a GPU it runs there and the tiles are copied to host memory before
    # anything is timed -- the CPU codec then runs on host memory only; without one, numpy makes the first 512 tiles
    # and the batch repeats them (numpy needs minutes for 4096).
    tiles = None
    try:
        import torch
        if torch.cuda.is_available():
            t = device_synth_tiles(ntiles, w, h, bands, dcode, torch.device("cuda", 0))
            tiles = t.cpu().numpy().view(np.dtype(dname)).reshape(ntiles, h, w, bands)
            del t
            torch.cuda.empty_cache()
    except Exception:  # noqa: BLE001
        tiles = None
    distinct = ntiles
    if tiles is None:
        distinct = min(ntiles, 512)
        first = synth_tiles(distinct, w, h, bands, np.dtype(dname))
        tiles = np.concatenate([first] * ((ntiles + distinct - 1) // distinct))[:ntiles]
    tile_bytes = tiles[0].nbytes
    bench = C.CDLL(REFBENCH_SO)
    bench.refbench_run.restype = C.c_int
    bench.refbench_run.argtypes = [C.c_char_p, C.c_size_t, C.c_size_t, C.c_size_t, C.c_size_t, C.c_int, C.c_int, C.c_void_p,
                                   C.c_uint64, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p, C.c_int, C.c_int,
                                   C.POINTER(C.c_double), C.POINTER(C.c_double)]
    slot = 1024 + int(tile_bytes * 1.14) + 64
    streams = np.zeros((ntiles, slot), np.uint8)
    sizes = np.zeros(ntiles, np.uint64)
    cb = (C.c_size_t * 256)(*cband) if cband else None
    times = []
    for i in range(args.warmup + args.steps):
        e, d = C.c_double(), C.c_double()
        rc = bench.refbench_run(REF_SO.encode(), ntiles, w, h, bands, dcode, mode, cb, QUANTA.get(wl, 1), tiles.ctypes.data,
                                streams.ctypes.data, slot, sizes.ctypes.data, None, ncores, 1, C.byref(e), C.byref(d))
        if rc:
            print(json.dumps({"impl": "reference", "unavailable": "refbench_run failed rc=%d" % rc}))
            return
        if i >= args.warmup:
            times.append((e.value, d.value))
    te = sum(t[0] for t in times) / len(times)
    td = sum(t[1] for t in times) / len(times)
    raw = ntiles * tile_bytes
    val = raw / (te + td) / 1e9
    sample = "%d tiles per step (%d distinct), one handle per tile, std::thread pool over tiles" % (ntiles, distinct)
    print(json.dumps({
        "impl": "reference", "metric": "QB3 encode+decode raw-pixel GB/s", "value": val, "unit": "GB/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * (te + td), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "u8" if dcode < 2 else dname, "data": "synthetic",
        "config": {"workload": desc, "tiles_per_gpu": ntiles, "tile": [w, h, bands], "mode": mode,
                   "l2": "inputs (%.2f GB per GPU) larger than L2, no flush needed" % (raw / 1e9),
                   "sharding": "contiguous tile ranges per rank, no collective"},
        "encode_gbs": raw / te / 1e9, "decode_gbs": raw / td / 1e9,
        "compressed_ratio": float(sizes.sum()) / raw,
        "cpu_baseline": {"value": val, "unit": "GB/s", "cores": ncores, "kind": "reference", "sample": sample},
        "e2e": {"value": val, "unit": "GB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))


def cpu_baseline(wl, ntiles=256):
    """Bounded CPU sample for the main arm's cpu_baseline object (rank 0, N=1)."""
    import numpy as np
    from helpers import REF_SO, REFBENCH_SO, synth_tiles
    w, h, bands, dcode, dname, mode, cband, desc = WORKLOADS[wl]
    if not (os.path.exists(REF_SO) and os.path.exists(REFBENCH_SO)):
        return None
    ncores = len(os.sched_getaffinity(0))
    tiles = synth_tiles(ntiles, w, h, bands, np.dtype(dname))
    tile_bytes = tiles[0].nbytes
    bench = C.CDLL(REFBENCH_SO)
    bench.refbench_run.restype = C.c_int
    bench.refbench_run.argtypes = [C.c_char_p, C.c_size_t, C.c_size_t, C.c_size_t, C.c_size_t, C.c_int, C.c_int, C.c_void_p,
                                   C.c_uint64, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p, C.c_int, C.c_int,
                                   C.POINTER(C.c_double), C.POINTER(C.c_double)]
    slot = 1024 + int(tile_bytes * 1.14) + 64
    streams = np.zeros((ntiles, slot), np.uint8)
    sizes = np.zeros(ntiles, np.uint64)
    cb = (C.c_size_t * 256)(*cband) if cband else None
    e, d = C.c_double(), C.c_double()
    reps = 3
    rc = bench.refbench_run(REF_SO.encode(), ntiles, w, h, bands, dcode, mode, cb, QUANTA.get(wl, 1), tiles.ctypes.data, streams.ctypes.data,
                            slot, sizes.ctypes.data, None, ncores, reps, C.byref(e), C.byref(d))
    if rc:
        return None
    raw = ntiles * tile_bytes
    return {"value": raw / (e.value + d.value) / 1e9, "unit": "GB/s", "cores": ncores, "kind": "reference",
            "encode_gbs": raw / e.value / 1e9, "decode_gbs": raw / d.value / 1e9,
            "sample": "%d tiles of the workload, best of %d, one handle per tile, thread pool over tiles" % (ntiles, reps)}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--workload", default="c2", choices=sorted(WORKLOADS))
    ap.add_argument("--tiles", type=int, default=0, help="tiles per GPU (default: the workload's)")
    ap.add_argument("--ref-tiles", type=int, default=0, help="tiles per step of the reference arm (0: the workload's batch)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--e2e-tiles", type=int, default=0, help="tiles per end-to-end step (0: the resident leg's batch when host memory allows, else 2048)")
    ap.add_argument("--no-others", action="store_true", help="skip the short runs of the other BASELINE configs")
    ap.add_argument("--e2e-enc-chunk", type=int, default=256, help="tiles per chunk of the encode pipe")
    ap.add_argument("--e2e-enc-depth", type=int, default=8, help="chunks in flight in the encode pipe")
    ap.add_argument("--e2e-dec-chunk", type=int, default=256, help="tiles per chunk of the decode pipe")
    ap.add_argument("--e2e-dec-depth", type=int, default=8, help="chunks in flight in the decode pipe")
    args = ap.parse_args()
    wl = args.workload
    if args.impl == "reference":
        return run_reference_arm(args, wl)

    import torch
    import qb3_b200 as q
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    assert torch.cuda.is_available(), "bench.py needs a CUDA device: there is no CPU path"
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    numa = bind_to_gpu_cores(local) if world > 1 else None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def resident(wname, ntiles, steps, warmup, keep=False):
        """encode + decode of one workload with inputs resident in HBM: CUDA events around the steps and around each
        pass, max over ranks. Every rank owns its own contiguous shard of the tile sequence (weak scaling, no exchange)."""
        w, h, bands, dcode, dname, mode, cband, desc = WORKLOADS[wname]
        ts = q.TYPESIZE[dcode]
        tile_bytes = w * h * bands * ts
        cfg = q.config(w, h, bands, dcode, mode=mode, cband=cband, quanta=QUANTA.get(wname, 1))
        slot = q.slot_bytes(cfg)
        t_lo, t_hi = q.shard_range(ntiles * world, rank, world)
        src = device_synth_tiles(t_hi - t_lo, w, h, bands, dcode, dev, t0=t_lo)
        dst = torch.empty((ntiles, slot), dtype=torch.uint8, device=dev)
        sizes = torch.empty(ntiles, dtype=torch.int64, device=dev)
        est = torch.empty(ntiles, dtype=torch.int32, device=dev)
        dstat = torch.empty(ntiles, dtype=torch.int32, device=dev)
        out = torch.empty((ntiles, tile_bytes), dtype=torch.uint8, device=dev)
        offsets = torch.arange(ntiles, device=dev, dtype=torch.int64) * slot

        def step(events=None):
            if events:
                events[0].record()
            q.encode_batch(cfg, src, ntiles, dst=dst, sizes=sizes, status=est)
            if events:
                events[1].record()
            q.decode_batch(cfg, dst, offsets, sizes, ntiles, out=out, status=dstat)
            if events:
                events[2].record()

        for _ in range(max(warmup, 3)):
            step()
        barrier()
        assert not est.any().item() and not dstat.any().item(), "tile status reports an error"
        assert wname in QUANTA or torch.equal(out, src), "decode(encode(x)) != x"
        comp_bytes = int(sizes.sum().item())
        evs = [[torch.cuda.Event(enable_timing=True) for _ in range(3)] for _ in range(steps)]
        launches0 = q.kernel_launches()
        barrier()
        t_start, t_end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t_start.record()
        for i in range(steps):
            step(evs[i])
        t_end.record()
        barrier()
        launches = q.kernel_launches() - launches0
        total_ms = t_start.elapsed_time(t_end)
        enc_ms = sum(e[0].elapsed_time(e[1]) for e in evs) / steps
        dec_ms = sum(e[1].elapsed_time(e[2]) for e in evs) / steps
        comp_all = comp_bytes
        if world > 1:
            t = torch.tensor([total_ms, enc_ms, dec_ms], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            total_ms, enc_ms, dec_ms = t.tolist()
            cb = torch.tensor([comp_bytes], device=dev, dtype=torch.int64)
            dist.all_reduce(cb)
            comp_all = int(cb.item())
        r = {"cfg": cfg, "slot": slot, "tile_bytes": tile_bytes, "ntiles": ntiles, "ts": ts, "desc": desc, "dname": dname,
             "geom": [w, h, bands], "mode": mode, "ms_per_step": total_ms / steps, "enc_ms": enc_ms, "dec_ms": dec_ms,
             "comp_rank": comp_bytes, "comp_all": comp_all, "launches": launches, "raw_rank": ntiles * tile_bytes}
        if keep:
            r["src"], r["step"] = src, step
        return r

    ntiles = args.tiles or DEFAULT_TILES[wl]
    sampler = ClockSampler(local)
    sampler.start()
    R = resident(wl, ntiles, args.steps, args.warmup, keep=True)
    sampler.stop_flag = True
    sampler.join()
    w, h, bands = R["geom"]
    cfg, slot, tile_bytes, ts, desc, dname, mode = R["cfg"], R["slot"], R["tile_bytes"], R["ts"], R["desc"], R["dname"], R["mode"]
    src, step = R["src"], R["step"]
    ms_per_step, enc_ms, dec_ms, launches = R["ms_per_step"], R["enc_ms"], R["dec_ms"], R["launches"]
    comp_bytes, comp_all = R["comp_rank"], R["comp_all"]
    raw_rank = R["raw_rank"]
    raw_all = raw_rank * world
    value = raw_all / (ms_per_step * 1e-3) / 1e9

    # End to end through the C ABI with HOST buffers (include/qb3cu.h, qb3cu_pipe_*): pinned host pixels ->
    # qb3cu_pipe_encode -> packed streams + index in host memory -> qb3cu_pipe_decode -> host pixels. Every copy in
    # either direction is inside the calls, hence inside the timed region. A step encodes one batch and decodes one
    # batch: the encode of batch k runs on one host thread while the decode of batch k-1 (the streams the previous
    # step produced, read from host memory) runs on another, the way a service that both ingests and serves tiles
    # would use the library; PCIe then carries pixels up and pixels down at the same time. The same two calls made
    # one after the other are timed as well ("sequential").
    e2e = None
    if not args.no_e2e:
        try:
            # the same batch as the resident leg when the host has the memory for its page locked copies (pixels in,
            # pixels out, three packed buffers), else the number asked for
            need = ntiles * (2 * tile_bytes + 3 * tile_bytes)
            try:
                import psutil
                avail = psutil.virtual_memory().available
            except Exception:  # noqa: BLE001
                avail = 0
            n2 = ntiles if (args.e2e_tiles == 0 and avail > 2.5 * need * max(1, world)) else min(ntiles, args.e2e_tiles or 2048)
            h_src = torch.empty((n2, tile_bytes), dtype=torch.uint8).pin_memory()
            h_src.copy_(src[:n2])
            h_out = torch.zeros((n2, tile_bytes), dtype=torch.uint8).pin_memory()
            NB = 3  # packed stream buffers in rotation between the encoding and the decoding thread
            # a stream is never longer than its raw tile plus headers (stored fallback), so raw size + 1 KB per tile holds it
            h_packed = [torch.empty((n2 * (tile_bytes + 1024),), dtype=torch.uint8).pin_memory() for _ in range(NB)]
            h_off = [torch.zeros(n2, dtype=torch.int64) for _ in range(NB)]
            h_sz = [torch.zeros(n2, dtype=torch.int64) for _ in range(NB)]
            h_stat = torch.zeros(n2, dtype=torch.int32)
            totals = [0] * NB
            enc_pipe = q.Pipe(cfg, args.e2e_enc_chunk, args.e2e_enc_depth)
            dec_pipe = q.Pipe(cfg, args.e2e_dec_chunk, args.e2e_dec_depth)

            def enc(k):
                totals[k % NB] = enc_pipe.encode(h_src, n2, h_packed[k % NB], h_off[k % NB], h_sz[k % NB])

            def dec(k):
                dec_pipe.decode(h_packed[k % NB], h_off[k % NB], h_sz[k % NB], n2, h_out, h_stat)

            def streaming(k0, reps):
                """Batches k0 .. k0+reps-1 are encoded by one thread while another decodes batches k0-1 .. k0+reps-2, each as
                soon as its streams are in host memory (batch k0-1 was encoded before the clock started): reps encodes and
                reps decodes, nothing waits at a step boundary."""
                encoded = [threading.Semaphore(0) for _ in range(reps + 1)]   # encoded[i]: batch k0-1+i is in host memory
                decoded = [threading.Semaphore(0) for _ in range(reps + 1)]   # decoded[i]: batch k0-1+i has been read back
                encoded[0].release()
                errors = []

                def enc_loop():
                    try:
                        for i in range(1, reps + 1):
                            if i - NB >= 0:
                                decoded[i - NB].acquire()      # the buffer this batch goes into has been decoded
                            enc(k0 - 1 + i)
                            encoded[i].release()
                    except Exception as e:  # noqa: BLE001 -- reported by the caller
                        errors.append(e)
                        for sem in encoded:
                            sem.release()

                def dec_loop():
                    try:
                        for i in range(reps):
                            encoded[i].acquire()
                            dec(k0 - 1 + i)
                            decoded[i].release()
                    except Exception as e:  # noqa: BLE001
                        errors.append(e)
                        for sem in decoded:
                            sem.release()

                ta, tb = threading.Thread(target=enc_loop), threading.Thread(target=dec_loop)
                ta.start(); tb.start(); ta.join(); tb.join()
                if errors:
                    raise errors[0]

            def timed(fn, reps, k0):
                barrier()
                t0 = time.perf_counter()
                for k in range(k0, k0 + reps):
                    fn(k)
                barrier()
                t = (time.perf_counter() - t0) / reps
                if world > 1:
                    tt = torch.tensor([t], device=dev, dtype=torch.float64)
                    dist.all_reduce(tt, op=dist.ReduceOp.MAX)
                    t = tt.item()
                return t

            launches_e0 = q.kernel_launches()
            enc(0); dec(0); enc(1); dec(1)         # warm-up: buffers of both pipes reach their final size
            streaming(2, 2)
            k2 = max(3, args.steps // 2)
            t_seq = timed(lambda k: (enc(k), dec(k)), k2, 4)
            h_out.zero_()
            k0 = 4 + k2
            barrier()
            t0 = time.perf_counter()
            streaming(k0, 2 * k2)                  # batch k0 - 1, the last one of the sequential run, is its first input
            barrier()
            t_e2e = (time.perf_counter() - t0) / (2 * k2)
            if world > 1:
                tt = torch.tensor([t_e2e], device=dev, dtype=torch.float64)
                dist.all_reduce(tt, op=dist.ReduceOp.MAX)
                t_e2e = tt.item()
            assert wl in QUANTA or torch.equal(h_out, h_src), "end to end round trip differs"
            assert not h_stat.any().item(), "tile status reports an error"
            index_bytes = 2 * 8 * n2
            h2d_b = n2 * tile_bytes + totals[0] + index_bytes
            d2h_b = totals[0] + index_bytes + n2 * tile_bytes + 4 * n2
            e2e = {"value": n2 * tile_bytes * world / t_e2e / 1e9, "unit": "GB/s",
                   "h2d_bytes_per_step": int(h2d_b), "d2h_bytes_per_step": int(d2h_b),
                   "tiles_per_step": n2, "ms_per_step": 1e3 * t_e2e,
                   "sequential": {"value": n2 * tile_bytes * world / t_seq / 1e9, "ms_per_step": 1e3 * t_seq},
                   "note": "qb3cu_pipe_encode + qb3cu_pipe_decode on pinned host buffers: one host thread encodes batch after "
                           "batch, a second one decodes each batch's streams from host memory as soon as they are there; a "
                           "step = one batch encoded and one decoded; 'sequential' is the two calls one after the other on "
                           "one thread",
                   "pipes": {"encode": [args.e2e_enc_chunk, args.e2e_enc_depth], "decode": [args.e2e_dec_chunk, args.e2e_dec_depth]}}
            enc_pipe.close(); dec_pipe.close()
            # The ceiling the host link sets: the same page locked buffers, the same bytes per step in each direction
            # (pixels + streams + index up, streams + index + pixels down), the pipes' chunk sizes, as plain
            # cudaMemcpyAsync calls on two streams and nothing else -- one call per copy.
            try:
                d_up = torch.empty(args.e2e_enc_chunk * tile_bytes, dtype=torch.uint8, device=dev)
                d_dn = torch.empty(args.e2e_dec_chunk * tile_bytes, dtype=torch.uint8, device=dev)
                s_up, s_dn = torch.cuda.Stream(dev), torch.cuda.Stream(dev)
                flat_src, flat_out, flat_pk = h_src.view(-1), h_out.view(-1), h_packed[0]

                def copies():
                    with torch.cuda.stream(s_up):
                        for o in range(0, n2 * tile_bytes, d_up.numel()):
                            k = min(d_up.numel(), n2 * tile_bytes - o)
                            d_up[:k].copy_(flat_src[o:o + k], non_blocking=True)
                        for o in range(0, totals[0], d_up.numel()):
                            k = min(d_up.numel(), totals[0] - o)
                            d_up[:k].copy_(flat_pk[o:o + k], non_blocking=True)
                    with torch.cuda.stream(s_dn):
                        for o in range(0, totals[0], d_dn.numel()):
                            k = min(d_dn.numel(), totals[0] - o)
                            flat_pk[o:o + k].copy_(d_dn[:k], non_blocking=True)
                        for o in range(0, n2 * tile_bytes, d_dn.numel()):
                            k = min(d_dn.numel(), n2 * tile_bytes - o)
                            flat_out[o:o + k].copy_(d_dn[:k], non_blocking=True)

                copies()
                reps_c = 3
                barrier()
                t0 = time.perf_counter()
                for _ in range(reps_c):
                    copies()
                barrier()
                t_copy = (time.perf_counter() - t0) / reps_c
                if world > 1:
                    tt = torch.tensor([t_copy], device=dev, dtype=torch.float64)
                    dist.all_reduce(tt, op=dist.ReduceOp.MAX)
                    t_copy = tt.item()
                e2e["copy_ceiling_gbs"] = n2 * tile_bytes * world / t_copy / 1e9
                e2e["copy_ceiling_ms"] = 1e3 * t_copy
                e2e["of_copy_ceiling"] = e2e["value"] / e2e["copy_ceiling_gbs"]
                del d_up, d_dn
            except Exception as exc:  # noqa: BLE001 -- the ceiling is commentary on the number above, not part of it
                e2e["copy_ceiling_error"] = "%s: %s" % (type(exc).__name__, exc)
            del h_src, h_out, h_packed
            # restore the device state for anything that follows
            step()
            barrier()
        except AssertionError:  # a wrong result is never downgraded to a note
            raise
        except Exception as exc:  # noqa: BLE001 -- the kernel-only numbers above stand on their own; say what went wrong
            e2e = {"error": "%s: %s" % (type(exc).__name__, exc)}
            barrier()

    # The other BASELINE configs, short runs with inputs resident in HBM (every rank its own shard, as for the headline):
    # config 3 (8 band u16, core band 0, BASE and BEST) and config 4 (1 band i32 lossless, u64 quanta 3).
    peak, peak_src = measured_peak()
    others = []
    if wl == "c2" and not args.no_others:
        del src, step, R
        torch.cuda.empty_cache()
        for wname in ("c3base", "c3best", "c4i32", "c4u64q3"):
            try:
                r = resident(wname, DEFAULT_TILES[wname], 3, 3)
                alg = r["raw_rank"] + r["comp_rank"]
                others.append({"workload": r["desc"], "name": wname, "tiles_per_gpu": r["ntiles"], "dtype": r["dname"],
                               "value": r["raw_rank"] * world / (r["ms_per_step"] * 1e-3) / 1e9, "unit": "GB/s",
                               "ms_per_step": r["ms_per_step"], "encode_ms": r["enc_ms"], "decode_ms": r["dec_ms"],
                               "compressed_ratio": r["comp_all"] / (r["raw_rank"] * world),
                               "roofline": {"bound": "hbm", "peak": peak, "unit": "GB/s",
                                            "encode": {"achieved": alg / (r["enc_ms"] * 1e-3) / 1e9, "frac": alg / (r["enc_ms"] * 1e-3) / 1e9 / peak},
                                            "decode": {"achieved": alg / (r["dec_ms"] * 1e-3) / 1e9, "frac": alg / (r["dec_ms"] * 1e-3) / 1e9 / peak}}})
                torch.cuda.empty_cache()
            except AssertionError:
                raise
            except Exception as exc:  # noqa: BLE001
                others.append({"name": wname, "error": "%s: %s" % (type(exc).__name__, exc)})

    if rank == 0:
        enc_bytes = raw_rank + comp_bytes
        enc_gbs_hbm = enc_bytes / (enc_ms * 1e-3) / 1e9
        dec_gbs_hbm = enc_bytes / (dec_ms * 1e-3) / 1e9
        dominant = "encode_kernel" if enc_ms >= dec_ms else "decode_kernel"
        traffic = ncu_traffic(wl, ntiles, dominant)
        dom_ach = enc_gbs_hbm if enc_ms >= dec_ms else dec_gbs_hbm
        line = {
            "metric": "QB3 encode+decode raw-pixel GB/s", "value": value, "unit": "GB/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "u8" if ts == 1 else dname, "data": "synthetic",
            "config": {"workload": desc, "tiles_per_gpu": ntiles, "tile": [w, h, bands], "mode": mode,
                       "l2": "inputs (%.2f GB per GPU) larger than L2, no flush needed" % (raw_rank / 1e9),
                       "sharding": "contiguous tile ranges per rank, no collective"},
            "encode_gbs": raw_all / (enc_ms * 1e-3) / 1e9, "decode_gbs": raw_all / (dec_ms * 1e-3) / 1e9,
            "encode_ms": enc_ms, "decode_ms": dec_ms, "compressed_ratio": comp_all / raw_all,
            "roofline": {"bound": "hbm", "kernel": dominant, "achieved": dom_ach, "peak": peak, "unit": "GB/s",
                         "frac": dom_ach / peak, "traffic": traffic["bytes"] if traffic else None, "traffic_source": traffic,
                         "peak_source": peak_src, "algorithmic_bytes_per_launch": enc_bytes,
                         "encode": {"achieved": enc_gbs_hbm, "frac": enc_gbs_hbm / peak},
                         "decode": {"achieved": dec_gbs_hbm, "frac": dec_gbs_hbm / peak,
                                    "note": "one serial parse per stream bounds this pass: time = groups per stream x "
                                            "cycles per group of the scanner warp, whatever the batch (DESIGN 4.4)"}},
            "gpu_launches": int(launches), "clocks": sampler.result(), "cores_bound_per_rank": numa, "e2e": e2e,
            "other_workloads": others,
        }
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline(wl)
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
