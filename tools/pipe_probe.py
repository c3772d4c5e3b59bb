"""Times the host buffer pipeline (qb3cu_pipe_*) alone and with an encode and a decode running side by side."""
import os, sys, time, threading
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import qb3_b200 as q
from bench import device_synth_tiles
dev = torch.device("cuda", 0)
n, w, h, b = 2048, 512, 512, 3
tb = w * h * b
cfg = q.config(w, h, b, 0, mode=8)
slot = q.slot_bytes(cfg)
h_src = torch.empty((n, tb), dtype=torch.uint8).pin_memory(); h_src.copy_(device_synth_tiles(n, w, h, b, 0, dev))
h_out = torch.empty((n, tb), dtype=torch.uint8).pin_memory()
h_packed = [torch.empty(n * slot, dtype=torch.uint8).pin_memory() for _ in range(2)]
h_off = [torch.zeros(n, dtype=torch.int64) for _ in range(2)]
h_sz = [torch.zeros(n, dtype=torch.int64) for _ in range(2)]
h_st = torch.zeros(n, dtype=torch.int32)
print("CUDA_DEVICE_MAX_CONNECTIONS", os.environ.get("CUDA_DEVICE_MAX_CONNECTIONS"))
os.environ.get("CUDA_DEVICE_MAX_CONNECTIONS") or sys.exit("set CUDA_DEVICE_MAX_CONNECTIONS=32")
for ec, ed, chunk, depth in [(128, 12, 512, 4), (128, 16, 512, 4), (64, 24, 512, 4), (128, 12, 256, 8), (256, 6, 512, 4), (128, 8, 512, 4)]:
    ep, dp = q.Pipe(cfg, ec, ed), q.Pipe(cfg, chunk, depth)
    def enc(k): ep.encode(h_src, n, h_packed[k % 2], h_off[k % 2], h_sz[k % 2])
    def dec(k): dp.decode(h_packed[k % 2], h_off[k % 2], h_sz[k % 2], n, h_out, h_st)
    enc(0); enc(1); dec(0); dec(1)
    def T(f, reps=4):
        t0 = time.perf_counter()
        for k in range(reps): f(k)
        return (time.perf_counter() - t0) / reps * 1e3
    te, td = T(enc), T(dec)
    def both(k):
        a = threading.Thread(target=enc, args=(k,)); bth = threading.Thread(target=dec, args=(k + 1,))
        a.start(); bth.start(); a.join(); bth.join()
    tb2 = T(both)
    assert torch.equal(h_out, h_src) and not h_st.any()
    print("enc %d x %d, dec chunk %4d depth %2d: encode %.1f ms, decode %.1f ms, sequential %.1f ms (%.1f GB/s), side by side %.1f ms (%.1f GB/s)"
          % (ec, ed, chunk, depth, te, td, te + td, n * tb / (te + td) / 1e6, tb2, n * tb / tb2 / 1e6))
    ep.close(); dp.close()
