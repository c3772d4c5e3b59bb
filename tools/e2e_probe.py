"""Times the pieces of the end-to-end step separately (PCIe copies alone, legs alone) to see what bounds it."""
import os, sys, time
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
src = device_synth_tiles(n, w, h, b, 0, dev)
h_src = torch.empty((n, tb), dtype=torch.uint8).pin_memory(); h_src.copy_(src)
h_out = torch.empty((n, tb), dtype=torch.uint8).pin_memory()
dst = torch.empty((n, slot), dtype=torch.uint8, device=dev)
sizes = torch.empty(n, dtype=torch.int64, device=dev); est = torch.empty(n, dtype=torch.int32, device=dev)
out = torch.empty((n, tb), dtype=torch.uint8, device=dev); dstat = torch.empty(n, dtype=torch.int32, device=dev)
def T(f, reps=5):
    f(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps): f()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps * 1e3
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
print("H2D 1.6GB ms", T(lambda: src.copy_(h_src, non_blocking=True)))
print("D2H 1.6GB ms", T(lambda: h_out.copy_(out, non_blocking=True)))
def both():
    with torch.cuda.stream(s1): src.copy_(h_src, non_blocking=True)
    with torch.cuda.stream(s2): h_out.copy_(out, non_blocking=True)
print("H2D+D2H concurrently ms", T(both))
print("encode ms", T(lambda: q.encode_batch(cfg, src, n, dst=dst, sizes=sizes, status=est)))
packed, offsets, total = q.pack_streams(dst, sizes, n)
print("pack ms", T(lambda: q.pack_streams(dst, sizes, n, packed=packed, offsets=offsets, total=total)))
print("decode(2048) ms", T(lambda: q.decode_batch(cfg, packed, offsets, sizes, n, out=out, status=dstat)))
for m in (256, 512, 1024):
    print("decode(%d) ms" % m, T(lambda: q.decode_batch(cfg, packed, offsets[:m], sizes[:m], m, out=out[:m], status=dstat[:m])))
def two():
    with torch.cuda.stream(s1): q.decode_batch(cfg, packed, offsets[:1024], sizes[:1024], 1024, out=out[:1024], status=dstat[:1024])
    with torch.cuda.stream(s2): q.decode_batch(cfg, packed, offsets[1024:], sizes[1024:], 1024, out=out[1024:], status=dstat[1024:])
print("decode 2x1024 on two streams ms", T(two))
ss = [torch.cuda.Stream() for _ in range(16)]
def many(k):
    m = n // k
    def f():
        for i in range(k):
            with torch.cuda.stream(ss[i]):
                q.decode_batch(cfg, packed, offsets[i*m:(i+1)*m], sizes[i*m:(i+1)*m], m, out=out[i*m:(i+1)*m], status=dstat[i*m:(i+1)*m])
    return f
for k in (4, 8, 16):
    print("decode %d x %d on %d streams ms" % (k, n // k, k), T(many(k)))
def manyenc(k):
    m = n // k
    def f():
        for i in range(k):
            with torch.cuda.stream(ss[i]):
                q.encode_batch(cfg, src[i*m:(i+1)*m], m, dst=dst[i*m:(i+1)*m], sizes=sizes[i*m:(i+1)*m], status=est[i*m:(i+1)*m])
    return f
print("encode 8 x 256 on 8 streams ms", T(manyenc(8)))
