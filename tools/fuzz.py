"""Randomised differential run against the oracle: geometry, type, mode (all nine), band map, quanta, content and batch
size drawn at random; every case is encoded (qb3cu_encode_batch), measured (qb3cu_encoded_size_batch) and decoded
(qb3cu_decode_batch) on the device, and the streams of its first and last tile are compared byte for byte with
oracle/qb3_oracle.c. Prints the failing configuration and stops at the first difference.

  python tools/fuzz.py [--seconds 120] [--seed 1] [--api]

--api: the same draw through the QB3.h calls instead (one image per handle, band maps as the caller would give them,
encoded twice on the handle so that its running state is carried, decoded through qb3_read_start / _info / _data).
"""
import argparse, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import torch
import qb3_b200 as q
from helpers import CONTENT_KINDS, DTYPES, PRODUCT_SO, QB3Lib, content, dtype_code, oracle, synth_tiles

ap = argparse.ArgumentParser()
ap.add_argument("--seconds", type=float, default=120)
ap.add_argument("--seed", type=int, default=1)
ap.add_argument("--api", action="store_true")
args = ap.parse_args()
rng = np.random.default_rng(args.seed)
O = oracle()
t_end = time.time() + args.seconds
ncases = 0
while time.time() < t_end:
    dt = np.dtype(DTYPES[rng.integers(0, len(DTYPES))])
    big = rng.random() < 0.2                      # now and then a tile large enough to be coded in parts
    w = int(rng.integers(4, 700 if big else 140))
    h = int(rng.integers(4, 700 if big else 100))
    if rng.random() < 0.3:
        w = (w + 3) & ~3
    b = int(rng.choice([1, 1, 2, 3, 3, 4, 5, 8, 17]))
    if w * h * b * dt.itemsize > (6 << 20):
        continue
    mode = int(rng.integers(0, 9))
    n = int(rng.choice([1, 2, 3, 7, 33])) if not big else int(rng.choice([1, 2, 5]))
    if not big and rng.random() < 0.08 and w * h * b * dt.itemsize <= 40000:
        n = int(rng.choice([2500, 3000, 4500, 5000, 9000]))   # 17 to 32 streams per CTA, and more than one CTA per SM holds
    kw = dict(mode=mode)
    if rng.random() < 0.4:                        # band map: a few core bands, the others derived from them
        cores = rng.choice(b, size=max(1, b // 3), replace=False)
        cb = [int(rng.choice(cores)) for _ in range(b)]
        for c in cores:
            cb[c] = int(c)
        kw["cband"] = cb
    if rng.random() < 0.25:
        kw["quanta"] = int(rng.choice([2, 3, 4, 5, 10, 37]))
        kw["away"] = bool(rng.integers(0, 2))
    if mode >= 4 and rng.random() < 0.1:          # an arbitrary 4x4 scan curve, written as an "SC" chunk
        perm = rng.permutation(16)
        kw["order"] = int(sum(int(v) << (4 * (15 - i)) for i, v in enumerate(perm)))
    packed_decode = rng.random() < 0.3            # decode from the packed blob instead of the slots
    kind = CONTENT_KINDS[rng.integers(0, len(CONTENT_KINDS))]
    few = [content(kind, w, h, b, dt, seed=int(rng.integers(1, 1 << 30))) if i % 2 == 0
           else synth_tiles(1, w, h, b, dt, seed=int(rng.integers(1, 1 << 30)))[0] for i in range(min(n, 34))]
    tiles = np.stack([few[i % len(few)] for i in range(n)])   # a large batch repeats its first 34 tiles
    desc = "dt=%s w=%d h=%d b=%d n=%d kind=%s packed=%s kw=%r" % (dt.name, w, h, b, n, kind, packed_decode, kw)
    if args.api:
        if b > 16 or "order" in kw:               # the header's QB3_MAXBANDS; the API has no setter for the curve
            continue
        try:
            P = P if "P" in globals() else QB3Lib(PRODUCT_SO, 16)
            got, want = P.encode(tiles[0], reps=2, **kw), O.encode(tiles[0], reps=2, **kw)
            assert got == want, "qb3_encode differs from the oracle (%d, %d vs %d, %d bytes)" % (len(got[0]), len(got[1]), len(want[0]), len(want[1]))
            for sbytes in want:
                ref, back = O.decode(sbytes), P.decode(sbytes)
                assert (ref is None) == (back is None), "qb3_read_data %s where the oracle %s" % ("fails" if back is None else "succeeds", "fails" if ref is None else "succeeds")
                assert ref is None or np.array_equal(ref, back), "qb3_read_data gives other pixels than the oracle"
        except Exception as exc:  # noqa: BLE001
            print("FAILED after %d cases: %s\n  %s: %s" % (ncases, desc, type(exc).__name__, exc), flush=True)
            sys.exit(1)
        ncases += 1
        continue
    try:
        cfg = q.config(w, h, b, dtype_code(dt), **kw)
        src = torch.from_numpy(np.ascontiguousarray(tiles).view(np.uint8).reshape(n, -1)).cuda()
        dst, sizes, st = q.encode_batch(cfg, src, n)
        only = q.encoded_size_batch(cfg, src, n)
        if packed_decode:
            blob, off, total = q.pack_streams(dst, sizes, n)
            out, st2 = q.decode_batch(cfg, blob, off, sizes, n)
        else:
            off = torch.arange(n, device="cuda", dtype=torch.int64) * dst.stride(0)
            out, st2 = q.decode_batch(cfg, dst, off, sizes, n)
        torch.cuda.synchronize()
        assert not st.any().item(), "encode status"
        sz, d = sizes.cpu().numpy(), dst.cpu().numpy()
        for t in {0, n - 1}:
            want = O.encode(tiles[t], **kw)
            assert int(sz[t]) == len(want) and d[t, :sz[t]].tobytes() == want, "tile %d differs from the oracle (%d vs %d bytes)" % (t, sz[t], len(want))
        assert torch.equal(only, sizes), "size-only pass differs: %r vs %r" % (only.cpu().tolist(), sizes.cpu().tolist())
        # A stream the reference itself cannot read back stays unreadable here: in the common factor modes the
        # reference drops a 64 bit group of more than 800 bits for an empty index group (QB3encode.h:704-708), and its
        # decoder, the oracle and the device all report failure on what follows.
        st2h = st2.cpu().tolist()
        for t in {0, n - 1}:
            ref = O.decode(O.encode(tiles[t], **kw))
            if ref is None:
                assert st2h[t] != 0, "tile %d: the oracle fails on this stream, the device does not" % t
                continue
            assert st2h[t] == 0, "decode status %r" % st2h
            back = out[t].cpu().numpy().view(dt).reshape(h, w, b)
            assert np.array_equal(back, ref), "tile %d decodes differently from the oracle" % t
    except Exception as exc:  # noqa: BLE001
        print("FAILED after %d cases: %s\n  %s: %s" % (ncases, desc, type(exc).__name__, exc), flush=True)
        sys.exit(1)
    ncases += 1
print("fuzz ok: %d cases in %.0f s, seed %d" % (ncases, args.seconds, args.seed))
