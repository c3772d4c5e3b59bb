"""Generates tests/golden/golden.json from the REFERENCE library compiled out of /root/reference
(oracle/_ref, see oracle/Makefile). Run in the build container only:

    python tests/golden/make_golden.py

Small cases carry the input pixels and the complete reference stream as hex; large synthetic tiles
(deterministic integer generator, helpers.synth_tiles) carry the stream length and sha256.
The reference has no golden vectors of its own; these pin parity on the GPU box where
/root/reference does not exist.
"""
import hashlib
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from helpers import (CONTENT_KINDS, MODE_BASE, MODE_BEST, MODE_FTL, content, ref, ref256, synth_tiles)  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden.json")


def small_case(name, img, lib, **kw):
    s = lib.encode(img, **kw)
    dec = lib.decode(s)
    return dict(name=name, kind="small", w=img.shape[1], h=img.shape[0], bands=img.shape[2], dtype=img.dtype.name,
                pixels=img.tobytes().hex(), stream=s.hex(),
                ref_decoded=None if dec is None else hashlib.sha256(dec.tobytes()).hexdigest(), **kw)


def synth_case(name, w, h, bands, dt, tile, lib, **kw):
    img = synth_tiles(1, w, h, bands, dt, t0=tile)[0]
    s = lib.encode(img, **kw)
    return dict(name=name, kind="synth", w=w, h=h, bands=bands, dtype=np.dtype(dt).name, tile=tile,
                length=len(s), sha256=hashlib.sha256(s).hexdigest(), head=s[:48].hex(), **kw)


def main():
    R, R256 = ref(), ref256()
    cases = []
    yy, xx = np.meshgrid(np.arange(8), np.arange(8), indexing="ij")
    # SURVEY 8a known-answer vectors K1..K8
    k1 = (xx + 2 * yy).astype(np.uint8)[:, :, None]
    for m, n in ((MODE_FTL, "K1"), (MODE_BASE, "K2"), (MODE_BEST, "K3")):
        cases.append(small_case(n, k1, R, mode=m))
    cases.append(small_case("K4", np.zeros((8, 8, 1), np.uint8), R, mode=MODE_FTL))
    k5 = (10 * np.arange(3)[None, None, :] + xx[:, :, None] + yy[:, :, None]).astype(np.uint8)
    cases.append(small_case("K5", k5, R, mode=MODE_FTL))
    cases.append(small_case("K6", (5 * (xx + 3 * yy)).astype(np.uint16)[:, :, None], R, mode=MODE_BEST))
    y7, x7 = np.meshgrid(np.arange(7), np.arange(9), indexing="ij")
    cases.append(small_case("K7", (100 - 7 * x7 + 3 * y7).astype(np.int32)[:, :, None], R, mode=MODE_BASE, quanta=3))
    cases.append(small_case("K8", np.zeros((64, 64, 1), np.uint8), R, mode=MODE_BEST))
    # all types x FTL/BASE/BEST x content kinds on a ragged multi band shape
    for dt in ("uint8", "int8", "uint16", "int16", "uint32", "int32", "uint64", "int64"):
        for i, kind in enumerate(CONTENT_KINDS):
            img = content(kind, 13, 10, 2, dt, seed=100 + i)
            for mode in (MODE_FTL, MODE_BASE, MODE_BEST):
                cases.append(small_case("c_%s_%s_%d" % (dt, kind, mode), img, R, mode=mode, cband=[0, 0]))
    # legacy modes
    for mode in (0, 1, 2, 3, 5, 6):
        cases.append(small_case("legacy_%d" % mode, content("steps", 12, 12, 1, "uint16", seed=3), R, mode=mode))
    # quanta
    for dt in ("uint8", "int16", "int32", "uint64"):
        for q, away in ((2, True), (3, False), (4, False), (7, True), (10, False)):
            cases.append(small_case("q_%s_%d_%d" % (dt, q, away), content("signed", 9, 7, 1, dt, seed=q), R,
                                    mode=MODE_BASE, quanta=q, away=away))
    # small / narrow / short images (patched reference build) and many bands
    for (w, h, b) in ((3, 40, 3), (1, 17, 1), (2, 9, 1), (40, 2, 3), (17, 1, 1), (5, 4, 1), (4, 4, 2), (3, 5, 1), (60, 3, 1)):
        for dt in ("uint8", "int32", "uint64"):
            for q in (1, 3):
                cases.append(small_case("small_%dx%dx%d_%s_q%d" % (w, h, b, dt, q), content("synth", w, h, b, dt, seed=w + h),
                                        R256, mode=MODE_BASE, quanta=q))
    for bands in (17, 256):
        cases.append(small_case("bands_%d" % bands, content("synth", 8, 4, bands, "uint16"), R256, mode=MODE_BEST,
                                cband=[0] * bands))
    # the BASELINE configs, as digests
    for t in range(4):
        cases.append(synth_case("C2_tile%d" % t, 512, 512, 3, np.uint8, t, R, mode=MODE_FTL))
    for t in range(2):
        for mode in (MODE_BASE, MODE_BEST):
            cases.append(synth_case("C3_tile%d_m%d" % (t, mode), 512, 512, 8, np.uint16, t, R, mode=mode, cband=[0] * 8))
    for dt in (np.int32, np.uint64):
        for q in (1, 3):
            cases.append(synth_case("C4_%s_q%d" % (np.dtype(dt).name, q), 513, 511, 1, dt, 0, R, mode=MODE_FTL, quanta=q))
    cases.append(synth_case("C5_64x64x16_u16", 64, 64, 16, np.uint16, 0, R, mode=MODE_BEST, cband=[0] * 16))
    cases.append(synth_case("C5_1024_u8", 1024, 1024, 1, np.uint8, 0, R, mode=MODE_BASE))
    with open(OUT, "w") as f:
        json.dump(dict(generator="tests/golden/make_golden.py", source="oracle/_ref (reference compiled from /root/reference)",
                       cases=cases), f, indent=0)
    print(len(cases), "cases,", os.path.getsize(OUT), "bytes")


if __name__ == "__main__":
    main()
