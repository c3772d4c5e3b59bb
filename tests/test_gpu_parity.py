"""GPU parity tests (run with -m gpu on the B200 box): the CUDA path, called through the C ABI, against the
oracle on seeded inputs, against the committed golden vectors made by the reference library, and -- at the
BASELINE sizes -- through size independent properties (round trip, checksum of stream sizes)."""
import ctypes as C
import hashlib
import os

import numpy as np
import pytest

import qb3_b200 as q
from helpers import (CONTENT_KINDS, DTYPES, MODE_BASE, MODE_BEST, MODE_CF, MODE_CF_H, MODE_FTL, MODE_RLE, MODE_RLE_H, PRODUCT_SO, REF_SO, QB3Lib, content, dtype_code,
                     golden_cases, golden_check_stream, golden_image, golden_kwargs, have_ref, oracle, synth_tiles)

pytestmark = pytest.mark.gpu

ENC_MODES_DONE = set(range(9))  # encoder modes implemented on the device


def product():
    return QB3Lib(PRODUCT_SO, 256)


def torch_mod():
    import torch
    assert torch.cuda.is_available(), "these tests need a CUDA device"
    return torch


# ------------------------------------------------------------------ QB3.h API, single images

def test_api_encode_matches_golden():
    P = product()
    bad = []
    for case in golden_cases():
        kw = golden_kwargs(case)
        if kw.get("mode", MODE_FTL) not in ENC_MODES_DONE:
            continue
        try:
            golden_check_stream(case, P.encode(golden_image(case), **kw))
        except (AssertionError, RuntimeError) as e:
            bad.append((case["name"], type(e).__name__))
    assert not bad, "%d of %d golden cases differ: %s" % (len(bad), len(golden_cases()), bad[:12])


def test_api_decode_matches_golden():
    P, O = product(), oracle()
    bad = []
    for case in golden_cases():
        if case["kind"] != "small":
            continue
        s = bytes.fromhex(case["stream"])
        os.environ["QB3_REF_COMPAT"] = "1"  # decode exactly like the reference decoder
        try:
            d = P.decode(s)
        finally:
            del os.environ["QB3_REF_COMPAT"]
        if case["ref_decoded"] is None:
            ok = d is None
        else:
            ok = d is not None and hashlib.sha256(d.tobytes()).hexdigest() == case["ref_decoded"]
        if ok and d is not None and case.get("quanta", 1) == 1:
            d2 = P.decode(s)  # spec behaviour: identity band map by default -> the original pixels
            ok = d2 is not None and np.array_equal(d2, golden_image(case))
        if not ok:
            bad.append(case["name"])
    assert not bad, "%d golden streams decode differently: %s" % (len(bad), bad[:12])


@pytest.mark.parametrize("dt", DTYPES)
def test_api_round_trip_and_oracle_all_content(dt):
    P, O = product(), oracle()
    for i, kind in enumerate(CONTENT_KINDS):
        for (w, h, b) in ((17, 9, 3), (33, 31, 1), (12, 8, 5)):
            img = content(kind, w, h, b, dt, seed=7 * i + w)
            for mode in (MODE_FTL, MODE_BASE, MODE_BEST, 0, 1, 2, 5):
                if mode not in ENC_MODES_DONE:
                    continue
                s = P.encode(img, mode=mode)
                assert s == O.encode(img, mode=mode), (kind, w, h, b, mode)
                d, od = P.decode(s), O.decode(s)
                assert (d is None) == (od is None) and (d is None or np.array_equal(d, od)), (kind, w, h, b, mode)
                # The reference cannot read back two kinds of its own streams, and neither may we: an RLE stream
                # that expands past the raw size (tiny images, QB3encode.cpp:543 vs QB3decode.cpp:401) and 64 bit
                # BEST groups longer than 800 bits, which it drops (QB3encode.h:564,705). Everything else round trips.
                known_defect = (od is None and s[10] in (2, 3, 6, 7)) or (np.dtype(dt).itemsize == 8 and mode in (1, 5, 7))
                assert known_defect or np.array_equal(d, img), (kind, w, h, b, mode)


def test_api_quanta_stride_state():
    P, O = product(), oracle()
    for dt in (np.uint8, np.int16, np.int32, np.uint64):
        img = content("signed", 21, 14, 2, dt, seed=3)
        for qv, away in ((2, False), (2, True), (3, False), (4, True), (5, False), (10, True)):
            s = P.encode(img, mode=MODE_BASE, quanta=qv, away=away)
            assert s == O.encode(img, mode=MODE_BASE, quanta=qv, away=away)
            assert np.array_equal(P.decode(s), O.decode(s))
    # strided source and destination (stride in values, QB3.h:116,147)
    back = content("synth", 40, 12, 3, np.uint16)
    view = np.ascontiguousarray(back.reshape(12, 120)[:, :81].reshape(12, 27, 3))
    L = P.lib
    e = L.qb3_create_encoder(27, 12, 3, dtype_code(np.uint16))
    L.qb3_set_encoder_stride(e, 120)
    dst = np.zeros(L.qb3_max_encoded_size(e), np.uint8)
    n = L.qb3_encode(e, back.ctypes.data, dst.ctypes.data)
    L.qb3_destroy_encoder(e)
    s = dst[:n].tobytes()
    assert s == O.encode(view)
    out = P.decode(s, stride=100)
    assert np.array_equal(out[:, :81].reshape(12, 27, 3), view) and not out[:, 81:].any()
    # running state persists across qb3_encode calls on one handle
    img = content("synth", 16, 16, 1, np.uint8)
    assert P.encode(img, mode=MODE_BASE, reps=2) == O.encode(img, mode=MODE_BASE, reps=2)
    # ... but not for quantised images or images with a side under 4: the reference codes those through a copy of the
    # handle (QB3encode.cpp:405, :352), so the second stream equals the first and decodes
    for kw in (dict(mode=MODE_BASE, quanta=3), dict(mode=MODE_BEST, quanta=5)):
        a = P.encode(img, reps=2, **kw)
        assert a == O.encode(img, reps=2, **kw) and a[0] == a[1]
    small = content("synth", 3, 40, 2, np.uint16)
    a = P.encode(small, mode=MODE_BASE, reps=2)
    assert a == O.encode(small, mode=MODE_BASE, reps=2) and a[0] == a[1]
    assert np.array_equal(P.decode(a[1]), small)
    # ... and for an image large enough to be coded in parts, BEST included: the second call starts from the values,
    # rungs and common factors the first one left behind
    big = (synth_tiles(1, 640, 512, 3, np.uint8)[0] >> 2) * np.uint8(5)
    for mode in (MODE_BASE, MODE_CF_H, MODE_BEST):
        a = P.encode(big, mode=mode, reps=2)
        assert a == O.encode(big, mode=mode, reps=2), "mode %d" % mode


def test_api_small_and_many_bands():
    P, O = product(), oracle()
    for (w, h, b) in ((3, 100, 3), (1, 17, 1), (2, 9, 1), (100, 2, 3), (17, 1, 1), (1000, 3, 1), (2, 2000, 1), (5, 4, 1), (4, 4, 2), (2, 2, 1)):
        for dt in (np.uint8, np.int32, np.uint64):
            img = content("synth", w, h, b, dt, seed=w + h)
            for qv in (1, 3):
                s = P.encode(img, mode=MODE_BASE, quanta=qv)
                assert s == O.encode(img, mode=MODE_BASE, quanta=qv), (w, h, b, qv)
                d, od = P.decode(s), O.decode(s)
                assert (d is None) == (od is None) and (d is None or np.array_equal(d, od)), (w, h, b, qv)
    for bands in (17, 64, 256):
        img = content("synth", 12, 8, bands, np.uint16)
        s = P.encode(img, mode=MODE_BASE, cband=[0] * bands)
        assert s == O.encode(img, mode=MODE_BASE, cband=[0] * bands)
        assert np.array_equal(P.decode(s), img)


def test_api_stored_fallback_and_errors():
    P, O = product(), oracle()
    img = content("noise", 32, 32, 1, np.uint8)
    s = P.encode(img)
    assert s[10] == 255 and s == O.encode(img)  # incompressible -> stored (QB3encode.cpp:570-573)
    assert np.array_equal(P.decode(s), img)
    assert P.decode(s + b"\x00") is None        # stored payload must be exact (QB3decode.cpp:360)
    ok = P.encode(content("synth", 16, 16, 1, np.uint8))
    assert P.decode(ok + b"\x00") is None       # one spare byte is an error (QB3decode.h:411)
    assert P.decode(ok[:-1] + b"\xff\xff") is None


# ------------------------------------------------------------------ batched C ABI

def encode_tiles(tiles, **kw):
    torch = torch_mod()
    n, h, w, b = tiles.shape
    cfg = q.config(w, h, b, dtype_code(tiles.dtype), **kw)
    src = torch.from_numpy(tiles.view(np.uint8).reshape(n, -1)).cuda()
    dst, sizes, status = q.encode_batch(cfg, src, n)
    torch.cuda.synchronize()
    return cfg, dst, sizes, status


@pytest.mark.parametrize("shape,dt,kw", [
    ((512, 512, 3), np.uint8, dict(mode=MODE_FTL)),                      # BASELINE config 2
    ((512, 512, 8), np.uint16, dict(mode=MODE_BASE, cband=[0] * 8)),     # config 3
    ((512, 512, 8), np.uint16, dict(mode=MODE_BEST, cband=[0] * 8)),
    ((513, 511, 1), np.int32, dict(mode=MODE_FTL)),                      # config 4
    ((513, 511, 1), np.uint64, dict(mode=MODE_FTL, quanta=3)),
    ((64, 64, 16), np.uint16, dict(mode=MODE_BASE)),                     # config 5 corners
    ((1024, 1024, 1), np.uint8, dict(mode=MODE_BASE)),
    ((100, 60, 3), np.int8, dict(mode=MODE_FTL)),
])
def test_batch_encode_matches_oracle_and_round_trips(shape, dt, kw):
    torch = torch_mod()
    if kw.get("mode", MODE_FTL) not in ENC_MODES_DONE:
        pytest.skip("mode not on the device yet")
    w, h, b = shape
    n = 6
    tiles = synth_tiles(n, w, h, b, dt)
    cfg, dst, sizes, status = encode_tiles(tiles, **kw)
    sizes_h, dst_h = sizes.cpu().numpy(), dst.cpu().numpy()
    assert not status.cpu().numpy().any()
    O = oracle()
    for t in range(n):
        want = O.encode(tiles[t], **kw)
        got = dst_h[t, :sizes_h[t]].tobytes()
        assert got == want, "tile %d: %d vs %d bytes" % (t, len(got), len(want))
    # decode the batch in place on the GPU
    offsets = (torch.arange(n, device="cuda", dtype=torch.int64) * dst.stride(0))
    out, st = q.decode_batch(cfg, dst, offsets, sizes, n)
    torch.cuda.synchronize()
    assert not st.cpu().numpy().any()
    dec = out.cpu().numpy().view(tiles.dtype).reshape(tiles.shape)
    if kw.get("quanta", 1) == 1:
        assert np.array_equal(dec, tiles)
    else:
        for t in range(n):
            assert np.array_equal(dec[t], O.decode(dst_h[t, :sizes_h[t]].tobytes()))


def test_batch_full_size_properties():
    """BASELINE config 2 at full size: 4096 tiles of 512x512x3 u8. Every tile must round trip, and the
    sizes of a deterministic sample must equal the oracle's."""
    torch = torch_mod()
    n, w, h, b = 4096, 512, 512, 3
    from bench import device_synth_tiles
    src = device_synth_tiles(n, w, h, b, 0, torch.device("cuda"))
    cfg = q.config(w, h, b, 0, mode=MODE_FTL)
    dst, sizes, status = q.encode_batch(cfg, src, n)
    offsets = torch.arange(n, device="cuda", dtype=torch.int64) * dst.stride(0)
    out, st = q.decode_batch(cfg, dst, offsets, sizes, n)
    torch.cuda.synchronize()
    assert not status.any().item() and not st.any().item()
    assert torch.equal(out.view(-1), src.view(-1))
    sizes_h = sizes.cpu().numpy()
    O = oracle()
    for t in (0, 1, 2047, 4095):
        tile = synth_tiles(1, w, h, b, np.uint8, t0=t)[0]
        assert np.array_equal(src[t].cpu().numpy().reshape(h, w, b), tile)  # device generator == host generator
        want = O.encode(tile)
        assert sizes_h[t] == len(want)
        assert dst[t, :sizes_h[t]].cpu().numpy().tobytes() == want


@pytest.mark.parametrize("mode", [MODE_BASE, MODE_BEST])
def test_batch_full_size_config3(mode):
    """BASELINE config 3 at full size: 1024 Landsat-like tiles of 512x512x8 u16, core band 0, BASE and BEST. Every tile
    round trips, eight tiles spread over the batch are byte-identical to the oracle's streams, and the size-only pass
    agrees with the encode on all 1024 sizes."""
    torch = torch_mod()
    n, w, h, b = 1024, 512, 512, 8
    from bench import device_synth_tiles
    src = device_synth_tiles(n, w, h, b, 2, torch.device("cuda"))
    cfg = q.config(w, h, b, 2, mode=mode, cband=[0] * b)
    dst, sizes, status = q.encode_batch(cfg, src, n)
    offsets = torch.arange(n, device="cuda", dtype=torch.int64) * dst.stride(0)
    out, st = q.decode_batch(cfg, dst, offsets, sizes, n)
    only = q.encoded_size_batch(cfg, src, n)
    torch.cuda.synchronize()
    assert not status.any().item() and not st.any().item()
    assert torch.equal(out.view(-1), src.view(-1))
    assert torch.equal(only, sizes)
    sizes_h = sizes.cpu().numpy()
    O = oracle()
    for t in (0, 1, 127, 300, 511, 777, 1022, 1023):
        tile = synth_tiles(1, w, h, b, np.uint16, t0=t)[0]
        assert np.array_equal(src[t].cpu().numpy().view(np.uint16).reshape(h, w, b), tile)  # device generator == host generator
        want = O.encode(tile, mode=mode, cband=[0] * b)
        assert sizes_h[t] == len(want) and dst[t, :sizes_h[t]].cpu().numpy().tobytes() == want, "tile %d" % t


def test_batch_decode_reports_bad_streams():
    torch = torch_mod()
    tiles = synth_tiles(4, 32, 32, 1, np.uint8)
    cfg, dst, sizes, status = encode_tiles(tiles)
    dst_h, sizes_h = dst.cpu().numpy().copy(), sizes.cpu().numpy().copy()
    dst_h[1, 0] = ord("X")           # bad signature
    sizes_h[2] += 3                  # trailing garbage
    dst_h[3, 4] = 63                 # geometry differs from the batch
    d2, s2 = torch.from_numpy(dst_h).cuda(), torch.from_numpy(sizes_h).cuda()
    offsets = torch.arange(4, device="cuda", dtype=torch.int64) * d2.stride(0)
    out, st = q.decode_batch(cfg, d2, offsets, s2, 4)
    torch.cuda.synchronize()
    assert st.cpu().numpy().tolist() == [q.TILE_OK, q.TILE_BAD_HEADER, q.TILE_CORRUPT, q.TILE_BAD_HEADER]
    assert np.array_equal(out[0].cpu().numpy().reshape(32, 32, 1), tiles[0])


def test_pack_streams_and_decode_packed():
    """qb3cu_pack_streams: streams back to back at 16 byte aligned offsets, decodable from there."""
    torch = torch_mod()
    tiles = synth_tiles(9, 100, 60, 3, np.uint8)
    cfg, dst, sizes, status = encode_tiles(tiles, mode=MODE_BASE)
    packed, offsets, total = q.pack_streams(dst, sizes, 9)
    out, st = q.decode_batch(cfg, packed, offsets, sizes, 9)
    torch.cuda.synchronize()
    sz, off, dst_h, pk = sizes.cpu().numpy(), offsets.cpu().numpy(), dst.cpu().numpy(), packed.cpu().numpy()
    want_off = np.concatenate([[0], np.cumsum((sz + 15) // 16 * 16)])
    assert off.tolist() == want_off[:-1].tolist() and int(total.item()) == want_off[-1]
    for t in range(9):
        assert pk[off[t]:off[t] + sz[t]].tobytes() == dst_h[t, :sz[t]].tobytes()
    assert not st.cpu().numpy().any()
    assert np.array_equal(out.cpu().numpy().reshape(tiles.shape), tiles)


# ------------------------------------------------------------------ host buffer pipeline (qb3cu_pipe_*)

@pytest.mark.parametrize("shape,dt,mode,chunk,depth", [
    ((37, 64, 48, 3), np.uint8, MODE_FTL, 8, 3),     # five chunks, the last one short, stages reused
    ((10, 40, 36, 2), np.uint16, MODE_BEST, 4, 2),
    ((5, 33, 21, 1), np.int32, MODE_BASE, 0, 0),     # default chunking: one chunk
    ((6, 3, 50, 2), np.uint8, MODE_BASE, 4, 2),      # narrow images go through the reorder path
    ((9, 40, 70, 3), np.uint8, MODE_BASE, 4, 2),     # several row bands per chunk, height not a multiple of four
    ((6, 36, 130, 1), np.int16, MODE_FTL, 3, 3),
])
def test_pipe_host_buffers_match_oracle_and_round_trip(shape, dt, mode, chunk, depth):
    """qb3cu_pipe_encode / qb3cu_pipe_decode: host pixels -> packed streams in host memory, byte identical to the
    oracle's, 16 byte aligned starts in tile order; and back to the same pixels."""
    torch_mod()
    n, w, h, b = shape
    tiles = synth_tiles(n, w, h, b, dt)
    cfg = q.config(w, h, b, dtype_code(dt), mode=mode)
    pipe = q.Pipe(cfg, chunk, depth)
    packed = np.zeros(n * q.slot_bytes(cfg), np.uint8)
    offsets, sizes = np.zeros(n, np.uint64), np.zeros(n, np.uint64)
    total = pipe.encode(tiles, n, packed, offsets, sizes)
    want_off = 0
    for t in range(n):
        want = oracle().encode(tiles[t], mode=mode)
        assert int(offsets[t]) == want_off and int(sizes[t]) == len(want)
        assert packed[want_off:want_off + len(want)].tobytes() == want, "tile %d differs from the oracle" % t
        want_off += (len(want) + 15) // 16 * 16
    assert total == want_off
    out = np.zeros_like(tiles)
    status = np.full(n, 99, np.uint32)
    pipe.decode(packed, offsets, sizes, n, out, status)
    assert not status.any() and np.array_equal(out, tiles)
    pipe.close()


def test_pipe_pitched_strided_pinned_and_errors():
    torch = torch_mod()
    n, w, h, b = 7, 24, 20, 3
    tiles = synth_tiles(n, w, h, b, np.uint8)
    # tiles further apart than their size, in page locked memory
    cfg = q.config(w, h, b, q.U8, mode=MODE_BASE)
    pitch = tiles[0].nbytes + 100
    h_src = torch.zeros(n * pitch, dtype=torch.uint8).pin_memory()
    h_src.view(n, pitch)[:, :tiles[0].nbytes] = torch.from_numpy(tiles.reshape(n, -1))
    pipe = q.Pipe(cfg, 3, 2)
    packed = torch.zeros(n * q.slot_bytes(cfg), dtype=torch.uint8).pin_memory()
    offsets, sizes = torch.zeros(n, dtype=torch.int64), torch.zeros(n, dtype=torch.int64)
    pipe.encode(h_src, n, packed, offsets, sizes, tile_pitch=pitch)
    for t in range(n):
        o, s = int(offsets[t]), int(sizes[t])
        assert packed[o:o + s].numpy().tobytes() == oracle().encode(tiles[t], mode=MODE_BASE)
    h_out = torch.full((n * pitch,), 7, dtype=torch.uint8).pin_memory()
    status = torch.zeros(n, dtype=torch.int32)
    pipe.decode(packed, offsets, sizes, n, h_out, status, tile_pitch=pitch)
    v = h_out.view(n, pitch).numpy()
    assert np.array_equal(v[:, :tiles[0].nbytes].reshape(tiles.shape), tiles)
    assert (v[:, tiles[0].nbytes:] == 7).all(), "the gap between tiles was written"
    # too little room for the streams
    with pytest.raises(RuntimeError):
        pipe.encode(h_src, n, packed[:int(offsets[3])], offsets, sizes, tile_pitch=pitch)
    # a damaged stream is reported per tile, its neighbours decode
    bad = packed.clone()
    bad[int(offsets[2])] = 0
    pipe.decode(bad, offsets, sizes, n, h_out, status, tile_pitch=pitch)
    assert status.tolist() == [0, 0, q.TILE_BAD_HEADER, 0, 0, 0, 0]
    pipe.close()
    # lines further apart than their length
    stride = w * b + 5
    cfg2 = q.config(w, h, b, q.U8, mode=MODE_FTL, stride=stride)
    buf = np.full((n, h, stride), 9, np.uint8)
    buf[:, :, :w * b] = tiles.reshape(n, h, w * b)
    pipe2 = q.Pipe(cfg2, 4, 2)
    packed2 = np.zeros(n * q.slot_bytes(cfg2), np.uint8)
    off2, sz2 = np.zeros(n, np.uint64), np.zeros(n, np.uint64)
    pipe2.encode(buf, n, packed2, off2, sz2)
    for t in range(n):
        assert packed2[int(off2[t]):int(off2[t] + sz2[t])].tobytes() == oracle().encode(tiles[t], mode=MODE_FTL)
    out2 = np.full_like(buf, 5)
    st2 = np.zeros(n, np.uint32)
    pipe2.decode(packed2, off2, sz2, n, out2, st2)
    assert not st2.any() and np.array_equal(out2[:, :, :w * b], buf[:, :, :w * b]) and (out2[:, :, w * b:] == 5).all()
    pipe2.close()


def test_pipe_rle_and_stored_tiles():
    """Tiles the two pass decode leaves to the kernels after it (RLE streams, stored tiles) travel back whole."""
    torch_mod()
    n, w, h, b = 8, 64, 72, 1
    tiles = synth_tiles(n, w, h, b, np.uint8)
    tiles[1] = 0                                                        # RLE pays: mode byte 7
    tiles[4] = np.random.default_rng(5).integers(0, 256, tiles[4].shape, dtype=np.uint8)   # incompressible: stored
    cfg = q.config(w, h, b, q.U8, mode=MODE_BEST)
    pipe = q.Pipe(cfg, 8, 2)
    packed = np.zeros(n * q.slot_bytes(cfg), np.uint8)
    offsets, sizes = np.zeros(n, np.uint64), np.zeros(n, np.uint64)
    pipe.encode(tiles, n, packed, offsets, sizes)
    for t in range(n):
        assert packed[int(offsets[t]):int(offsets[t] + sizes[t])].tobytes() == oracle().encode(tiles[t], mode=MODE_BEST)
    assert packed[int(offsets[1]) + 10] == 7 and packed[int(offsets[4]) + 10] == 255
    out = np.zeros_like(tiles)
    status = np.full(n, 99, np.uint32)
    pipe.decode(packed, offsets, sizes, n, out, status)
    assert not status.any() and np.array_equal(out, tiles)
    pipe.close()


# ------------------------------------------------------------------ cqb3cu, the command line tool

def _write_pnm(path, img):
    h, w, b = img.shape
    big = img.astype(">u2") if img.dtype == np.uint16 else img
    with open(path, "wb") as f:
        f.write(b"P%d\n%d %d\n%d\n" % (5 if b == 1 else 6, w, h, 65535 if img.dtype == np.uint16 else 255))
        f.write(big.tobytes())


def test_cli_encode_decode_bandmix_and_folder(tmp_path):
    """cqb3cu with the reference tool's options: streams byte identical to the oracle for the same settings, PNM and
    raw round trips, -m x = the smallest of the ten band maps, folder mode through the batched pipeline."""
    import subprocess
    torch_mod()
    exe = os.path.join(os.path.dirname(PRODUCT_SO), "..", "apps", "cqb3cu")
    if not os.path.exists(exe):
        subprocess.run(["make", "-s", "-C", os.path.dirname(exe)], check=True)
    rgb = synth_tiles(3, 70, 50, 3, np.uint8)
    gray16 = synth_tiles(1, 37, 41, 1, np.uint16)[0]
    p = str(tmp_path / "a.ppm")
    _write_pnm(p, rgb[0])
    run = lambda *a: subprocess.run([exe, *a], capture_output=True, text=True, cwd=str(tmp_path))
    # default (BASE), -b (BEST), -f (FTL), -q 3, explicit band map
    for flags, kw in (([], dict(mode=MODE_BASE)), (["-b"], dict(mode=MODE_BEST)), (["-f"], dict(mode=MODE_FTL)),
                      (["-q", "3"], dict(mode=MODE_BASE, quanta=3)), (["-m", "0,0,2"], dict(mode=MODE_BASE, cband=[0, 0, 2]))):
        r = run(*flags, p, "o.qb3")
        assert r.returncode == 0, r
        assert (tmp_path / "o.qb3").read_bytes() == oracle().encode(rgb[0], **kw), flags
    # decode back to PNM
    r = run("-f", p, "f.qb3"); assert r.returncode == 0
    r = run("-d", "f.qb3", "back.ppm"); assert r.returncode == 0, r
    assert (tmp_path / "back.ppm").read_bytes() == (tmp_path / "a.ppm").read_bytes()
    # 16 bit gray through PNM, int32 through raw
    _write_pnm(str(tmp_path / "g.pgm"), gray16)
    assert run("-b", "g.pgm", "g.qb3").returncode == 0
    assert (tmp_path / "g.qb3").read_bytes() == oracle().encode(gray16, mode=MODE_BEST)
    i32 = synth_tiles(1, 21, 9, 2, np.int32)[0]
    (tmp_path / "i.raw").write_bytes(i32.tobytes())
    assert run("-s", "21x9x2:i32", "i.raw", "i.qb3").returncode == 0
    assert (tmp_path / "i.qb3").read_bytes() == oracle().encode(i32, mode=MODE_BASE)
    assert run("-d", "-s", "21x9x2:i32", "i.qb3", "i2.raw").returncode == 0
    assert (tmp_path / "i2.raw").read_bytes() == i32.tobytes()
    # band mix search: the smallest of the reference's ten maps, the first one on ties
    combos = [[1, 1, 1], [0, 0, 0], [0, 0, 2], [0, 1, 0], [0, 1, 1], [0, 1, 2], [0, 2, 2], [1, 1, 2], [2, 1, 2], [2, 2, 2]]
    streams = [oracle().encode(rgb[0], mode=MODE_BASE, cband=c) for c in combos]
    best = min(range(10), key=lambda k: (len(streams[k]), k))
    assert run("-m", "x", p, "x.qb3").returncode == 0
    assert (tmp_path / "x.qb3").read_bytes() == streams[best]
    # folder mode
    d = tmp_path / "in"; d.mkdir(); o = tmp_path / "out"; o.mkdir()
    for t in range(3):
        _write_pnm(str(d / ("t%d.ppm" % t)), rgb[t])
    _write_pnm(str(d / "g.pgm"), gray16)
    r = run("-f", str(d), str(o)); assert r.returncode == 0, r
    for t in range(3):
        assert (o / ("t%d.qb3" % t)).read_bytes() == oracle().encode(rgb[t], mode=MODE_FTL)
    assert (o / "g.qb3").read_bytes() == oracle().encode(gray16, mode=MODE_FTL)
    # and back: the folder of streams decoded as batches per geometry
    back = tmp_path / "back"; back.mkdir()
    r = run("-d", str(o), str(back)); assert r.returncode == 0, r
    for t in range(3):
        assert (back / ("t%d.pnm" % t)).read_bytes() == (d / ("t%d.ppm" % t)).read_bytes()
    assert (back / "g.pnm").read_bytes() == (d / "g.pgm").read_bytes()


def test_two_pipes_on_two_threads():
    """An encode pipe and a decode pipe used at the same time from two host threads (what bench.py's end to end leg
    does): results identical to the calls made alone."""
    import threading
    torch = torch_mod()
    n, w, h, b = 96, 128, 96, 3
    tiles = synth_tiles(n, w, h, b, np.uint8)
    cfg = q.config(w, h, b, q.U8, mode=MODE_FTL)
    ep, dp = q.Pipe(cfg, 16, 3), q.Pipe(cfg, 32, 2)
    h_src = torch.from_numpy(tiles.reshape(n, -1)).pin_memory()
    packed = [torch.zeros(n * q.slot_bytes(cfg), dtype=torch.uint8).pin_memory() for _ in range(2)]
    off = [torch.zeros(n, dtype=torch.int64) for _ in range(2)]
    sz = [torch.zeros(n, dtype=torch.int64) for _ in range(2)]
    out = torch.zeros_like(h_src).pin_memory()
    st = torch.ones(n, dtype=torch.int32)
    ep.encode(h_src, n, packed[0], off[0], sz[0])
    errors = []

    def run(f, *a):
        try:
            f(*a)
        except Exception as e:  # noqa: BLE001
            errors.append(e)

    for k in range(1, 5):   # encode into one buffer while the other one is decoded
        ta = threading.Thread(target=run, args=(ep.encode, h_src, n, packed[k % 2], off[k % 2], sz[k % 2]))
        tb = threading.Thread(target=run, args=(dp.decode, packed[(k - 1) % 2], off[(k - 1) % 2], sz[(k - 1) % 2], n, out, st))
        out.zero_(); st.fill_(1)
        ta.start(); tb.start(); ta.join(); tb.join()
        assert not errors, errors
        assert not st.any().item() and torch.equal(out, h_src)
    for t in (0, n // 2, n - 1):
        o, s = int(off[0][t]), int(sz[0][t])
        assert packed[0][o:o + s].numpy().tobytes() == oracle().encode(tiles[t], mode=MODE_FTL)
    ep.close(); dp.close()


@pytest.mark.parametrize("shape,dt,kw", [
    ((1, 1024, 1024, 3), np.uint8, dict(mode=MODE_FTL)),          # one large tile: 32 parts
    ((2, 517, 1030, 1), np.uint16, dict(mode=MODE_BASE)),         # ragged size, last block row shifted up
    ((3, 300, 260, 4), np.int16, dict(mode=MODE_BASE, quanta=3)), # quantised, derived bands
    ((1, 8, 2048, 2), np.uint8, dict(mode=MODE_BASE, cband=[1, 1])),  # two blocks to a row: the neighbour look-ups wrap rows
    ((1, 4, 512, 1), np.int32, dict(mode=MODE_FTL)),              # one block to a row
    ((2, 640, 400, 3), np.uint8, dict(mode=6)),                   # BASE + RLE over stitched parts
])
def test_batch_large_tiles_in_parts_match_oracle(shape, dt, kw):
    """Few large tiles are coded by several CTAs each and stitched (encode_kernel parts + stitch_kernel): byte identical
    to the oracle, and decodable."""
    torch = torch_mod()
    n, w, h, b = shape
    tiles = synth_tiles(n, w, h, b, dt)
    cfg, dst, sizes, status = encode_tiles(tiles, **kw)
    sz, dst_h = sizes.cpu().numpy(), dst.cpu().numpy()
    for t in range(n):
        want = oracle().encode(tiles[t], **kw)
        assert int(sz[t]) == len(want) and dst_h[t, :sz[t]].tobytes() == want, "tile %d differs from the oracle" % t
    offsets = torch.arange(n, device="cuda", dtype=torch.int64) * dst.stride(0)
    out, st = q.decode_batch(cfg, dst, offsets, sizes, n)
    torch.cuda.synchronize()
    assert not st.cpu().numpy().any()
    if kw.get("quanta", 1) == 1:
        assert np.array_equal(out.cpu().numpy().view(dt).reshape(tiles.shape), tiles)
    # incompressible content: the parts do not pay, the tile is stored
    noise = np.random.default_rng(3).integers(0, 256, (1, 256, 512, 3), dtype=np.uint8)
    cfg2, dst2, sizes2, _ = encode_tiles(noise, mode=MODE_BASE)
    want = oracle().encode(noise[0], mode=MODE_BASE)
    assert dst2.cpu().numpy()[0, :int(sizes2[0])].tobytes() == want and want[10] == 255


@pytest.mark.parametrize("shape,dt,kw", [
    ((6, 512, 512, 3), np.uint8, dict(mode=MODE_FTL)),
    ((3, 300, 200, 4), np.uint16, dict(mode=MODE_BASE, cband=[1, 1, 1, 3])),
    ((3, 128, 128, 1), np.int32, dict(mode=MODE_CF_H, quanta=5)),
    ((2, 256, 256, 3), np.uint8, dict(mode=MODE_BEST)),            # RLE mode: measured on streams made in scratch
    ((1, 1024, 2048, 3), np.uint8, dict(mode=MODE_BASE)),          # one large tile: in parts
    ((1, 1024, 1024, 1), np.uint16, dict(mode=MODE_CF_H)),         # BEST without RLE, in parts
    ((4, 3, 50, 2), np.uint8, dict(mode=MODE_FTL)),                # narrow images go through the reorder
])
def test_size_only_pass_matches_encode(shape, dt, kw):
    """qb3cu_encoded_size_batch (the pass behind cqb3cu -m x, cqb3.cpp:561-586): the sizes of the streams without the
    streams -- equal to what the encode reports and to the oracle's stream lengths, stored fallback included."""
    torch = torch_mod()
    n, w, h, b = shape
    tiles = synth_tiles(n, w, h, b, dt)
    tiles[n - 1] = np.random.default_rng(9).integers(0, 1 << 8 * np.dtype(dt).itemsize, tiles[0].shape).astype(np.uint64).astype(tiles.dtype) \
        if np.dtype(dt).kind == "u" else tiles[n - 1]     # noise: stored
    cfg, dst, sizes, status = encode_tiles(tiles, **kw)
    src = torch.from_numpy(tiles.view(np.uint8).reshape(n, -1)).cuda()
    only = q.encoded_size_batch(cfg, src, n)
    torch.cuda.synchronize()
    assert np.array_equal(only.cpu().numpy(), sizes.cpu().numpy())
    for t in (0, n - 1):
        assert int(only[t]) == len(oracle().encode(tiles[t], **kw))


def _best_parts_content(kind, w, h, b, dt):
    """Content whose BEST coding leans on the band's last written factor across the parts of a tile."""
    base = synth_tiles(1, w, h, b, np.uint8)[0].astype(np.uint64)
    yy = np.arange(h, dtype=np.uint64)[:, None, None]
    if kind == "synth":
        v = synth_tiles(1, w, h, b, dt)[0]
        return v
    if kind == "cf5":            # one factor everywhere: every part starts by meeting the factor that comes in
        v = base * np.uint64(5)
    elif kind == "cfbands":      # the factor changes every 40 rows, out of step with the parts
        f = np.array([5, 5, 3, 7, 5, 2, 2, 9, 3, 3], dtype=np.uint64)[(yy // np.uint64(40)) % np.uint64(10)]
        v = (base >> np.uint64(2)) * f
    elif kind == "cfrare":       # plain data with a stripe of factor 6 now and then (most parts never meet a factor)
        v = base.copy()
        stripe = ((yy // np.uint64(4)) % np.uint64(37)) == 0
        v = np.where(stripe, (base >> np.uint64(3)) * np.uint64(6), v)
    elif kind == "fewcf":        # few distinct values, all multiples of 3: index groups compete with factor groups
        rng = np.random.default_rng(11)
        pal = np.array([0, 3, 6, 30, 33, 96], dtype=np.uint64)
        v = pal[rng.integers(0, 6, size=(h, w, b))]
        v[h // 3: h // 3 + 50] = pal[rng.integers(0, 6, size=(50, w, b))] * np.uint64(2)
    else:
        raise ValueError(kind)
    dt = np.dtype(dt)
    udt = np.dtype("uint%d" % (8 * dt.itemsize))
    return np.ascontiguousarray(v.astype(udt)).view(dt)


@pytest.mark.parametrize("kind,shape,dt,kw", [
    ("synth", (1024, 1024, 3), np.uint8, dict(mode=MODE_BEST)),
    ("cf5", (512, 768, 3), np.uint8, dict(mode=MODE_BEST)),
    ("cf5", (260, 517, 2), np.uint16, dict(mode=MODE_CF_H, cband=[1, 1])),
    ("cfbands", (512, 1024, 3), np.uint8, dict(mode=MODE_BEST)),
    ("cfbands", (300, 1000, 1), np.int16, dict(mode=MODE_CF)),
    ("cfrare", (512, 2048, 3), np.uint8, dict(mode=MODE_BEST)),
    ("fewcf", (256, 1536, 2), np.uint8, dict(mode=MODE_BEST)),
    ("cfbands", (128, 1024, 1), np.uint32, dict(mode=MODE_BEST)),
    ("cf5", (64, 640, 1), np.uint64, dict(mode=MODE_BEST)),
])
def test_best_tiles_in_parts_match_oracle(kind, shape, dt, kw):
    """BEST on few large tiles: parts coded without knowing the band's last written factor, the factors handed down by
    best_resolve_kernel, dependent parts coded again (EncArgs::best_pass), RLE by chunks afterwards: byte identical
    to the oracle, decodable, and the running state that comes back is the oracle's."""
    torch = torch_mod()
    w, h, b = shape
    tiles = np.stack([_best_parts_content(kind, w, h, b, dt), _best_parts_content("synth", w, h, b, dt)])
    cfg, dst, sizes, status = encode_tiles(tiles, **kw)
    sz, dst_h = sizes.cpu().numpy(), dst.cpu().numpy()
    for t in range(2):
        want = oracle().encode(tiles[t], **kw)
        assert int(sz[t]) == len(want) and dst_h[t, :sz[t]].tobytes() == want, "tile %d differs from the oracle" % t
    offsets = torch.arange(2, device="cuda", dtype=torch.int64) * dst.stride(0)
    out, st = q.decode_batch(cfg, dst, offsets, sizes, 2)
    torch.cuda.synchronize()
    assert not st.cpu().numpy().any()
    assert np.array_equal(out.cpu().numpy().view(dt).reshape(tiles.shape), tiles)


@pytest.mark.parametrize("kind,shape,dt,mode", [
    ("zeros", (512, 512, 3), np.uint8, MODE_BEST),        # one long run of zero bytes
    ("lownoise", (512, 512, 1), np.uint8, MODE_BEST),     # rungs 0..2: streams full of 00 and FF bytes
    ("lownoise", (700, 300, 2), np.uint16, MODE_RLE_H),
    ("steps", (512, 512, 3), np.uint8, MODE_BEST),
    ("halfzero", (1024, 512, 1), np.uint8, MODE_BEST),    # runs that start and end inside chunks
    ("halfzero", (2048, 2048, 1), np.uint8, MODE_RLE),    # in parts, then RLE
    ("ffpairs", (512, 256, 1), np.uint8, MODE_BEST),
    ("many", (256, 256, 1), np.uint8, MODE_BEST),         # a batch: one CTA per tile does all the steps (rle_kernel)
])
def test_rle_by_chunks_matches_oracle(kind, shape, dt, mode):
    """rle_kernel cuts a stream into chunks at bytes that are neither 00 nor FF and codes them independently: the bytes
    must be the serial transducer's (QB3encode.cpp:271-332), whatever falls on a chunk boundary."""
    torch = torch_mod()
    w, h, b = shape
    if kind in ("halfzero", "many"):
        rng = np.random.default_rng(5)
        v = rng.integers(0, 3, size=(h, w, b)).astype(dt)
        v[:, : w // 2] = 0
        v[h // 3: h // 3 + 64] = 0
        v[::17, ::5] = 7
    elif kind == "ffpairs":   # top half: every value four less than the one before it in coding order, which codes as
        v = np.zeros((h, w, b), dtype=dt)          # runs of one bits (thousands of FF FF pairs); bottom half: zeros
        cur = 0
        for by in range(h // 8):
            for bx in range(w // 4):
                for i in range(16):
                    n = (0x01548cd9aefb7623 >> (4 * (15 - i))) & 15
                    cur = (cur - 4) & 255
                    v[4 * by + (n >> 2), 4 * bx + (n & 3), 0] = cur
    else:
        v = content(kind, w, h, b, dt)
    tiles = np.stack([v, v[::-1].copy()] * (160 if kind == "many" else 1))
    n = len(tiles)
    cfg, dst, sizes, status = encode_tiles(tiles, mode=mode)
    sz, dst_h = sizes.cpu().numpy(), dst.cpu().numpy()
    for t in (0, 1, n - 2, n - 1):
        want = oracle().encode(tiles[t], mode=mode)
        assert int(sz[t]) == len(want) and dst_h[t, :sz[t]].tobytes() == want, "tile %d differs from the oracle (mode byte %d)" % (t, want[10])
    assert len(set(sz[0::2])) == 1 and len(set(sz[1::2])) == 1
    offsets = torch.arange(n, device="cuda", dtype=torch.int64) * dst.stride(0)
    out, st = q.decode_batch(cfg, dst, offsets, sizes, n)
    torch.cuda.synchronize()
    assert not st.cpu().numpy().any()
    assert np.array_equal(out.cpu().numpy().view(dt).reshape(tiles.shape), tiles)


def test_streams_the_reference_encoder_breaks():
    """In the common factor modes the reference drops a 64 bit group of more than 800 bits for an empty index group
    (QB3encode.h:704-708). The streams are the reference's byte for byte all the same -- down to a tile in parts whose
    parts hold no bits at all -- and decode like the reference decodes them: failure where it fails, and where it
    does not, the same (meaningless) pixels, which at a width that is not a multiple of four depends on the moved-back
    last block of a row being written after the one it overlaps. Found by tools/fuzz.py."""
    torch = torch_mod()
    O = oracle()
    # every group dropped: the stream is its headers; three parts, two of them empty
    tiles = np.stack([content("highrung", 71, 93, 2, np.int64, seed=s) for s in (3, 4, 5)])
    kw = dict(mode=MODE_CF, quanta=37)
    cfg, dst, sizes, status = encode_tiles(tiles, **kw)
    sz, d = sizes.cpu().numpy(), dst.cpu().numpy()
    for t in range(3):
        want = O.encode(tiles[t], **kw)
        assert len(want) < 64 and int(sz[t]) == len(want) and d[t, :sz[t]].tobytes() == want
    # some groups dropped, the decoder runs on regardless
    checked = 0
    for seed in range(1, 7):
        for dt, kw in ((np.int64, dict(mode=MODE_BEST, quanta=5, away=True)), (np.uint64, dict(mode=MODE_CF_H, quanta=5))):
            img = content("fewvals", 9, 62, 1, dt, seed=seed)
            cfg, dst, sizes, status = encode_tiles(img[None], **kw)
            want = O.encode(img, **kw)
            assert dst.cpu().numpy()[0, :int(sizes[0])].tobytes() == want
            offsets = torch.zeros(1, device="cuda", dtype=torch.int64)
            out, st = q.decode_batch(cfg, dst, offsets, sizes, 1)
            torch.cuda.synchronize()
            ref = O.decode(want)
            if ref is None:
                assert int(st[0]) != 0
            else:
                assert int(st[0]) == 0 and np.array_equal(out[0].cpu().numpy().view(dt).reshape(img.shape), ref)
                checked += not np.array_equal(ref, img)
    assert checked > 0   # at least one stream that decodes, and not to what went in


@pytest.mark.parametrize("dt,mode", [(np.uint8, MODE_BASE), (np.uint8, MODE_BEST), (np.uint16, MODE_BEST), (np.int32, MODE_BASE),
                                     (np.uint8, MODE_FTL), (np.uint64, MODE_BEST), (np.int16, MODE_BASE), (np.int8, 1)])
def test_damaged_streams_decode_like_the_oracle(dt, mode):
    """One flipped payload bit per stream: wherever the oracle (and the compiled reference, when it is there) still
    decodes, the device must produce the very same pixels; where it reports failure, so must the device. A damaged
    stream wanders through rung switches, factors and index tables no encoder would write."""
    torch = torch_mod()
    n, w, h, b = 48, 64, 64, 3
    kinds = [k for k in CONTENT_KINDS]
    tiles = np.stack([content(kinds[t % len(kinds)], w, h, b, dt, seed=100 + t) for t in range(n)])
    cfg = q.config(w, h, b, dtype_code(dt), mode=mode)
    rng = np.random.default_rng(2024)
    streams = []
    for t in range(n):
        s = bytearray(oracle().encode(tiles[t], mode=mode))
        hdr = oracle().info(bytes(s))["data_offset"]
        if s[10] != 255 and len(s) > hdr + 4:
            pos = int(rng.integers(hdr * 8, len(s) * 8))
            s[pos >> 3] ^= 1 << (pos & 7)
        streams.append(bytes(s))
    slot = max(len(s) for s in streams) + 64
    buf = np.zeros((n, slot), np.uint8)
    for t, s in enumerate(streams):
        buf[t, :len(s)] = np.frombuffer(s, np.uint8)
    d = torch.from_numpy(buf).cuda()
    sizes = torch.tensor([len(s) for s in streams], dtype=torch.int64, device="cuda")
    offsets = torch.arange(n, device="cuda", dtype=torch.int64) * slot
    out, st = q.decode_batch(cfg, d, offsets, sizes, n)
    torch.cuda.synchronize()
    out_h, st_h = out.cpu().numpy().view(dt).reshape(tiles.shape), st.cpu().numpy()
    ref = QB3Lib(REF_SO, 16) if have_ref() else None
    agree = 0
    for t in range(n):
        want = oracle().decode(streams[t])
        if ref is not None and want is not None:
            # What the oracle decodes, the reference decodes (its pixels differ only by its band map default).
            # The converse does not hold: in a damaged common factor group whose factor claims a rung of its own equal
            # to 0 the reference reads its table at index -1 (QB3decode.h:651) and carries on with whatever is there;
            # the oracle and the device report failure.
            assert ref.decode(streams[t]) is not None, "tile %d: the oracle decodes what the reference rejects" % t
        if want is None:
            assert st_h[t] != q.TILE_OK, "tile %d: the oracle rejects the stream, the device accepts it" % t
        else:
            assert st_h[t] == q.TILE_OK, "tile %d: the oracle decodes the stream, the device reports %d" % (t, st_h[t])
            assert np.array_equal(out_h[t], want), "tile %d decodes differently" % t
            agree += 1
    assert agree > 0


def test_many_caller_streams():
    """Every caller stream gets helper streams inside the decoder, remembered in a table that never forgets; more
    caller streams than the table holds (QB3.h handles come and go by the hundred) must still decode, on either kind of
    helper stream (scans on their own SMs for a lone tile, shared SMs inside the host pipeline)."""
    torch = torch_mod()
    n, w, h, b = 4, 64, 64, 3
    tiles = synth_tiles(n, w, h, b, np.uint8)
    cfg, dst, sizes, status = encode_tiles(tiles, mode=MODE_BASE)
    offsets = torch.arange(n, device="cuda", dtype=torch.int64) * dst.stride(0)
    streams = [torch.cuda.Stream() for _ in range(200)]   # kept alive: two hundred distinct handles
    outs = []
    for st in streams:
        with torch.cuda.stream(st):
            st.wait_stream(torch.cuda.current_stream())
            out, s1 = q.decode_batch(cfg, dst, offsets[:1], sizes[:1], 1, stream=st)
            outs.append((out, s1))
    torch.cuda.synchronize()
    for out, s1 in outs[::17]:
        assert not s1.cpu().numpy().any() and np.array_equal(out.cpu().numpy().reshape(1, h, w, b)[0], tiles[0])
    pipe = q.Pipe(cfg, 2, 2)
    packed = np.zeros(n * q.slot_bytes(cfg), np.uint8)
    off, sz = np.zeros(n, np.uint64), np.zeros(n, np.uint64)
    pipe.encode(tiles, n, packed, off, sz)
    back = np.zeros_like(tiles)
    stat = np.full(n, 9, np.uint32)
    pipe.decode(packed, off, sz, n, back, stat)
    assert not stat.any() and np.array_equal(back, tiles)
    pipe.close()


def test_arbitrary_scan_curves():
    """Streams whose "SC" chunk names a curve that is neither Hilbert nor Z (QB3decode.cpp:231-250) decode on every
    path -- the QB3.h wrapper, the batch entry point -- and the batch encoder writes them when asked to (cfg.order)."""
    P, O = product(), oracle()
    for order in (0x0123456789abcdef, 0xfedcba9876543210, 0x048c159d26ae37bf, 0x5a0f3c96e17d48b2):
        for (w, h, b, dt, mode, cb) in ((16, 12, 1, np.uint8, MODE_FTL, None), (64, 40, 3, np.uint8, MODE_BASE, None),
                                       (21, 9, 3, np.uint16, MODE_BASE, None), (12, 8, 2, np.int32, MODE_BEST, [0, 0]),
                                       (20, 16, 1, np.uint64, MODE_FTL, None)):
            n = 5
            tiles = np.stack([content("synth", w, h, b, dt, seed=t) for t in range(n)])
            want = [O.encode(tiles[t], mode=mode, order=order, cband=cb) for t in range(n)]
            assert np.array_equal(P.decode(want[0]), tiles[0])
            torch = torch_mod()
            cfg = q.config(w, h, b, dtype_code(dt), mode=mode, cband=cb)
            cfg.order = order
            src = torch.from_numpy(tiles.view(np.uint8).reshape(n, -1)).cuda()
            dst, sizes, status = q.encode_batch(cfg, src, n)
            offsets = torch.arange(n, device="cuda", dtype=torch.int64) * dst.stride(0)
            out, st = q.decode_batch(cfg, dst, offsets, sizes, n)
            torch.cuda.synchronize()
            assert not status.cpu().numpy().any() and not st.cpu().numpy().any()
            sizes_h, dst_h = sizes.cpu().numpy(), dst.cpu().numpy()
            assert [dst_h[t, :sizes_h[t]].tobytes() for t in range(n)] == want
            assert np.array_equal(out.cpu().numpy().view(tiles.dtype).reshape(tiles.shape), tiles)


def test_multi_device_entry_point():
    """qb3cu_multi_*: one call shards a host batch over several devices (here every device of the box, and -- so that
    the sharding runs on a single GPU box too -- device 0 named three times). Streams equal the oracle's, tile by tile,
    whatever the sharding; decode gives the tiles back."""
    torch = torch_mod()
    w, h, b, n = 96, 64, 3, 37
    tiles = synth_tiles(n, w, h, b, np.uint8)
    want = [oracle().encode(tiles[t], mode=MODE_BASE) for t in range(n)]
    cfg = q.config(w, h, b, 0, mode=MODE_BASE)
    for devices in (None, [0, 0, 0], [0]):
        m = q.Multi(cfg, devices, chunk_tiles=5, depth=2)
        assert m.ndevices == (torch.cuda.device_count() if devices is None else len(devices))
        packed = np.zeros(n * q.slot_bytes(cfg), np.uint8)
        offs, sizes = np.zeros(n, np.uint64), np.zeros(n, np.uint64)
        total = m.encode(tiles, n, packed, offs, sizes)
        assert total == sum((len(s) + 15) // 16 * 16 for s in want)
        assert [packed[int(o):int(o + l)].tobytes() for o, l in zip(offs, sizes)] == want
        out, status = np.zeros_like(tiles), np.ones(n, np.uint32)
        m.decode(packed, offs, sizes, n, out, status)
        assert not status.any() and np.array_equal(out, tiles)
        m.close()
