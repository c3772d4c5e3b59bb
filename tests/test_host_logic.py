"""CPU-only checks of the product library (qb3_b200/libQB3.so): it loads, exports every symbol the
headers declare, and its host-side logic (closed-form code tables, header bytes, option setters, header
parser) agrees with the oracle / the reference. No kernel is launched here."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import qb3_b200 as q
from helpers import (DTYPES, MODE_BASE, MODE_BEST, MODE_FTL, PRODUCT_SO, QB3Lib, content, golden_cases, have_ref,
                     oracle, ref, ref256)

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def product():
    return QB3Lib(PRODUCT_SO, 256)


def declared_symbols():
    names = []
    for h in ("QB3.h", "qb3cu.h"):
        text = open(os.path.join(ROOT, "include", h)).read()
        names += re.findall(r"LIBQB3_EXPORT\s+[\w\s\*]+?\b(qb3\w+)\s*\(", text)
    return names


def test_library_exports_every_declared_symbol():
    L = q.lib()
    names = declared_symbols()
    assert len([n for n in names if not n.startswith("qb3cu_")]) == 21  # the QB3.h API, SURVEY 8b
    for n in names:
        assert hasattr(L, n), n


def test_no_test_probes_in_the_product_library():
    assert not [n for n in os.popen("nm -D --defined-only " + PRODUCT_SO).read().split() if "debug" in n]


def test_closed_form_tables_equal_oracle_tables():
    L, O = q.testing_lib(), oracle().lib
    for f in (L.qb3cu_debug_cs_entry, L.qb3cu_debug_ds_entry, L.qb3cu_debug_code, L.qb3cu_debug_decode, L.qb3cu_debug_cs_signal):
        f.restype = C.c_uint32
    for U in (3, 4, 5, 6):
        assert L.qb3cu_debug_cs_signal(U) == O.qb3o_signal(U)
        for d in range(1 << U):
            assert L.qb3cu_debug_cs_entry(U, d) == O.qb3o_csw(U, d)
        for x in range(1 << (U + 1)):
            assert L.qb3cu_debug_ds_entry(U, x) == O.qb3o_dsw(U, x)
    for r in range(0, 11):
        for v in range(2 ** (r + 1)):
            assert L.qb3cu_debug_code(r, v, 0) == O.qb3o_crg(r, v), (r, v)
        for x in range(2 ** (r + 2)):
            assert L.qb3cu_debug_decode(r, x, 0) == O.qb3o_drg(r, x), (r, x)
    # group values: rungs 1 and 2 are swapped too (inline tables of the reference, QB3encode.h:185-197)
    assert [L.qb3cu_debug_code(1, v, 1) & 0xFFF for v in range(4)] == [0, 3, 1, 7]
    assert [L.qb3cu_debug_code(1, v, 1) >> 12 for v in range(4)] == [1, 3, 2, 3]
    assert [L.qb3cu_debug_code(2, v, 1) & 0xFFF for v in range(8)] == [0, 2, 1, 3, 5, 7, 11, 15]
    assert [L.qb3cu_debug_code(2, v, 1) >> 12 for v in range(8)] == [2, 2, 3, 4, 3, 4, 4, 4]


def test_step_rule():
    L = q.testing_lib()
    for M in range(1, 1 << 16, 97):
        n = bin(M).count("1")
        is_step = (M & (M + 1)) == 0
        assert L.qb3cu_debug_step(M, 0) == (n - 1 if is_step else -1)
        assert L.qb3cu_debug_step(M, 1) == (n if is_step and M != 0xFFFF else -1)
    assert L.qb3cu_debug_step(0, 1) == 0 and L.qb3cu_debug_step(0xFFFF, 0) == 15


def test_header_bytes_match_oracle_streams():
    L = q.testing_lib()
    out = (C.c_uint8 * 320)()
    for (w, h, b, dt, kw) in [(8, 8, 1, np.uint8, {}), (17, 9, 3, np.uint16, {}), (12, 8, 5, np.int32, dict(cband=[2, 2, 2, 2, 4])),
                              (9, 7, 1, np.int32, dict(quanta=3)), (8, 8, 2, np.uint64, dict(quanta=70000, mode=MODE_BASE)),
                              (16, 16, 3, np.uint8, dict(mode=1)), (12, 8, 200, np.uint16, dict(cband=[0] * 200))]:
        img = content("ramp", w, h, b, dt)
        s = oracle().encode(img, **kw)
        mode_byte = s[10]
        cfg = q.config(w, h, b, DTYPES.index(dt) if dt in DTYPES else [np.dtype(d) for d in DTYPES].index(np.dtype(dt)),
                       mode=kw.get("mode", MODE_FTL), quanta=kw.get("quanta", 1))
        if "cband" in kw:
            arr = (C.c_size_t * 256)(*kw["cband"])
            e = product().lib.qb3_create_encoder(w, h, b, cfg.dtype)
            product().lib.qb3_set_encoder_coreband(e, b, arr)
            product().lib.qb3_destroy_encoder(e)
            for i in range(b):
                cfg.cband[i] = arr[i]
        n = L.qb3cu_debug_headers(C.byref(cfg), mode_byte, out)
        off = oracle().info(s)["data_offset"]
        assert n == off and bytes(out[:n]) == s[:off]


def test_max_encoded_size_and_slot():
    for (w, h, b, dt) in [(512, 512, 3, 0), (513, 511, 1, 5), (7, 5, 16, 6), (65536, 3, 1, 2), (1, 1, 1, 0), (64, 64, 256, 3)]:
        cfg = q.config(w, h, b, dt)
        assert q.max_encoded_size(cfg) == oracle().max_encoded_size(w, h, b, DTYPES[dt])
        assert q.slot_bytes(cfg) % 16 == 0 and q.slot_bytes(cfg) >= q.max_encoded_size(cfg) + 16
    assert q.max_encoded_size(q.config(512, 512, 3, 0)) == 891904


def test_create_encoder_validation():
    P = product().lib
    for bad in [(0, 4, 1, 0), (65537, 4, 1, 0), (4, 0, 1, 0), (4, 65537, 1, 0), (4, 4, 0, 0), (4, 4, 257, 0), (4, 4, 1, 8)]:
        assert not P.qb3_create_encoder(*bad)
    e = P.qb3_create_encoder(65536, 65536, 256, 7)
    assert e
    P.qb3_destroy_encoder(e)


@pytest.mark.skipif(not have_ref(), reason="oracle/_ref not built")
def test_setters_match_reference():
    P, R = product(), ref256()
    rng = np.random.default_rng(5)
    for bands in (1, 2, 3, 4, 7, 16):
        for _ in range(20):
            cb = [int(x) for x in rng.integers(0, bands + 2, size=bands)]
            res = []
            for lib in (P, R):
                e = lib.lib.qb3_create_encoder(8, 8, bands, 0)
                arr = (C.c_size_t * 256)(*cb)
                ok = lib.lib.qb3_set_encoder_coreband(e, bands, arr)
                bad = lib.lib.qb3_set_encoder_coreband(e, bands + 1, arr)
                res.append((ok, bad, list(arr[:bands])))
                lib.lib.qb3_destroy_encoder(e)
            assert res[0] == res[1], cb
    for dt in range(8):
        for qv in (0, 1, 2, 127, 128, 255, 256, 32767, 32768, 65535, 65536, 2 ** 31 - 1, 2 ** 31, 2 ** 32 - 1, 2 ** 32,
                   2 ** 63 - 1, 2 ** 63):
            res = []
            for lib in (P, R):
                e = lib.lib.qb3_create_encoder(8, 8, 1, dt)
                res.append(lib.lib.qb3_set_encoder_quanta(e, qv, False))
                lib.lib.qb3_destroy_encoder(e)
            assert res[0] == res[1], (dt, qv)
    for mode in (-1, 0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 255):
        res = []
        for lib in (P, R):
            e = lib.lib.qb3_create_encoder(8, 8, 1, 0)
            res.append((lib.lib.qb3_set_encoder_mode(e, 4), lib.lib.qb3_set_encoder_mode(e, mode)))
            lib.lib.qb3_destroy_encoder(e)
        assert res[0] == res[1], mode


def test_header_parser_on_golden_streams():
    P = product()
    os.environ["QB3_REF_COMPAT"] = "1"
    try:
        for case in golden_cases():
            if case["kind"] != "small":
                continue
            s = bytes.fromhex(case["stream"])
            mine, want = P.info(s), oracle().info(s)
            if want is None or len(s) < 15:
                assert mine is None, case["name"]
                continue
            assert mine is not None, case["name"]
            for k in ("w", "h", "bands", "type", "mode"):
                assert mine[k] == want[k], (case["name"], k)
            assert mine["quanta"] == (want["quanta"] or 0) or (want["quanta"] == 0 and mine["quanta"] in (0, 1))
            assert mine["cband"] == want["cband"], case["name"]
    finally:
        del os.environ["QB3_REF_COMPAT"]


@pytest.mark.skipif(not have_ref(), reason="oracle/_ref not built")
def test_header_parser_matches_reference_getters():
    P, R = product(), ref256()
    os.environ["QB3_REF_COMPAT"] = "1"
    try:
        for case in golden_cases()[:120]:
            if case["kind"] != "small":
                continue
            s = bytes.fromhex(case["stream"])
            assert P.info(s) == R.info(s), case["name"]
        # malformed inputs: same verdict as the reference
        base = bytes.fromhex(golden_cases()[0]["stream"])
        for mut in (base[:10], base[:14], b"QB4" + base[3:], base[:10] + bytes([9]) + base[11:], base[:11] + b"\x80" + base[12:],
                    base[:11] + b"ZZ\x01\x00\x00" + base[11:], base[:9] + bytes([8]) + base[10:]):
            assert (P.info(mut) is None) == (R.info(mut) is None)
    finally:
        del os.environ["QB3_REF_COMPAT"]


def test_no_cpu_fallback():
    """Without a CUDA device the codec entry points must fail loudly instead of computing on the host."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    P = product().lib
    img = np.zeros((8, 8, 1), np.uint8)
    e = P.qb3_create_encoder(8, 8, 1, 0)
    dst = np.zeros(P.qb3_max_encoded_size(e), np.uint8)
    assert P.qb3_encode(e, img.ctypes.data, dst.ctypes.data) == 0
    assert P.qb3_get_encoder_state(e) == 255  # QB3E_LIBERR
    P.qb3_destroy_encoder(e)
    s = np.frombuffer(bytes.fromhex(golden_cases()[0]["stream"]), np.uint8).copy()
    dims = (C.c_size_t * 3)()
    d = P.qb3_read_start(s.ctypes.data, len(s), dims)
    assert d and P.qb3_read_info(d)
    out = np.zeros(64, np.uint8)
    assert P.qb3_read_data(d, out.ctypes.data) == 0
    P.qb3_destroy_decoder(d)


def test_pipe_needs_a_device_and_valid_arguments():
    """The host pipeline has no CPU path either: without a device qb3cu_pipe_create returns NULL; bad arguments are
    refused before anything touches CUDA."""
    import qb3_b200 as q
    L = q.lib()
    cfg = q.config(64, 64, 3, q.U8)
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:  # noqa: BLE001
        has_gpu = False
    if not has_gpu:
        with pytest.raises(RuntimeError):
            q.Pipe(cfg)
        assert not L.qb3cu_host_alloc(1 << 20)
    assert not L.qb3cu_pipe_create(None, 0, 0)
    bad = q.config(64, 64, 3, q.U8)
    bad.width = 0
    assert not L.qb3cu_pipe_create(C.byref(bad), 0, 0)
    assert not L.qb3cu_pipe_create(C.byref(cfg), 0, -1)
    assert L.qb3cu_pipe_encode(None, None, 0, None, 0, None, None, None, 0) == 1      # QB3CU_ERR_PARAM
    assert L.qb3cu_pipe_decode(None, None, None, None, None, 0, None, 0, 0) == 1
    L.qb3cu_pipe_destroy(None)
    L.qb3cu_host_free(None)


def test_band_limit_is_the_references_until_a_caller_raises_it():
    """QB3_MAXBANDS is 16 in the reference's header and library (QB3.h:34, QB3encode.cpp:28-30, QB3decode.cpp:152):
    the QB3.h functions here refuse a 17th band the same way until qb3cu_api_max_bands() is told otherwise."""
    L = product().lib
    assert "#define QB3_MAXBANDS 16" in open(os.path.join(ROOT, "include", "QB3.h")).read()
    try:
        assert L.qb3cu_api_max_bands(16) == 16
        assert not L.qb3_create_encoder(8, 8, 17, 0)
        e = L.qb3_create_encoder(8, 8, 16, 0)
        assert e
        L.qb3_destroy_encoder(e)
        s17 = np.frombuffer(oracle().encode(content("ramp", 8, 8, 17, np.uint8)), np.uint8).copy()
        dims = (C.c_size_t * 3)()
        assert not L.qb3_read_start(s17.ctypes.data, s17.size, dims)
        assert L.qb3cu_api_max_bands(0) == 16 and L.qb3cu_api_max_bands(257) == 16  # out of range: unchanged
        assert L.qb3cu_api_max_bands(256) == 256
        e = L.qb3_create_encoder(8, 8, 17, 0)
        assert e
        L.qb3_destroy_encoder(e)
        d = L.qb3_read_start(s17.ctypes.data, s17.size, dims)
        assert d and dims[2] == 17
        L.qb3_destroy_decoder(d)
    finally:
        L.qb3cu_api_max_bands(256)
