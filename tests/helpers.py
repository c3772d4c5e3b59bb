"""Shared test helpers: ctypes bindings for the oracle, the compiled reference (oracle/_ref)
and the product library, plus the deterministic integer tile generator of BASELINE.md section 3.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may touch oracle/.
"""
import ctypes as C
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_SO = os.path.join(ROOT, "oracle", "libqb3oracle.so")
REF_SO = os.path.join(ROOT, "oracle", "_ref", "libQB3ref.so")
REF256_SO = os.path.join(ROOT, "oracle", "_ref", "libQB3ref256.so")
REFBENCH_SO = os.path.join(ROOT, "oracle", "_ref", "libqb3refbench.so")
PRODUCT_SO = os.path.join(ROOT, "qb3_b200", "libQB3.so")

DTYPES = [np.uint8, np.int8, np.uint16, np.int16, np.uint32, np.int32, np.uint64, np.int64]
U8, I8, U16, I16, U32, I32, U64, I64 = range(8)
MODE_BASE_Z, MODE_CF, MODE_RLE, MODE_CF_RLE, MODE_BASE, MODE_CF_H, MODE_RLE_H, MODE_BEST, MODE_FTL = range(9)
MODE_STORED = 255
ZCURVE = 0x0145236789CDABEF
HILBERT = 0x01548CD9AEFB7623


def dtype_code(dt):
    return [np.dtype(d) for d in DTYPES].index(np.dtype(dt))


# --------------------------------------------------------------------------- QB3.h bindings

class QB3Lib:
    """ctypes binding of the 21 QB3.h functions; works for the reference builds and the product."""

    def __init__(self, path, maxbands):
        self.lib = L = C.CDLL(path)
        self.maxbands = maxbands
        if hasattr(L, "qb3cu_api_max_bands"):  # the product: 16 bands like the reference until a caller asks for more
            L.qb3cu_api_max_bands.restype, L.qb3cu_api_max_bands.argtypes = C.c_uint32, [C.c_uint32]
            L.qb3cu_api_max_bands(maxbands)
        vp, sz, u64 = C.c_void_p, C.c_size_t, C.c_uint64
        sig = {
            "qb3_create_encoder": (vp, [sz, sz, sz, C.c_int]),
            "qb3_destroy_encoder": (None, [vp]),
            "qb3_reset_encoder": (None, [vp]),
            "qb3_set_encoder_coreband": (C.c_bool, [vp, sz, C.POINTER(sz)]),
            "qb3_set_encoder_quanta": (C.c_bool, [vp, u64, C.c_bool]),
            "qb3_max_encoded_size": (sz, [vp]),
            "qb3_set_encoder_mode": (C.c_int, [vp, C.c_int]),
            "qb3_set_encoder_stride": (None, [vp, sz]),
            "qb3_encode": (sz, [vp, vp, vp]),
            "qb3_get_encoder_state": (C.c_int, [vp]),
            "qb3_read_start": (vp, [vp, sz, C.POINTER(sz)]),
            "qb3_read_info": (C.c_bool, [vp]),
            "qb3_read_data": (sz, [vp, vp]),
            "qb3_destroy_decoder": (None, [vp]),
            "qb3_decoded_size": (sz, [vp]),
            "qb3_get_type": (C.c_int, [vp]),
            "qb3_set_decoder_stride": (None, [vp, sz]),
            "qb3_get_mode": (C.c_int, [vp]),
            "qb3_get_quanta": (u64, [vp]),
            "qb3_get_order": (u64, [vp]),
            "qb3_get_coreband": (C.c_bool, [vp, C.POINTER(sz)]),
        }
        for name, (res, args) in sig.items():
            f = getattr(L, name)
            f.restype, f.argtypes = res, args
        self.symbols = list(sig)

    def encode(self, img, mode=MODE_FTL, cband=None, quanta=1, away=False, stride=None, reps=1, ret_handle=False):
        """img: (h, w, bands) array (or (h, stride) backing array when stride is given with shape=).
        Returns the stream bytes (or a list, one per repetition on the same handle, when reps > 1)."""
        img = np.ascontiguousarray(img)
        h, w, b = img.shape
        L = self.lib
        e = L.qb3_create_encoder(w, h, b, dtype_code(img.dtype))
        if not e:
            raise ValueError("qb3_create_encoder failed")
        try:
            if cband is not None:
                arr = (C.c_size_t * self.maxbands)(*cband)
                if not L.qb3_set_encoder_coreband(e, len(cband), arr):
                    raise ValueError("qb3_set_encoder_coreband failed")
            if quanta != 1:
                if not L.qb3_set_encoder_quanta(e, quanta, away):
                    raise ValueError("qb3_set_encoder_quanta failed")
            L.qb3_set_encoder_mode(e, mode)
            if stride is not None:
                L.qb3_set_encoder_stride(e, stride)
            cap = L.qb3_max_encoded_size(e)
            out = []
            for _ in range(reps):
                dst = np.zeros(cap + 64, dtype=np.uint8)
                n = L.qb3_encode(e, img.ctypes.data, dst.ctypes.data)
                if n == 0:
                    raise RuntimeError("qb3_encode failed, state %d" % L.qb3_get_encoder_state(e))
                out.append(dst[:n].tobytes())
            return out[0] if reps == 1 else out
        finally:
            L.qb3_destroy_encoder(e)

    def max_encoded_size(self, w, h, b, dt):
        e = self.lib.qb3_create_encoder(w, h, b, dtype_code(dt))
        n = self.lib.qb3_max_encoded_size(e)
        self.lib.qb3_destroy_encoder(e)
        return n

    def info(self, stream):
        buf = np.frombuffer(stream, dtype=np.uint8).copy()
        dims = (C.c_size_t * 3)()
        L = self.lib
        d = L.qb3_read_start(buf.ctypes.data, len(buf), dims)
        if not d:
            return None
        try:
            if not L.qb3_read_info(d):
                return None
            cb = (C.c_size_t * self.maxbands)()
            L.qb3_get_coreband(d, cb)
            return dict(w=dims[0], h=dims[1], bands=dims[2], type=L.qb3_get_type(d), mode=L.qb3_get_mode(d),
                        quanta=L.qb3_get_quanta(d), order=L.qb3_get_order(d), cband=list(cb[:dims[2]]),
                        size=L.qb3_decoded_size(d))
        finally:
            L.qb3_destroy_decoder(d)

    def decode(self, stream, stride=None):
        """Returns the decoded (h, w, bands) array, or None when the library reports failure."""
        buf = np.frombuffer(stream, dtype=np.uint8).copy()
        dims = (C.c_size_t * 3)()
        L = self.lib
        d = L.qb3_read_start(buf.ctypes.data, len(buf), dims)
        if not d:
            return None
        try:
            if not L.qb3_read_info(d):
                return None
            w, h, b = dims[0], dims[1], dims[2]
            dt = DTYPES[L.qb3_get_type(d)]
            if stride is None:
                out = np.zeros((h, w, b), dtype=dt)
            else:
                L.qb3_set_decoder_stride(d, stride)
                out = np.zeros((h, stride), dtype=dt)
            n = L.qb3_read_data(d, out.ctypes.data)
            if n == 0:
                return None
            return out
        finally:
            L.qb3_destroy_decoder(d)


_cache = {}


def ref():
    """The unmodified reference library (QB3_MAXBANDS 16)."""
    if "ref" not in _cache:
        _cache["ref"] = QB3Lib(REF_SO, 16)
    return _cache["ref"]


def ref256():
    """Reference built with QB3_MAXBANDS 256 and the small-image scope fix (oracle/Makefile)."""
    if "ref256" not in _cache:
        _cache["ref256"] = QB3Lib(REF256_SO, 256)
    return _cache["ref256"]


def have_ref():
    return os.path.exists(REF_SO) and os.path.exists(REF256_SO)


# --------------------------------------------------------------------------- oracle binding

class _OEnc(C.Structure):
    _fields_ = [("xsize", C.c_size_t), ("ysize", C.c_size_t), ("nbands", C.c_size_t), ("stride", C.c_size_t),
                ("order", C.c_uint64), ("quanta", C.c_uint64), ("away", C.c_int), ("mode", C.c_int),
                ("type", C.c_int), ("error", C.c_int), ("cband", C.c_uint8 * 256),
                ("prev", C.c_uint64 * 256), ("runbits", C.c_uint64 * 256), ("cf", C.c_uint64 * 256)]


class _OInfo(C.Structure):
    _fields_ = [("xsize", C.c_size_t), ("ysize", C.c_size_t), ("nbands", C.c_size_t),
                ("order", C.c_uint64), ("quanta", C.c_uint64), ("mode", C.c_int), ("type", C.c_int),
                ("has_cb", C.c_int), ("cband", C.c_uint8 * 256), ("data_offset", C.c_size_t)]


class Oracle:
    """The plain-C restatement, oracle/qb3_oracle.c."""

    def __init__(self, path=ORACLE_SO):
        self.lib = L = C.CDLL(path)
        sz, vp = C.c_size_t, C.c_void_p
        L.qb3o_init.restype, L.qb3o_init.argtypes = C.c_int, [C.POINTER(_OEnc), sz, sz, sz, C.c_int]
        L.qb3o_set_mode.restype, L.qb3o_set_mode.argtypes = C.c_int, [C.POINTER(_OEnc), C.c_int]
        L.qb3o_set_coreband.restype, L.qb3o_set_coreband.argtypes = C.c_int, [C.POINTER(_OEnc), sz, C.POINTER(sz)]
        L.qb3o_reset.restype, L.qb3o_reset.argtypes = None, [C.POINTER(_OEnc)]
        L.qb3o_max_encoded_size.restype, L.qb3o_max_encoded_size.argtypes = sz, [C.POINTER(_OEnc)]
        L.qb3o_encode.restype, L.qb3o_encode.argtypes = sz, [C.POINTER(_OEnc), vp, vp]
        L.qb3o_read_info.restype, L.qb3o_read_info.argtypes = C.c_int, [vp, sz, C.POINTER(_OInfo)]
        L.qb3o_decode.restype, L.qb3o_decode.argtypes = sz, [vp, sz, vp, sz, C.c_int]
        for n in ("qb3o_crg", "qb3o_drg", "qb3o_csw", "qb3o_dsw"):
            f = getattr(L, n)
            f.restype, f.argtypes = C.c_uint16, [C.c_uint, C.c_uint]
        L.qb3o_signal.restype, L.qb3o_signal.argtypes = C.c_uint16, [C.c_uint]

    def encode(self, img, mode=MODE_FTL, cband=None, quanta=1, away=False, stride=None, reps=1, order=None):
        img = np.ascontiguousarray(img)
        h, w, b = img.shape
        e = _OEnc()
        if self.lib.qb3o_init(C.byref(e), w, h, b, dtype_code(img.dtype)):
            raise ValueError("qb3o_init failed")
        if cband is not None:
            arr = (C.c_size_t * 256)(*cband)
            if not self.lib.qb3o_set_coreband(C.byref(e), len(cband), arr):
                raise ValueError("coreband")
        e.quanta, e.away = quanta, int(away)
        self.lib.qb3o_set_mode(C.byref(e), mode)
        if order is not None:  # an arbitrary 4x4 scan curve, written as an "SC" chunk (QB3encode.cpp:246-252)
            e.order = order
        if stride is not None:
            e.stride = stride
        cap = self.lib.qb3o_max_encoded_size(C.byref(e))
        out = []
        for _ in range(reps):
            dst = np.zeros(cap + 64, dtype=np.uint8)
            n = self.lib.qb3o_encode(C.byref(e), img.ctypes.data, dst.ctypes.data)
            if n == 0:
                raise RuntimeError("qb3o_encode failed")
            out.append(dst[:n].tobytes())
        return out[0] if reps == 1 else out

    def max_encoded_size(self, w, h, b, dt):
        e = _OEnc()
        self.lib.qb3o_init(C.byref(e), w, h, b, dtype_code(dt))
        return self.lib.qb3o_max_encoded_size(C.byref(e))

    def info(self, stream):
        buf = np.frombuffer(stream, dtype=np.uint8).copy()
        o = _OInfo()
        if self.lib.qb3o_read_info(buf.ctypes.data, len(buf), C.byref(o)):
            return None
        return dict(w=o.xsize, h=o.ysize, bands=o.nbands, type=o.type, mode=o.mode, quanta=o.quanta,
                    order=o.order, cband=list(o.cband[:o.nbands]), has_cb=bool(o.has_cb), data_offset=o.data_offset)

    def decode(self, stream, stride=None, identity_default=True):
        i = self.info(stream)
        if i is None:
            return None
        buf = np.frombuffer(stream, dtype=np.uint8).copy()
        dt = DTYPES[i["type"]]
        out = np.zeros((i["h"], i["w"], i["bands"]) if stride is None else (i["h"], stride), dtype=dt)
        n = self.lib.qb3o_decode(buf.ctypes.data, len(buf), out.ctypes.data, stride or 0, int(identity_default))
        return out if n else None


def oracle():
    if "oracle" not in _cache:
        _cache["oracle"] = Oracle()
    return _cache["oracle"]


# --------------------------------------------------------------------------- synthetic tiles

_M64 = np.uint64(0xFFFFFFFFFFFFFFFF)


def splitmix64(x):
    """splitmix64 finaliser over a uint64 array."""
    with np.errstate(over="ignore"):
        z = (x + np.uint64(0x9E3779B97F4A7C15)) & _M64
        z = ((z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)) & _M64
        z = ((z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)) & _M64
        return z ^ (z >> np.uint64(31))


NOISE_BITS = {1: 3, 2: 6, 4: 8, 8: 10}


def synth_tiles(ntiles, w, h, bands, dt, seed=12345, nb=None, t0=0):
    """BASELINE.md section 3 generator: smooth triangles + per band offset + nb bits of hash noise.
    Returns (ntiles, h, w, bands) of dtype dt. Integer only, so the device generator matches bit for bit."""
    dt = np.dtype(dt)
    bits = dt.itemsize * 8
    if nb is None:
        nb = NOISE_BITS[dt.itemsize]
    A = (1 << 40) if bits == 64 else (1 << (bits - 1)) - 1
    t = np.arange(t0, t0 + ntiles, dtype=np.uint64)[:, None, None, None]
    y = np.arange(h, dtype=np.uint64)[None, :, None, None]
    x = np.arange(w, dtype=np.uint64)[None, None, :, None]
    c = np.arange(bands, dtype=np.uint64)[None, None, None, :]

    def tri(u, P, a):
        m = u % np.uint64(P)
        return np.minimum(m, np.uint64(P) - m) * np.uint64(a) // np.uint64(P // 2)

    with np.errstate(over="ignore"):
        v = tri(x + np.uint64(37) * t, 211, A // 2) + tri(y + np.uint64(91) * t, 157, A // 2)
        v = v + (c * np.uint64(A)) // np.uint64(8 * bands)
        idx = ((t * np.uint64(h) + y) * np.uint64(w) + x) * np.uint64(bands) + c
        noise = splitmix64(np.uint64(seed) ^ idx) & np.uint64((1 << nb) - 1) if nb else np.uint64(0)
        v = (v + noise) & _M64
    return v.astype(np.dtype("uint%d" % bits)).view(dt) if dt.kind == "i" else v.astype(dt)


def content(kind, w, h, bands, dt, seed=1):
    """Coverage content kinds for parity tests; returns (h, w, bands)."""
    dt = np.dtype(dt)
    bits = dt.itemsize * 8
    udt = np.dtype("uint%d" % bits)
    rng = np.random.default_rng(seed)
    yy, xx, cc = np.meshgrid(np.arange(h), np.arange(w), np.arange(bands), indexing="ij")
    if kind == "zeros":
        v = np.zeros((h, w, bands), dtype=udt)
    elif kind == "ramp":
        v = (xx + 2 * yy + 10 * cc).astype(np.uint64).astype(udt)
    elif kind == "synth":
        v = synth_tiles(1, w, h, bands, udt, seed=seed)[0]
    elif kind == "noise":
        v = rng.integers(0, 1 << min(bits, 63), size=(h, w, bands), dtype=np.uint64).astype(udt)
        if bits == 64:
            v = v | (rng.integers(0, 2, size=v.shape, dtype=np.uint64) << np.uint64(63))
    elif kind == "lownoise":
        v = (rng.integers(0, 4, size=(h, w, bands), dtype=np.uint64)).astype(udt)
    elif kind == "signed":
        s = (xx.astype(np.int64) - yy * 3 + rng.integers(-5, 6, size=xx.shape))
        v = s.astype(np.int64).view(np.uint64).astype(udt)
    elif kind == "cf5":
        v = (synth_tiles(1, w, h, bands, np.uint8, seed=seed)[0].astype(np.uint64) * np.uint64(5)).astype(udt)
    elif kind == "cfshift":
        k = max(bits - 8, 0)
        v = (synth_tiles(1, w, h, bands, np.uint8, seed=seed)[0].astype(np.uint64) << np.uint64(k)).astype(udt)
    elif kind == "fewvals":
        pal = rng.integers(0, 1 << min(bits, 63), size=5, dtype=np.uint64)
        v = pal[rng.integers(0, 5, size=(h, w, bands))].astype(udt)
    elif kind == "highrung":
        v = rng.integers(0, 1 << min(bits, 63), size=(h, w, bands), dtype=np.uint64).astype(udt)
        v[::2, ::3] = 0
    elif kind == "steps":
        v = ((xx // 3 + yy // 5) * 16 + rng.integers(0, 2, size=xx.shape)).astype(np.uint64).astype(udt)
    else:
        raise ValueError(kind)
    return np.ascontiguousarray(v).view(dt) if dt.kind == "i" else np.ascontiguousarray(v.astype(dt))


CONTENT_KINDS = ["zeros", "ramp", "synth", "noise", "lownoise", "signed", "cf5", "cfshift", "fewvals", "highrung", "steps"]


# --------------------------------------------------------------------------- golden vectors

def golden_cases():
    """tests/golden/golden.json (made by tests/golden/make_golden.py from the reference library)."""
    import json
    with open(os.path.join(ROOT, "tests", "golden", "golden.json")) as f:
        return json.load(f)["cases"]


def golden_image(case):
    dt = np.dtype(case["dtype"])
    if case["kind"] == "small":
        return np.frombuffer(bytes.fromhex(case["pixels"]), dtype=dt).reshape(case["h"], case["w"], case["bands"]).copy()
    return synth_tiles(1, case["w"], case["h"], case["bands"], dt, t0=case["tile"])[0]


def golden_kwargs(case):
    return {k: case[k] for k in ("mode", "cband", "quanta", "away") if k in case}


def golden_check_stream(case, stream):
    import hashlib
    if case["kind"] == "small":
        assert stream.hex() == case["stream"], case["name"]
    else:
        assert len(stream) == case["length"], case["name"]
        assert hashlib.sha256(stream).hexdigest() == case["sha256"], case["name"]
