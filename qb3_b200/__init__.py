"""qb3_b200 -- thin ctypes binding of libQB3.so (QB3.h API + qb3cu.h batched C ABI).

The product is the shared library built from qb3_b200/csrc (hand-written CUDA for sm_100a behind a C
ABI); this module only loads it and passes device pointers of torch tensors through. There is no
Python or CPU implementation of the codec here: if the library is missing, loading fails loudly.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libQB3.so")

MAXBANDS = 256
U8, I8, U16, I16, U32, I32, U64, I64 = range(8)
TYPESIZE = (1, 1, 2, 2, 4, 4, 8, 8)
M_BASE_Z, M_CF, M_RLE, M_CF_RLE, M_BASE, M_CF_H, M_RLE_H, M_BEST, M_FTL = range(9)
M_STORED = 255
TILE_OK, TILE_BAD_HEADER, TILE_CORRUPT, TILE_RLE_TOO_BIG = range(4)


class Config(C.Structure):
    """struct qb3cu_config (include/qb3cu.h)."""
    _fields_ = [("width", C.c_uint32), ("height", C.c_uint32), ("bands", C.c_uint32), ("dtype", C.c_uint32),
                ("mode", C.c_uint32), ("away", C.c_uint32), ("quanta", C.c_uint64), ("order", C.c_uint64),
                ("stride", C.c_uint64), ("cband", C.c_uint8 * MAXBANDS)]


_lib = None
_testing = None


def testing_lib():
    """libqb3cu_testing.so: the closed forms and the header writer as host functions, for the CPU-only tests."""
    global _testing
    if _testing is None:
        _testing = C.CDLL(os.path.join(_HERE, "libqb3cu_testing.so"))
    return _testing


def lib():
    """The loaded shared library; raises if it has not been built (see __graft_entry__.build)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError("%s is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                               "(there is no CPU fallback)" % LIB_PATH)
        L = C.CDLL(LIB_PATH)
        vp, sz, u64p, u32p = C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p
        cfgp = C.POINTER(Config)
        L.qb3cu_config_init.restype, L.qb3cu_config_init.argtypes = C.c_int, [cfgp, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32]
        L.qb3cu_max_encoded_size.restype, L.qb3cu_max_encoded_size.argtypes = sz, [cfgp]
        L.qb3cu_slot_bytes.restype, L.qb3cu_slot_bytes.argtypes = sz, [cfgp]
        L.qb3cu_encode_batch.restype = C.c_int
        L.qb3cu_encode_batch.argtypes = [cfgp, vp, sz, vp, sz, u64p, u32p, u64p, sz, vp]
        L.qb3cu_encoded_size_batch.restype = C.c_int
        L.qb3cu_encoded_size_batch.argtypes = [cfgp, vp, sz, u64p, sz, vp]
        L.qb3cu_decode_batch.restype = C.c_int
        L.qb3cu_decode_batch.argtypes = [cfgp, vp, u64p, u64p, vp, sz, u32p, C.c_int, sz, vp]
        L.qb3cu_pack_streams.restype = C.c_int
        L.qb3cu_pack_streams.argtypes = [vp, sz, u64p, vp, u64p, u64p, sz, vp]
        L.qb3cu_pipe_create.restype, L.qb3cu_pipe_create.argtypes = vp, [cfgp, sz, C.c_int]
        L.qb3cu_pipe_destroy.restype, L.qb3cu_pipe_destroy.argtypes = None, [vp]
        L.qb3cu_pipe_encode.restype = C.c_int
        L.qb3cu_pipe_encode.argtypes = [vp, vp, sz, vp, sz, u64p, u64p, u64p, sz]
        L.qb3cu_pipe_decode.restype = C.c_int
        L.qb3cu_pipe_decode.argtypes = [vp, vp, u64p, u64p, vp, sz, u32p, C.c_int, sz]
        L.qb3cu_multi_create.restype, L.qb3cu_multi_create.argtypes = vp, [cfgp, C.POINTER(C.c_int), C.c_int, sz, C.c_int]
        L.qb3cu_multi_destroy.restype, L.qb3cu_multi_destroy.argtypes = None, [vp]
        L.qb3cu_multi_devices.restype, L.qb3cu_multi_devices.argtypes = C.c_int, [vp]
        L.qb3cu_multi_encode.restype = C.c_int
        L.qb3cu_multi_encode.argtypes = [vp, vp, sz, vp, sz, u64p, u64p, u64p, sz]
        L.qb3cu_multi_decode.restype = C.c_int
        L.qb3cu_multi_decode.argtypes = [vp, vp, u64p, u64p, vp, sz, u32p, C.c_int, sz]
        L.qb3cu_host_alloc.restype, L.qb3cu_host_alloc.argtypes = vp, [sz]
        L.qb3cu_host_free.restype, L.qb3cu_host_free.argtypes = None, [vp]
        L.qb3cu_last_cuda_error.restype, L.qb3cu_last_cuda_error.argtypes = C.c_int, []
        L.qb3cu_kernel_launches.restype, L.qb3cu_kernel_launches.argtypes = C.c_uint64, []
        _lib = L
    return _lib


def config(width, height, bands, dtype, mode=M_FTL, cband=None, quanta=1, away=False, order=0, stride=0):
    """qb3cu_config with the reference defaults (FTL, identity or RGB band map) unless overridden."""
    cfg = Config()
    if lib().qb3cu_config_init(C.byref(cfg), width, height, bands, dtype) != 0:
        raise ValueError("bad geometry or type")
    cfg.mode, cfg.quanta, cfg.away, cfg.order, cfg.stride = mode, quanta, int(away), order, stride
    if cband is not None:
        if len(cband) != bands:
            raise ValueError("band map length")
        for i, b in enumerate(cband):
            cfg.cband[i] = b
    return cfg


def max_encoded_size(cfg):
    return lib().qb3cu_max_encoded_size(C.byref(cfg))


def slot_bytes(cfg):
    return lib().qb3cu_slot_bytes(C.byref(cfg))


def kernel_launches():
    return lib().qb3cu_kernel_launches()


def _check(rc, what):
    if rc != 0:
        raise RuntimeError("%s failed: rc=%d cuda=%d" % (what, rc, lib().qb3cu_last_cuda_error()))


def _stream_handle(stream):
    import torch
    if stream is None:
        stream = torch.cuda.current_stream()
    return C.c_void_p(stream.cuda_stream)


def encode_batch(cfg, src, ntiles, dst=None, sizes=None, status=None, state=None, tile_pitch=None, stream=None):
    """Encode ntiles tiles resident on the GPU. src: torch tensor holding the tiles back to back
    (tile_pitch bytes apart, default one compact tile). Returns (dst uint8 [ntiles, slot], sizes uint64-as-int64 [ntiles],
    status int32 [ntiles]); asynchronous on the given / current torch stream."""
    import torch
    slot = slot_bytes(cfg)
    dev = src.device
    if tile_pitch is None:
        tile_pitch = cfg.width * cfg.height * cfg.bands * TYPESIZE[cfg.dtype] if not cfg.stride else \
            cfg.stride * cfg.height * TYPESIZE[cfg.dtype]
    if dst is None:
        dst = torch.empty((ntiles, slot), dtype=torch.uint8, device=dev)
    if sizes is None:
        sizes = torch.empty((ntiles,), dtype=torch.int64, device=dev)
    if status is None:
        status = torch.empty((ntiles,), dtype=torch.int32, device=dev)
    rc = lib().qb3cu_encode_batch(C.byref(cfg), src.data_ptr(), tile_pitch, dst.data_ptr(), dst.stride(0),
                                  sizes.data_ptr(), status.data_ptr(), state.data_ptr() if state is not None else None,
                                  ntiles, _stream_handle(stream))
    _check(rc, "qb3cu_encode_batch")
    return dst, sizes, status


def encoded_size_batch(cfg, src, ntiles, sizes=None, tile_pitch=None, stream=None):
    """The stream sizes encode_batch would report, without making the streams (qb3cu_encoded_size_batch)."""
    import torch
    if tile_pitch is None:
        tile_pitch = cfg.width * cfg.height * cfg.bands * TYPESIZE[cfg.dtype] if not cfg.stride else \
            cfg.stride * cfg.height * TYPESIZE[cfg.dtype]
    if sizes is None:
        sizes = torch.empty((ntiles,), dtype=torch.int64, device=src.device)
    rc = lib().qb3cu_encoded_size_batch(C.byref(cfg), src.data_ptr(), tile_pitch, sizes.data_ptr(), ntiles, _stream_handle(stream))
    _check(rc, "qb3cu_encoded_size_batch")
    return sizes


def decode_batch(cfg, streams, offsets, lens, ntiles, out=None, status=None, ref_compat=False, tile_pitch=None,
                 out_dtype=None, stream=None):
    """Decode ntiles streams resident on the GPU. streams: uint8 tensor; offsets, lens: int64 tensors [ntiles] (bytes).
    Returns (out, status)."""
    import torch
    ts = TYPESIZE[cfg.dtype]
    if tile_pitch is None:
        tile_pitch = (cfg.stride if cfg.stride else cfg.width * cfg.bands) * cfg.height * ts
    dev = streams.device
    if out is None:
        out = torch.empty((ntiles, tile_pitch), dtype=torch.uint8, device=dev)
    if status is None:
        status = torch.empty((ntiles,), dtype=torch.int32, device=dev)
    rc = lib().qb3cu_decode_batch(C.byref(cfg), streams.data_ptr(), offsets.data_ptr(), lens.data_ptr(), out.data_ptr(),
                                  tile_pitch, status.data_ptr(), int(ref_compat), ntiles, _stream_handle(stream))
    _check(rc, "qb3cu_decode_batch")
    return out, status


def pack_streams(slots, sizes, ntiles, packed=None, offsets=None, total=None, stream=None):
    """Pack the streams of encode_batch back to back (16 byte aligned starts). Returns (packed, offsets, total);
    total is a 1-element int64 tensor on the device."""
    import torch
    dev = slots.device
    if packed is None:
        packed = torch.empty(slots.numel(), dtype=torch.uint8, device=dev)
    if offsets is None:
        offsets = torch.empty((ntiles,), dtype=torch.int64, device=dev)
    if total is None:
        total = torch.empty((1,), dtype=torch.int64, device=dev)
    rc = lib().qb3cu_pack_streams(slots.data_ptr(), slots.stride(0), sizes.data_ptr(), packed.data_ptr(),
                                  offsets.data_ptr(), total.data_ptr(), ntiles, _stream_handle(stream))
    _check(rc, "qb3cu_pack_streams")
    return packed, offsets, total


class Pipe:
    """qb3cu_pipe: batches held in HOST memory (numpy arrays or CPU torch tensors, ideally page locked) through the
    device in overlapped chunks. encode() returns (total bytes used in packed); decode() fills out and status."""

    def __init__(self, cfg, chunk_tiles=0, depth=0):
        self.cfg = cfg
        self.handle = lib().qb3cu_pipe_create(C.byref(cfg), chunk_tiles, depth)
        if not self.handle:
            raise RuntimeError("qb3cu_pipe_create failed: cuda=%d (there is no CPU fallback)" % lib().qb3cu_last_cuda_error())

    def close(self):
        h, self.handle = getattr(self, "handle", None), None
        if h and _lib is not None:  # at interpreter exit the module globals may be gone already
            _lib.qb3cu_pipe_destroy(h)

    __del__ = close

    @staticmethod
    def _ptr(a):
        return a.data_ptr() if hasattr(a, "data_ptr") else a.ctypes.data

    def encode(self, src, ntiles, packed, offsets, sizes, tile_pitch=None):
        """src: host tiles back to back (tile_pitch bytes apart); packed: host byte buffer; offsets, sizes: host
        uint64 / int64 arrays [ntiles]. Returns the number of bytes of packed used."""
        ts = TYPESIZE[self.cfg.dtype]
        if tile_pitch is None:
            tile_pitch = (self.cfg.stride if self.cfg.stride else self.cfg.width * self.cfg.bands) * self.cfg.height * ts
        total = C.c_uint64(0)
        cap = packed.numel() * packed.element_size() if hasattr(packed, "numel") else packed.nbytes
        rc = lib().qb3cu_pipe_encode(self.handle, self._ptr(src), tile_pitch, self._ptr(packed), cap, self._ptr(offsets),
                                     self._ptr(sizes), C.addressof(total), ntiles)
        _check(rc, "qb3cu_pipe_encode")
        return total.value

    def decode(self, packed, offsets, lens, ntiles, out, status, tile_pitch=None, ref_compat=False):
        ts = TYPESIZE[self.cfg.dtype]
        if tile_pitch is None:
            tile_pitch = (self.cfg.stride if self.cfg.stride else self.cfg.width * self.cfg.bands) * self.cfg.height * ts
        rc = lib().qb3cu_pipe_decode(self.handle, self._ptr(packed), self._ptr(offsets), self._ptr(lens), self._ptr(out),
                                     tile_pitch, self._ptr(status), int(ref_compat), ntiles)
        _check(rc, "qb3cu_pipe_decode")


class Multi(Pipe):
    """qb3cu_multi: the same two calls, the batch sharded over several devices from this one process (a host thread
    and a pipe per device inside the library). devices: list of device indices, None = all."""

    def __init__(self, cfg, devices=None, chunk_tiles=0, depth=0):
        self.cfg = cfg
        arr = (C.c_int * len(devices))(*devices) if devices else None
        self.handle = lib().qb3cu_multi_create(C.byref(cfg), arr, len(devices) if devices else 0, chunk_tiles, depth)
        if not self.handle:
            raise RuntimeError("qb3cu_multi_create failed: cuda=%d (there is no CPU fallback)" % lib().qb3cu_last_cuda_error())
        self.ndevices = lib().qb3cu_multi_devices(self.handle)

    def close(self):
        h, self.handle = getattr(self, "handle", None), None
        if h and _lib is not None:
            _lib.qb3cu_multi_destroy(h)

    __del__ = close

    def encode(self, src, ntiles, packed, offsets, sizes, tile_pitch=None):
        ts = TYPESIZE[self.cfg.dtype]
        if tile_pitch is None:
            tile_pitch = (self.cfg.stride if self.cfg.stride else self.cfg.width * self.cfg.bands) * self.cfg.height * ts
        total = C.c_uint64(0)
        cap = packed.numel() * packed.element_size() if hasattr(packed, "numel") else packed.nbytes
        rc = lib().qb3cu_multi_encode(self.handle, self._ptr(src), tile_pitch, self._ptr(packed), cap, self._ptr(offsets),
                                      self._ptr(sizes), C.addressof(total), ntiles)
        _check(rc, "qb3cu_multi_encode")
        return total.value

    def decode(self, packed, offsets, lens, ntiles, out, status, tile_pitch=None, ref_compat=False):
        ts = TYPESIZE[self.cfg.dtype]
        if tile_pitch is None:
            tile_pitch = (self.cfg.stride if self.cfg.stride else self.cfg.width * self.cfg.bands) * self.cfg.height * ts
        rc = lib().qb3cu_multi_decode(self.handle, self._ptr(packed), self._ptr(offsets), self._ptr(lens), self._ptr(out),
                                      tile_pitch, self._ptr(status), int(ref_compat), ntiles)
        _check(rc, "qb3cu_multi_decode")


def shard_range(ntiles, rank, world):
    """Contiguous tile range [begin, end) of one rank when a batch of ntiles independent tiles is split over world
    GPUs (no exchange step, every rank encodes / decodes its own range; only sizes go back to the host)."""
    if world < 1 or not 0 <= rank < world:
        raise ValueError("bad rank / world")
    return ntiles * rank // world, ntiles * (rank + 1) // world
