/*
 * qb3_pipe.cu -- the host buffer pipeline of include/qb3cu.h (qb3cu_pipe_*): host code only.
 *
 * The reference's callers hold images in host memory and run one qb3_encode / qb3_read_data per image
 * (cqb3.cpp:405-493, 276-323); a batch of tiles here goes through the device in chunks, each chunk on its own CUDA
 * stream with its own staging memory:
 *   encode:  pixels up -> qb3cu_encode_batch -> qb3cu_pack_streams -> index (sizes, offsets, total) down;
 *            the host reads the total and brings down exactly the bytes produced, behind the next chunk's upload
 *   decode:  the chunk's span of stream bytes and its index up -> qb3cu_decode_batch's kernels, packed onto as few
 *            SMs as hold the chunk's streams -> pixels down
 * The copy engines serve both directions at once and a chunk's kernels run beside other chunks' copies. The parse of a
 * chunk takes about 9 ms however few tiles it has (one serial walk per stream), so several chunks are kept in flight.
 */
#include <cstdlib>
#include <cstring>
#include <new>
#include <vector>

#include "qb3_device.cuh"

namespace qb3 {
int note_cuda(cudaError_t e);
int decode_batch_shared(const qb3cu_config *cfg, const void *d_streams, const uint64_t *d_offsets, const uint64_t *d_lens,
                        void *d_dst, size_t dst_tile_pitch, uint32_t *d_status, int ref_compat, size_t ntiles, void *stream);

static const uint32_t PIPE_TYPESIZE[8] = {1, 1, 2, 2, 4, 4, 8, 8};

/* device memory that only ever grows */
struct DevBuf {
    uint8_t *p = nullptr;
    size_t cap = 0;
    bool ensure(size_t n, bool headroom = false) /* headroom: sizes that vary from call to call should not reallocate */
    {
        if (n <= cap) return true;
        if (headroom) n += n / 4 + (1u << 16);
        if (p) cudaFree(p);
        p = nullptr; cap = 0;
        n = (n + 255) & ~(size_t)255;
        if (note_cuda(cudaMalloc(reinterpret_cast<void **>(&p), n)) != QB3CU_OK) return false;
        cap = n;
        return true;
    }
    void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
};

/* one chunk in flight */
struct Stage {
    cudaStream_t st = nullptr;
    cudaEvent_t index_ready = nullptr, uploaded = nullptr;
    DevBuf pix, slots, packed, meta;
    uint64_t *h_meta = nullptr; /* pinned: [sizes or lens | offsets | total] then the status words */
    size_t first = 0, n = 0;    /* the tiles it holds */
    bool busy = false;
};
} // namespace qb3

using namespace qb3;

struct qb3cu_pipe {
    qb3cu_config cfg;      /* as given */
    qb3cu_config dev_cfg;  /* what the kernels see: the same with the line stride of the device copy */
    int device;            /* the device the pipe was created on: every call runs there, whatever thread makes it */
    size_t chunk, depth;
    size_t tsize, line_bytes, tile_bytes, slot;
    size_t dev_pitch;      /* bytes between tiles in device staging memory */
    bool strided;          /* lines are further apart than their length: copied line by line */
    std::vector<Stage> stages;
};

namespace qb3 {

static size_t meta_words(size_t n) { return 2 * n + 2; }
static uint32_t *status_of(uint64_t *meta, size_t n) { return reinterpret_cast<uint32_t *>(meta + meta_words(n)); }
static size_t meta_bytes(size_t n) { return meta_words(n) * 8 + n * 4; }

/* tiles [first, first + n) between host memory (tile pitch hpitch) and a stage's device staging memory */
static cudaError_t copy_tiles(const qb3cu_pipe *p, void *dst, size_t dpitch, const void *src, size_t spitch, size_t n,
                              cudaMemcpyKind kind, cudaStream_t st)
{
    if (!p->strided) {
        if (dpitch == p->tile_bytes && spitch == p->tile_bytes) return cudaMemcpyAsync(dst, src, n * p->tile_bytes, kind, st);
        return cudaMemcpy2DAsync(dst, dpitch, src, spitch, p->tile_bytes, n, kind, st);
    }
    const size_t lpitch = (size_t)p->cfg.stride * p->tsize;
    for (size_t t = 0; t < n; t++) {
        cudaError_t e = cudaMemcpy2DAsync(static_cast<uint8_t *>(dst) + t * dpitch, lpitch,
                                          static_cast<const uint8_t *>(src) + t * spitch, lpitch, p->line_bytes, p->cfg.height, kind, st);
        if (e != cudaSuccess) return e;
    }
    return cudaSuccess;
}

/* the calling thread on the pipe's device for the length of a call (the current device is per thread) */
struct OnDevice {
    int before = -1;
    bool ok;
    explicit OnDevice(int dev) { ok = cudaGetDevice(&before) == cudaSuccess && (before == dev || cudaSetDevice(dev) == cudaSuccess); if (before == dev) before = -1; }
    ~OnDevice() { if (before >= 0) cudaSetDevice(before); }
};

static bool drain(qb3cu_pipe *p)
{
    bool ok = true;
    for (Stage &s : p->stages) {
        if (s.st) ok &= note_cuda(cudaStreamSynchronize(s.st)) == QB3CU_OK;
        s.busy = false;
    }
    return ok;
}

} // namespace qb3

extern "C" {

void *qb3cu_host_alloc(size_t bytes)
{
    void *p = nullptr;
    if (note_cuda(cudaHostAlloc(&p, bytes ? bytes : 1, cudaHostAllocDefault)) != QB3CU_OK) return nullptr;
    return p;
}

void qb3cu_host_free(void *p) { if (p) cudaFreeHost(p); }

qb3cu_pipe *qb3cu_pipe_create(const qb3cu_config *cfg, size_t chunk_tiles, int depth)
{
    const size_t slot = qb3cu_slot_bytes(cfg); /* 0 for a bad geometry */
    if (!slot || depth < 0 || depth > 64) return nullptr;
    const size_t line = (size_t)cfg->width * cfg->bands;
    if (cfg->stride && cfg->stride < line) return nullptr;
    qb3cu_pipe *p = new (std::nothrow) qb3cu_pipe;
    if (!p) return nullptr;
    if (note_cuda(cudaGetDevice(&p->device)) != QB3CU_OK) { delete p; return nullptr; }
    p->cfg = *cfg;
    p->dev_cfg = *cfg;
    p->tsize = PIPE_TYPESIZE[cfg->dtype];
    p->line_bytes = line * p->tsize;
    p->tile_bytes = p->line_bytes * cfg->height;
    p->slot = slot;
    p->strided = cfg->stride && cfg->stride != line;
    p->dev_pitch = p->strided ? (size_t)cfg->stride * cfg->height * p->tsize : p->tile_bytes;
    if (!chunk_tiles) {
        chunk_tiles = (200u << 20) / p->tile_bytes;
        if (chunk_tiles >= 32) chunk_tiles &= ~(size_t)31; /* the scan walks 32 streams per warp */
        if (chunk_tiles < 1) chunk_tiles = 1;
        if (chunk_tiles > 4096) chunk_tiles = 4096;
    }
    p->chunk = chunk_tiles;
    p->depth = depth ? (size_t)(depth < 2 ? 2 : depth) : 6; /* a chunk is collected while the next one uploads: two at least */
    p->stages.resize(p->depth);
    bool ok = true;
    for (Stage &s : p->stages) {
        ok = ok && note_cuda(cudaStreamCreateWithFlags(&s.st, cudaStreamNonBlocking)) == QB3CU_OK
                && note_cuda(cudaEventCreateWithFlags(&s.index_ready, cudaEventDisableTiming)) == QB3CU_OK
                && note_cuda(cudaEventCreateWithFlags(&s.uploaded, cudaEventDisableTiming)) == QB3CU_OK
                && note_cuda(cudaHostAlloc(reinterpret_cast<void **>(&s.h_meta), meta_bytes(p->chunk), cudaHostAllocMapped)) == QB3CU_OK;
    }
    if (!ok) { qb3cu_pipe_destroy(p); return nullptr; }
    return p;
}

void qb3cu_pipe_destroy(qb3cu_pipe *p)
{
    if (!p) return;
    OnDevice here(p->device);
    for (Stage &s : p->stages) {
        if (s.st) { cudaStreamSynchronize(s.st); cudaStreamDestroy(s.st); }
        if (s.index_ready) cudaEventDestroy(s.index_ready);
        if (s.uploaded) cudaEventDestroy(s.uploaded);
        if (s.h_meta) cudaFreeHost(s.h_meta);
        s.pix.release(); s.slots.release(); s.packed.release(); s.meta.release();
    }
    delete p;
}

int qb3cu_pipe_encode(qb3cu_pipe *p, const void *h_src, size_t src_tile_pitch, void *h_packed, size_t packed_capacity,
                      uint64_t *h_offsets, uint64_t *h_sizes, uint64_t *h_total, size_t ntiles)
{
    if (!p || !h_src || !h_packed || !h_offsets || !h_sizes || !h_total) return QB3CU_ERR_PARAM;
    OnDevice here(p->device);
    if (!here.ok) return QB3CU_ERR_CUDA;
    if (src_tile_pitch < (p->strided ? p->dev_pitch - ((size_t)p->cfg.stride * p->tsize - p->line_bytes) : p->tile_bytes))
        return QB3CU_ERR_PARAM;
    *h_total = 0;
    const size_t nchunks = (ntiles + p->chunk - 1) / p->chunk;
    uint64_t base = 0; /* bytes of h_packed used so far */
    int rc = QB3CU_OK;

    /* second half of a chunk: its index is on the host, bring down the bytes it produced */
    auto collect = [&](Stage &s) -> int {
        if (note_cuda(cudaEventSynchronize(s.index_ready)) != QB3CU_OK) return QB3CU_ERR_CUDA;
        const uint64_t *sizes = s.h_meta, *offs = s.h_meta + s.n, total = s.h_meta[2 * s.n];
        if (base + total > packed_capacity) return QB3CU_ERR_PARAM;
        if (note_cuda(cudaMemcpyAsync(static_cast<uint8_t *>(h_packed) + base, s.packed.p, total, cudaMemcpyDeviceToHost, s.st)) != QB3CU_OK)
            return QB3CU_ERR_CUDA;
        for (size_t k = 0; k < s.n; k++) {
            h_sizes[s.first + k] = sizes[k];
            h_offsets[s.first + k] = base + offs[k];
        }
        base += total;
        return QB3CU_OK;
    };

    for (size_t i = 0; i < nchunks && rc == QB3CU_OK; i++) {
        Stage &s = p->stages[i % p->depth];
        if (s.busy && note_cuda(cudaStreamSynchronize(s.st)) != QB3CU_OK) { rc = QB3CU_ERR_CUDA; break; }
        /* Never more than two uploads queued: the copy engine serves its queue in order, and a pipe that decodes at
           the same time has small uploads that head a long chain (parse, rebuild, pixels down) and must not wait
           behind a whole batch of pixels. Stages may still be many: they wait for their download, not their upload. */
        if (i >= 2 && note_cuda(cudaEventSynchronize(p->stages[(i - 2) % p->depth].uploaded)) != QB3CU_OK) { rc = QB3CU_ERR_CUDA; break; }
        s.first = i * p->chunk;
        s.n = ntiles - s.first < p->chunk ? ntiles - s.first : p->chunk;
        s.busy = true;
        if (!s.pix.ensure(s.n * p->dev_pitch) || !s.slots.ensure(s.n * p->slot) || !s.packed.ensure(s.n * p->slot)
            || !s.meta.ensure(meta_bytes(p->chunk))) { rc = QB3CU_ERR_CUDA; break; }
        /* The index (sizes, offsets, total) is written by the kernels straight into the stage's page locked host
           block, which the device addresses directly: a copy, however small, would queue in the download engine
           behind whatever another pipe has waiting there, and the next upload cannot be issued before it is read. */
        uint64_t *d_sizes = s.h_meta, *d_offs = d_sizes + s.n, *d_total = d_sizes + 2 * s.n;
        if (note_cuda(copy_tiles(p, s.pix.p, p->dev_pitch, static_cast<const uint8_t *>(h_src) + s.first * src_tile_pitch,
                                 src_tile_pitch, s.n, cudaMemcpyHostToDevice, s.st)) != QB3CU_OK
            || note_cuda(cudaEventRecord(s.uploaded, s.st)) != QB3CU_OK) { rc = QB3CU_ERR_CUDA; break; }
        rc = qb3cu_encode_batch(&p->dev_cfg, s.pix.p, p->dev_pitch, s.slots.p, p->slot, d_sizes, nullptr, nullptr, s.n, s.st);
        if (rc == QB3CU_OK) rc = qb3cu_pack_streams(s.slots.p, p->slot, d_sizes, s.packed.p, d_offs, d_total, s.n, s.st);
        if (rc != QB3CU_OK) break;
        if (note_cuda(cudaEventRecord(s.index_ready, s.st)) != QB3CU_OK) { rc = QB3CU_ERR_CUDA; break; }
        /* the chunk before this one: by now its upload is behind it and this chunk's is queued */
        if (i > 0) rc = collect(p->stages[(i - 1) % p->depth]);
    }
    if (rc == QB3CU_OK && nchunks) rc = collect(p->stages[(nchunks - 1) % p->depth]);
    if (!drain(p) && rc == QB3CU_OK) rc = QB3CU_ERR_CUDA;
    if (rc == QB3CU_OK) *h_total = base;
    return rc;
}

int qb3cu_pipe_decode(qb3cu_pipe *p, const void *h_streams, const uint64_t *h_offsets, const uint64_t *h_lens, void *h_dst,
                      size_t dst_tile_pitch, uint32_t *h_status, int ref_compat, size_t ntiles)
{
    if (!p || !h_streams || !h_offsets || !h_lens || !h_dst || !h_status) return QB3CU_ERR_PARAM;
    OnDevice here(p->device);
    if (!here.ok) return QB3CU_ERR_CUDA;
    if (dst_tile_pitch < (p->strided ? p->dev_pitch - ((size_t)p->cfg.stride * p->tsize - p->line_bytes) : p->tile_bytes))
        return QB3CU_ERR_PARAM;
    const size_t nchunks = (ntiles + p->chunk - 1) / p->chunk;
    int rc = QB3CU_OK;

    auto retire = [&](Stage &s) -> int { /* the chunk that used this stage is complete: hand its status words out */
        if (!s.busy) return QB3CU_OK;
        if (note_cuda(cudaStreamSynchronize(s.st)) != QB3CU_OK)
            return QB3CU_ERR_CUDA;
        memcpy(h_status + s.first, status_of(s.h_meta, s.n), s.n * 4);
        s.busy = false;
        return QB3CU_OK;
    };

    for (size_t i = 0; i < nchunks && rc == QB3CU_OK; i++) {
        Stage &s = p->stages[i % p->depth];
        if ((rc = retire(s)) != QB3CU_OK) break;
        s.first = i * p->chunk;
        s.n = ntiles - s.first < p->chunk ? ntiles - s.first : p->chunk;
        /* the bytes this chunk's streams span in host memory, from a 16 byte boundary so that alignment carries over */
        uint64_t lo = ~0ull, hi = 0;
        for (size_t k = 0; k < s.n; k++) {
            const uint64_t o = h_offsets[s.first + k], e = o + h_lens[s.first + k];
            if (e < o) rc = QB3CU_ERR_PARAM;
            if (o < lo) lo = o;
            if (e > hi) hi = e;
        }
        if (rc != QB3CU_OK) break;
        lo &= ~15ull;
        if (hi < lo) hi = lo;
        uint64_t *lens = s.h_meta, *offs = s.h_meta + s.n;
        for (size_t k = 0; k < s.n; k++) {
            lens[k] = h_lens[s.first + k];
            offs[k] = h_offsets[s.first + k] - lo;
        }
        s.busy = true;
        if (!s.pix.ensure(s.n * p->dev_pitch) || !s.packed.ensure(hi - lo + 64, true) || !s.meta.ensure(meta_bytes(p->chunk))) {
            rc = QB3CU_ERR_CUDA;
            break;
        }
        uint64_t *d_lens = reinterpret_cast<uint64_t *>(s.meta.p), *d_offs = d_lens + s.n;
        uint32_t *d_status = status_of(d_lens, s.n);
        if (note_cuda(cudaMemcpyAsync(s.meta.p, s.h_meta, 2 * s.n * 8, cudaMemcpyHostToDevice, s.st)) != QB3CU_OK
            || (hi > lo && note_cuda(cudaMemcpyAsync(s.packed.p, static_cast<const uint8_t *>(h_streams) + lo, hi - lo,
                                                     cudaMemcpyHostToDevice, s.st)) != QB3CU_OK)) { rc = QB3CU_ERR_CUDA; break; }
        rc = decode_batch_shared(&p->dev_cfg, s.packed.p, d_offs, d_lens, s.pix.p, p->dev_pitch, d_status, ref_compat, s.n, s.st);
        if (rc != QB3CU_OK) break;
        if (note_cuda(copy_tiles(p, static_cast<uint8_t *>(h_dst) + s.first * dst_tile_pitch, dst_tile_pitch, s.pix.p,
                                 p->dev_pitch, s.n, cudaMemcpyDeviceToHost, s.st)) != QB3CU_OK) { rc = QB3CU_ERR_CUDA; break; }
        if (note_cuda(cudaMemcpyAsync(status_of(s.h_meta, s.n), d_status, s.n * 4, cudaMemcpyDeviceToHost, s.st)) != QB3CU_OK) {
            rc = QB3CU_ERR_CUDA;
            break;
        }
    }
    for (Stage &s : p->stages) {
        const int r = retire(s);
        if (rc == QB3CU_OK) rc = r;
    }
    if (rc != QB3CU_OK) drain(p);
    return rc;
}

} /* extern "C" */

/* ------------------------------------------------------------------ several devices from one process */

/*
 * qb3cu_multi_*: a batch in host memory sharded over G devices by one call (contiguous tile ranges, one
 * host thread and one pipe per device, nothing exchanged between devices -- tiles are independent QB3 streams).
 * What a C++ caller (GDAL / MRF, cqb3cu) needs to use all the GPUs of a box without writing the sharding itself.
 */
#include <thread>

struct qb3cu_multi {
    std::vector<qb3cu_pipe *> pipes; /* pipe g lives on device devs[g] */
    std::vector<int> devs;
};

namespace qb3 {
/* tiles [lo, hi) of n for share g of G: contiguous, sizes differ by at most one */
static void share_of(size_t n, size_t g, size_t G, size_t &lo, size_t &hi)
{
    lo = n * g / G;
    hi = n * (g + 1) / G;
}
} // namespace qb3

extern "C" {

qb3cu_multi *qb3cu_multi_create(const qb3cu_config *cfg, const int *devices, int ndevices, size_t chunk_tiles, int depth)
{
    int count = 0;
    if (!cfg || ndevices < 0 || ndevices > 64 || note_cuda(cudaGetDeviceCount(&count)) != QB3CU_OK || count < 1) return nullptr;
    if (ndevices == 0) ndevices = count; /* all of them */
    qb3cu_multi *m = new (std::nothrow) qb3cu_multi;
    if (!m) return nullptr;
    int before = 0;
    cudaGetDevice(&before);
    for (int g = 0; g < ndevices; g++) {
        const int dev = devices ? devices[g] : g;
        qb3cu_pipe *p = nullptr;
        if (dev >= 0 && dev < count && cudaSetDevice(dev) == cudaSuccess) p = qb3cu_pipe_create(cfg, chunk_tiles, depth);
        if (!p) {
            cudaSetDevice(before);
            qb3cu_multi_destroy(m);
            return nullptr;
        }
        m->pipes.push_back(p);
        m->devs.push_back(dev);
    }
    cudaSetDevice(before);
    return m;
}

void qb3cu_multi_destroy(qb3cu_multi *m)
{
    if (!m) return;
    for (qb3cu_pipe *p : m->pipes) qb3cu_pipe_destroy(p);
    delete m;
}

int qb3cu_multi_devices(const qb3cu_multi *m) { return m ? (int)m->pipes.size() : 0; }

int qb3cu_multi_encode(qb3cu_multi *m, const void *h_src, size_t src_tile_pitch, void *h_packed, size_t packed_capacity,
                       uint64_t *h_offsets, uint64_t *h_sizes, uint64_t *h_total, size_t ntiles)
{
    if (!m || m->pipes.empty() || !h_src || !h_packed || !h_offsets || !h_sizes || !h_total) return QB3CU_ERR_PARAM;
    const size_t G = m->pipes.size();
    /* every device packs its tile range into its own share of h_packed, shares starting at 16 byte multiples */
    const size_t share = (packed_capacity / G) & ~(size_t)15;
    std::vector<int> rc(G, QB3CU_OK);
    std::vector<uint64_t> used(G, 0);
    std::vector<std::thread> th;
    for (size_t g = 0; g < G; g++)
        th.emplace_back([&, g] {
            size_t lo, hi;
            share_of(ntiles, g, G, lo, hi);
            if (hi == lo) return;
            rc[g] = qb3cu_pipe_encode(m->pipes[g], static_cast<const uint8_t *>(h_src) + lo * src_tile_pitch, src_tile_pitch,
                                      static_cast<uint8_t *>(h_packed) + g * share, share, h_offsets + lo, h_sizes + lo,
                                      &used[g], hi - lo);
            for (size_t t = lo; t < hi && rc[g] == QB3CU_OK; t++) h_offsets[t] += g * share;
        });
    for (std::thread &t : th) t.join();
    uint64_t total = 0;
    for (size_t g = 0; g < G; g++) {
        if (rc[g] != QB3CU_OK) return rc[g];
        total += used[g];
    }
    *h_total = total;
    return QB3CU_OK;
}

int qb3cu_multi_decode(qb3cu_multi *m, const void *h_streams, const uint64_t *h_offsets, const uint64_t *h_lens, void *h_dst,
                       size_t dst_tile_pitch, uint32_t *h_status, int ref_compat, size_t ntiles)
{
    if (!m || m->pipes.empty() || !h_streams || !h_offsets || !h_lens || !h_dst || !h_status) return QB3CU_ERR_PARAM;
    const size_t G = m->pipes.size();
    std::vector<int> rc(G, QB3CU_OK);
    std::vector<std::thread> th;
    for (size_t g = 0; g < G; g++)
        th.emplace_back([&, g] {
            size_t lo, hi;
            share_of(ntiles, g, G, lo, hi);
            if (hi == lo) return;
            rc[g] = qb3cu_pipe_decode(m->pipes[g], h_streams, h_offsets + lo, h_lens + lo,
                                      static_cast<uint8_t *>(h_dst) + lo * dst_tile_pitch, dst_tile_pitch, h_status + lo,
                                      ref_compat, hi - lo);
        });
    for (std::thread &t : th) t.join();
    for (size_t g = 0; g < G; g++)
        if (rc[g] != QB3CU_OK) return rc[g];
    return QB3CU_OK;
}

} /* extern "C" */
