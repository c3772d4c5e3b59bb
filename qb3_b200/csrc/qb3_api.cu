/*
 * qb3_api.cu -- the QB3.h C API (reference: QB3lib/QB3.h:85-162) as batch-of-one wrappers over the
 * device entry points of qb3cu.h. Handle life cycle, option validation and header parsing are host code
 * and keep the reference's observable behaviour (QB3encode.cpp:26-134, QB3decode.cpp:36-264); pixels and
 * bits are only ever touched by the CUDA kernels. Without a usable CUDA device qb3_encode and
 * qb3_read_data fail (return 0, encoder state QB3E_LIBERR): there is no CPU codec here.
 */
#include <atomic>
#include <cstdlib>
#include <cstring>
#include <limits>

#include "../../include/QB3.h"
#include "qb3_device.cuh"

using namespace qb3;

namespace qb3 {
int note_cuda(cudaError_t e);
static const size_t TSIZE[8] = {1, 1, 2, 2, 4, 4, 8, 8};

/* device scratch owned by a handle, grown on demand */
struct DevBuf {
    void *p = nullptr;
    size_t cap = 0;
    bool reserve(size_t n)
    {
        if (n <= cap) return true;
        if (p) cudaFree(p);
        p = nullptr; cap = 0;
        if (note_cuda(cudaMalloc(&p, n)) != QB3CU_OK) return false;
        cap = n;
        return true;
    }
    void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
};
} // namespace qb3

/* Encoder handle: the settings of the reference's struct encs (QB3common.h:68-88) plus device scratch */
struct encs {
    qb3cu_config cfg;      /* cfg.order doubles as the reference's sticky 'order' field */
    uint64_t state[3 * QB3CU_MAXBANDS]; /* per band prev, runbits, cf; survives qb3_encode (QB3encode.h:446-449) */
    int error;
    cudaStream_t stream;
    DevBuf src, dst, aux;
};

/* Decoder handle: the reference's struct decs (QB3common.h:91-111) */
struct decs {
    size_t xsize, ysize, nbands, stride;
    uint64_t order, quanta;
    int error, stage;
    uint8_t cband[QB3CU_MAXBANDS];
    int mode, type;
    uint8_t *s_in;          /* caller's buffer, from the first chunk on (after qb3_read_info: the payload) */
    size_t s_size;
    uint8_t *s_start;       /* caller's buffer, whole stream */
    size_t s_total;
    cudaStream_t stream;
    DevBuf src, dst, aux;
};

/* QB3_REF_COMPAT=1 in the environment: a stream without a CB chunk decodes with the reference's all-zero band map
   (SURVEY 4.3 D1). A user facing switch, looked up when a decoder handle is made so that a program can change it
   between images; the only environment variable this library reads. */
static int ref_compat_env()
{
    const char *e = getenv("QB3_REF_COMPAT");
    return e && e[0] == '1' ? 1 : 0;
}

static bool ensure_stream(cudaStream_t &s)
{
    if (s) return true;
    return note_cuda(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking)) == QB3CU_OK;
}

/* what QB3_MAXBANDS is for the callers of the QB3.h functions: the reference's 16 unless a caller says otherwise */
static std::atomic<uint32_t> g_api_max_bands(16);

extern "C" {

uint32_t qb3cu_api_max_bands(uint32_t bands)
{
    if (bands >= 1 && bands <= QB3CU_MAXBANDS) g_api_max_bands = bands;
    return g_api_max_bands;
}

/* ------------------------------------------------------------------ encoder */

encsp qb3_create_encoder(size_t w, size_t h, size_t b, qb3_dtype dt)
{
    if (w == 0 || w > 0x10000 || h == 0 || h > 0x10000 || b == 0 || b > g_api_max_bands || (unsigned)dt > QB3_I64)
        return nullptr;
    encsp p = new encs();
    if (qb3cu_config_init(&p->cfg, (uint32_t)w, (uint32_t)h, (uint32_t)b, (uint32_t)dt) != QB3CU_OK) {
        delete p;
        return nullptr;
    }
    p->stream = nullptr;
    qb3_reset_encoder(p);
    return p;
}

void qb3_reset_encoder(encsp p)
{
    memset(p->state, 0, sizeof(p->state));
    p->error = 0;
}

void qb3_destroy_encoder(encsp p)
{
    if (!p) return;
    p->src.release(); p->dst.release(); p->aux.release();
    if (p->stream) cudaStreamDestroy(p->stream);
    delete p;
}

/* reference: QB3encode.cpp:63-77 */
bool qb3_set_encoder_coreband(encsp p, size_t bands, size_t *cband)
{
    if (bands != p->cfg.bands) return false;
    uint8_t *m = p->cfg.cband;
    for (size_t i = 0; i < bands; i++) m[i] = (uint8_t)(cband[i] < bands ? cband[i] : i);
    for (size_t i = 0; i < bands; i++) if (m[i] != i) m[m[i]] = m[i]; /* a core band is never itself derived */
    for (size_t i = 0; i < bands; i++) cband[i] = m[i];
    return true;
}

void qb3_set_encoder_stride(encsp p, size_t stride) { p->cfg.stride = stride; }

/* reference: QB3encode.cpp:87-110; the value is kept even when it is reported as too large */
bool qb3_set_encoder_quanta(encsp p, uint64_t q, bool away)
{
    if (q < 1) return false;
    p->cfg.quanta = q;
    p->cfg.away = away;
    if (q == 1) return true;
    static const uint64_t tmax[8] = {0xff, 0x7f, 0xffff, 0x7fff, 0xffffffffull, 0x7fffffffull,
                                     ~0ull, 0x7fffffffffffffffull};
    /* the reference's switch falls through from the type's case to the end, so the bound that applies is the
       smallest among the listed types from this one on (I8,U8,I16,U16,I32,U32,I64) */
    static const int order[7] = {QB3_I8, QB3_U8, QB3_I16, QB3_U16, QB3_I32, QB3_U32, QB3_I64};
    bool error = false, on = false;
    for (int i = 0; i < 7; i++) {
        on |= order[i] == (int)p->cfg.dtype;
        if (on) error |= q > tmax[order[i]];
    }
    return !error;
}

size_t qb3_max_encoded_size(const encsp p) { return qb3cu_max_encoded_size(&p->cfg); }

/* reference: QB3encode.cpp:120-134; a legacy mode switches the curve to Z for good */
qb3_mode qb3_set_encoder_mode(encsp p, qb3_mode mode)
{
    if (mode >= 0 && mode < QB3M_END) p->cfg.mode = (uint32_t)mode;
    if (p->cfg.mode <= QB3M_CF_RLE) p->cfg.order = ZCURVE;
    return (qb3_mode)p->cfg.mode;
}

int qb3_get_encoder_state(encsp p) { return p->error; }

size_t qb3_encode(encsp p, void *source, void *destination)
{
    const qb3cu_config &c = p->cfg;
    const size_t ts = TSIZE[c.dtype], line = (size_t)c.width * c.bands * ts;
    const size_t pitch = (c.stride ? c.stride : (size_t)c.width * c.bands) * ts;
    for (uint32_t b = 0; b < c.bands; b++)
        if (c.cband[b] >= c.bands) { p->error = 2; return 0; } /* reference: check_info, QB3encode.h:364-373 */
    if (pitch < line) { p->error = QB3E_EINV; return 0; }
    const size_t slot = qb3cu_slot_bytes(&c), nstate = 3 * (size_t)c.bands;
    if (!ensure_stream(p->stream) || !p->src.reserve(line * c.height) || !p->dst.reserve(slot)
        || !p->aux.reserve(16 + nstate * 8)) {
        p->error = QB3E_LIBERR;
        return 0;
    }
    /* device copy is compact; the caller's stride is honoured by the 2D copy */
    qb3cu_config dc = c;
    dc.stride = 0;
    uint64_t *d_size = static_cast<uint64_t *>(p->aux.p);
    uint32_t *d_status = reinterpret_cast<uint32_t *>(d_size + 1);
    uint64_t *d_state = d_size + 2;
    cudaStream_t st = p->stream;
    uint64_t size = 0;
    const bool keep_state = c.quanta > 1 || c.width < 4 || c.height < 4;
    bool ok = note_cuda(cudaMemcpy2DAsync(p->src.p, line, source, pitch, line, c.height, cudaMemcpyHostToDevice, st)) == QB3CU_OK
        && note_cuda(cudaMemcpyAsync(d_state, p->state, nstate * 8, cudaMemcpyHostToDevice, st)) == QB3CU_OK
        && qb3cu_encode_batch(&dc, p->src.p, line * c.height, p->dst.p, slot, d_size, d_status, d_state, 1, st) == QB3CU_OK
        && note_cuda(cudaMemcpyAsync(&size, d_size, 8, cudaMemcpyDeviceToHost, st)) == QB3CU_OK
        /* The running state comes back into the handle for plain images only. Quantised images and images with a side
           under 4 are coded through a copy of the handle in the reference (encs subimg(*p), QB3encode.cpp:405, and
           smallimg, :352), so there the handle's own state stays as it was and a second qb3_encode gives the same
           stream again. */
        && (keep_state || note_cuda(cudaMemcpyAsync(p->state, d_state, nstate * 8, cudaMemcpyDeviceToHost, st)) == QB3CU_OK)
        && note_cuda(cudaStreamSynchronize(st)) == QB3CU_OK;
    if (ok && size > 0 && size <= slot)
        ok = note_cuda(cudaMemcpy(destination, p->dst.p, size, cudaMemcpyDeviceToHost)) == QB3CU_OK;
    else ok = false;
    if (!ok) {
        p->error = QB3E_LIBERR;
        return 0;
    }
    return (size_t)size;
}

/* ------------------------------------------------------------------ decoder */

namespace {
/* little endian field of up to 8 bytes at a bit-stream position, zero beyond the end like iBits::peek (bitstream.h:39-50) */
uint64_t field(const uint8_t *p, size_t size, size_t at, size_t bytes)
{
    uint64_t v = 0;
    for (size_t i = 0; i < bytes && i < 8; i++)
        if (at + i < size) v |= (uint64_t)p[at + i] << (8 * i);
    return v;
}
} // namespace

/* reference: QB3decode.cpp:130-172 */
decsp qb3_read_start(void *source, size_t source_size, size_t *image_size)
{
    const size_t HDRSZ = 11;
    if (source_size < HDRSZ + 4 || !image_size || !source) return nullptr;
    const uint8_t *s = static_cast<const uint8_t *>(source);
    if (s[0] != 'Q' || s[1] != 'B' || s[2] != '3' || s[3] != 0x80) return nullptr;
    decsp p = new decs();
    p->xsize = 1 + field(s, source_size, 4, 2);
    p->ysize = 1 + field(s, source_size, 6, 2);
    p->nbands = 1 + s[8];
    p->type = s[9];
    p->mode = s[10];
    if (p->nbands > g_api_max_bands || (p->mode >= QB3M_END && p->mode != QB3M_STORED)
        || ((s[11] | s[12]) & 0x80) || p->type > QB3_I64) {
        delete p;
        return nullptr;
    }
    p->s_start = static_cast<uint8_t *>(source);
    p->s_total = source_size;
    p->s_in = p->s_start + HDRSZ;
    p->s_size = source_size - HDRSZ;
    image_size[0] = p->xsize;
    image_size[1] = p->ysize;
    image_size[2] = p->nbands;
    if (p->mode <= QB3M_CF_RLE) p->order = ZCURVE;
    /* band map when no CB chunk follows: identity, which is what the format says (doc/QB3.md:255). The reference
       leaves it zeroed; QB3_REF_COMPAT=1 in the environment selects that behaviour. */
    if (!ref_compat_env())
        for (size_t c = 0; c < p->nbands; c++) p->cband[c] = (uint8_t)c;
    p->error = QB3E_OK;
    p->stage = 1;
    return p;
}

/* reference: QB3decode.cpp:176-264 */
bool qb3_read_info(decsp p)
{
    if (p->stage != 1 || p->error || !p->s_in || p->s_size < 4) {
        if (p->error == QB3E_OK) p->error = QB3E_EINV;
        return false;
    }
    const uint8_t *s = p->s_in;
    const size_t n = p->s_size;
    size_t at = 0;
    do {
        const uint32_t chunk = (uint32_t)field(s, n, at, 2), len = (uint32_t)field(s, n, at + 2, 2);
        if (chunk == ('Q' | ('V' << 8))) {
            if (len > 4 || len < 1) { p->error = QB3E_EINV; break; }
            p->quanta = field(s, n, at + 4, len);
            at += 4 + len;
            if (p->quanta < 2) p->error = QB3E_EINV;
        }
        else if (chunk == ('C' | ('B' << 8))) {
            if (len != p->nbands) { p->error = QB3E_EINV; break; }
            for (size_t i = 0; i < p->nbands; i++) {
                p->cband[i] = (uint8_t)field(s, n, at + 4 + i, 1);
                if (p->cband[i] >= p->nbands) p->error = QB3E_EINV;
            }
            at += 4 + len;
        }
        else if (chunk == ('D' | ('T' << 8))) {
            at += 2;
            if (n <= at) { p->error = QB3E_EINV; break; }
            p->s_in += at;
            p->s_size -= at;
            p->stage = 2;
        }
        else if (chunk == ('S' | ('C' << 8))) {
            if (len != 8) { p->error = QB3E_EINV; break; }
            if (p->mode < QB3M_BASE_H || p->mode == QB3M_STORED) { p->error = QB3E_EINV; break; }
            p->order = field(s, n, at + 4, 8);
            at += 12;
            uint32_t seen = 0;
            for (int i = 0; i < 16; i++) seen |= 1u << ((p->order >> (4 * i)) & 15);
            if (seen != 0xffff) { p->error = QB3E_EINV; break; }
        }
        else {
            /* The reference steps over a lower case chunk by its payload length only (QB3decode.cpp:254-255), which
               re-reads the same chunk; no stream with such a chunk decodes there, so report it as unknown. */
            p->error = QB3E_UNKN;
        }
    } while (p->stage != 2 && p->error == QB3E_OK && at < n);
    if (p->error == QB3E_OK && p->stage != 2) p->error = QB3E_EINV;
    return p->error == QB3E_OK;
}

void qb3_destroy_decoder(decsp p)
{
    if (!p) return;
    p->src.release(); p->dst.release(); p->aux.release();
    if (p->stream) cudaStreamDestroy(p->stream);
    delete p;
}

size_t qb3_decoded_size(const decsp p) { return p->xsize * p->ysize * p->nbands * (p->type > QB3_I64 ? 0 : TSIZE[p->type]); }
qb3_dtype qb3_get_type(const decsp p) { return (qb3_dtype)p->type; }
qb3_mode qb3_get_mode(const decsp p) { return p->stage == 2 ? (qb3_mode)p->mode : QB3M_INVALID; }
uint64_t qb3_get_quanta(const decsp p) { return p->stage == 2 ? p->quanta : 0; }
/* reports the Z curve when no curve was read, as the reference does (QB3decode.cpp:56-60) */
uint64_t qb3_get_order(const decsp p) { return p->stage != 2 ? 0 : (p->order ? p->order : ZCURVE); }
void qb3_set_decoder_stride(decsp p, size_t stride) { p->stride = stride; }

bool qb3_get_coreband(const decsp p, size_t *coreband)
{
    if (p->stage != 2) return false;
    for (size_t c = 0; c < p->nbands; c++) coreband[c] = p->cband[c];
    return true;
}

/* reference: QB3decode.cpp:455-464, 380-452 */
size_t qb3_read_data(decsp p, void *destination)
{
    if (p->stage != 2 || p->error != QB3E_OK || !p->s_in || p->s_size == 0) {
        if (p->error == QB3E_OK) p->error = QB3E_EINV;
        return 0;
    }
    qb3cu_config c;
    if (qb3cu_config_init(&c, (uint32_t)p->xsize, (uint32_t)p->ysize, (uint32_t)p->nbands, (uint32_t)p->type) != QB3CU_OK) {
        p->error = QB3E_EINV;
        return 0;
    }
    if (p->mode >= 0 && p->mode < QB3M_END) c.mode = (uint32_t)p->mode; /* an RLE stream: expanded ahead of the parallel decode */
    const size_t ts = TSIZE[p->type], line = p->xsize * p->nbands * ts, out = line * p->ysize;
    const size_t pitch = (p->stride ? p->stride : p->xsize * p->nbands) * ts;
    if (pitch < line) { p->error = QB3E_EINV; return 0; }
    const int ref_compat = ref_compat_env();
    if (!ensure_stream(p->stream) || !p->src.reserve(p->s_total + 16) || !p->dst.reserve(out) || !p->aux.reserve(32)) {
        p->error = QB3E_LIBERR;
        return 0;
    }
    uint64_t meta[2] = {0, p->s_total};
    uint64_t *d_meta = static_cast<uint64_t *>(p->aux.p);
    uint32_t *d_status = reinterpret_cast<uint32_t *>(d_meta + 2);
    uint32_t status = 0xffffffffu;
    cudaStream_t st = p->stream;
    bool ok = note_cuda(cudaMemcpyAsync(p->src.p, p->s_start, p->s_total, cudaMemcpyHostToDevice, st)) == QB3CU_OK
        && note_cuda(cudaMemcpyAsync(d_meta, meta, 16, cudaMemcpyHostToDevice, st)) == QB3CU_OK
        && qb3cu_decode_batch(&c, p->src.p, d_meta, d_meta + 1, p->dst.p, out, d_status, ref_compat, 1, st) == QB3CU_OK
        && note_cuda(cudaMemcpyAsync(&status, d_status, 4, cudaMemcpyDeviceToHost, st)) == QB3CU_OK
        && note_cuda(cudaStreamSynchronize(st)) == QB3CU_OK;
    if (!ok) { p->error = QB3E_LIBERR; return 0; }
    if (status != QB3CU_TILE_OK) {
        p->error = status == QB3CU_TILE_RLE_TOO_BIG ? QB3E_ERR : QB3E_EINV;
        return 0;
    }
    if (note_cuda(cudaMemcpy2D(destination, pitch, p->dst.p, line, line, p->ysize, cudaMemcpyDeviceToHost)) != QB3CU_OK) {
        p->error = QB3E_LIBERR;
        return 0;
    }
    return out;
}

} /* extern "C" */
