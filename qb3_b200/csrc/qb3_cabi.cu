/*
 * qb3_cabi.cu -- the batched C ABI declared in include/qb3cu.h: argument checking, header bytes,
 * launch geometry. No pixel or bit work happens on the host.
 */
#include <atomic>
#include <cstdlib>
#include <cstring>

#include "qb3_device.cuh"
#include "qb3_host.h"

namespace qb3 {
cudaError_t launch_encode(const EncArgs &a, uint32_t tsize, size_t ntiles, uint32_t threads, size_t smem, cudaStream_t st);
cudaError_t launch_decode(const DecArgs &a, uint32_t tsize, cudaStream_t st, uint32_t &launches);
cudaMemPool_t scratch_pool(); /* qb3_decode.cu: the library's own stream ordered pool, which keeps its memory */
cudaError_t launch_pack(const uint8_t *slots, uint64_t slot, const unsigned long long *sizes, uint8_t *packed,
                        unsigned long long *offsets, unsigned long long *total, uint32_t ntiles, cudaStream_t st);

int decode_batch_shared(const qb3cu_config *cfg, const void *d_streams, const uint64_t *d_offsets, const uint64_t *d_lens,
                        void *d_dst, size_t dst_tile_pitch, uint32_t *d_status, int ref_compat, size_t ntiles, void *stream);

static thread_local int g_last_cuda_error = 0;
static std::atomic<uint64_t> g_launches(0);

static const uint32_t TYPESIZE[8] = {1, 1, 2, 2, 4, 4, 8, 8};

int note_cuda(cudaError_t e)
{
    if (e == cudaSuccess) return QB3CU_OK;
    g_last_cuda_error = (int)e;
    return QB3CU_ERR_CUDA;
}
void count_launches(uint64_t n) { g_launches += n; }

static bool rle_requested(uint32_t mode) { return mode == 2 || mode == 3 || mode == 6 || mode == 7; }

static bool geometry_ok(const qb3cu_config *c)
{
    return c && c->width >= 1 && c->width <= 0x10000 && c->height >= 1 && c->height <= 0x10000
        && c->bands >= 1 && c->bands <= QB3CU_MAXBANDS && c->dtype <= 7;
}

} // namespace qb3

using namespace qb3;

extern "C" {

int qb3cu_config_init(qb3cu_config *cfg, uint32_t width, uint32_t height, uint32_t bands, uint32_t dtype)
{
    if (!cfg) return QB3CU_ERR_PARAM;
    memset(cfg, 0, sizeof(*cfg));
    cfg->width = width; cfg->height = height; cfg->bands = bands; cfg->dtype = dtype;
    if (!geometry_ok(cfg)) return QB3CU_ERR_PARAM;
    cfg->mode = M_FTL;
    cfg->quanta = 1;
    for (uint32_t c = 0; c < bands; c++) cfg->cband[c] = (uint8_t)c;
    if (bands == 3 || bands == 4) cfg->cband[0] = cfg->cband[2] = 1;
    return QB3CU_OK;
}

size_t qb3cu_max_encoded_size(const qb3cu_config *cfg)
{
    if (!geometry_ok(cfg)) return 0;
    /* same expression, in double, as the reference (QB3encode.cpp:112-118): it also gates the RLE pass */
    const size_t n = (size_t)16 * ((cfg->width + 3) / 4) * ((cfg->height + 3) / 4) * cfg->bands;
    const double bits_per_value = 17.0 / 16.0 + 8 * TYPESIZE[cfg->dtype];
    return 1024 + static_cast<size_t>(bits_per_value * n / 8);
}

size_t qb3cu_slot_bytes(const qb3cu_config *cfg)
{
    const size_t m = qb3cu_max_encoded_size(cfg);
    return m ? ((m + 15) & ~(size_t)15) + 16 : 0;
}

int qb3cu_last_cuda_error(void) { return g_last_cuda_error; }
uint64_t qb3cu_kernel_launches(void) { return g_launches.load(); }

} /* extern "C" */

/* qb3cu_encode_batch, and with size_only qb3cu_encoded_size_batch: the same launch with the packing left out */
static int encode_impl(const qb3cu_config *cfg, const void *d_src, size_t src_tile_pitch, void *d_dst,
                       size_t dst_slot_bytes, uint64_t *d_sizes, uint32_t *d_status, uint64_t *d_state,
                       size_t ntiles, void *stream, bool size_only)
{
    if (!geometry_ok(cfg) || cfg->mode > M_FTL || cfg->quanta < 1 || !d_src || (!d_dst && !size_only) || !d_sizes) return QB3CU_ERR_PARAM;
    if (ntiles == 0) return QB3CU_OK;
    if (ntiles > 0x7fffffffull) return QB3CU_ERR_PARAM;
    const uint32_t tsize = TYPESIZE[cfg->dtype], bits = 8 * tsize;
    for (uint32_t c = 0; c < cfg->bands; c++) if (cfg->cband[c] >= cfg->bands) return QB3CU_ERR_PARAM;
    if (((uintptr_t)d_src | src_tile_pitch) % tsize) return QB3CU_ERR_PARAM;
    if (size_only) dst_slot_bytes = qb3cu_slot_bytes(cfg);
    if (((uintptr_t)d_dst | dst_slot_bytes) % 16 || dst_slot_bytes < qb3cu_slot_bytes(cfg)) return QB3CU_ERR_PARAM;
    const uint64_t line = (uint64_t)cfg->width * cfg->bands;
    if (cfg->stride && cfg->stride < line) return QB3CU_ERR_PARAM;

    EncArgs a;
    memset(&a, 0, sizeof(a));
    a.src = static_cast<const uint8_t *>(d_src);
    a.src_pitch = src_tile_pitch;
    a.dst = static_cast<uint8_t *>(d_dst);
    a.slot = dst_slot_bytes;
    a.sizes = reinterpret_cast<unsigned long long *>(d_sizes);
    a.status = d_status;
    a.state = reinterpret_cast<unsigned long long *>(d_state);
    a.stride = cfg->stride ? cfg->stride : line;
    a.quanta = cfg->quanta;
    a.w = cfg->width; a.h = cfg->height; a.bands = cfg->bands;
    a.raw_size = line * cfg->height * tsize;
    a.max_size = qb3cu_max_encoded_size(cfg);
    a.is_signed = cfg->dtype & 1;
    a.away = cfg->away != 0;
    a.size_only = size_only;
    memcpy(a.cband, cfg->cband, sizeof(a.cband));

    /* mode: the RLE variants code as their base mode, RLE is a byte pass afterwards (reference: QB3encode.cpp:494-506) */
    a.rle_mode = rle_requested(cfg->mode) ? cfg->mode : 0;
    a.mode = a.rle_mode ? cfg->mode - 2 : cfg->mode;
    a.order = cfg->order ? cfg->order : (cfg->mode <= 3 ? ZCURVE : HILBERT);
    a.hdr_len = build_headers(cfg, a.mode, a.order, a.hdr);
    a.hdr_stored_len = build_headers(cfg, M_STORED, a.order, a.hdr_stored);

    /* coded geometry; images with a side under 4 are reordered into 4 wide / 4 high strips (reference: QB3encode.cpp:351-389) */
    a.vw = a.w; a.vh = a.h;
    if ((uint64_t)a.w * a.h <= 16) a.small = 3; /* stored outright (reference: QB3encode.cpp:490-491) */
    else if (a.w < 4) { a.small = 1; a.vw = 4; a.vh = 4 * ((a.w * a.h + 15) / 16); }
    else if (a.h < 4) { a.small = 2; a.vh = 4; a.vw = 4 * ((a.w * a.h + 15) / 16); }
    a.nbx = (a.vw + 3) / 4;
    a.nby = (a.vh + 3) / 4;
    a.vec_stage = a.quanta == 1 && a.small == 0;

    /* one thread per group: as many whole blocks per iteration as fit the CTA, block rows split evenly. CTAs of up to
       384 threads have builds that fit more of them on an SM (encode_kernel's DENSE: three instead of two for 8 bit
       FTL / BASE, two instead of one for 8 and 16 bit BEST, Hilbert curve): the cap that puts more threads on an SM wins */
    const bool best_mode = a.mode == M_CF_Z || a.mode == M_CF_H;
    auto segments = [&](uint32_t max_threads) -> uint32_t {
        uint32_t seg_blocks = max_threads / a.bands;
        if (seg_blocks < 1) seg_blocks = 1;
        if (seg_blocks > a.nbx) seg_blocks = a.nbx;
        a.segs = (a.nbx + seg_blocks - 1) / seg_blocks;
        a.seg_blocks = (a.nbx + a.segs - 1) / a.segs;
        /* segments that start on 16 byte boundaries of the row can be staged by bulk copies */
        uint32_t unit = 16, bb = 4 * a.bands * tsize;
        while (bb % unit) unit >>= 1;
        const uint32_t al = 16 / unit, up = (a.seg_blocks + al - 1) / al * al;
        if (up <= seg_blocks) a.seg_blocks = up;
        else if (a.seg_blocks >= al) a.seg_blocks = a.seg_blocks / al * al;
        a.segs = (a.nbx + a.seg_blocks - 1) / a.seg_blocks;
        return (a.seg_blocks * a.bands + 31) & ~31u;
    };
    uint32_t threads = segments(tsize <= 2 ? 512 : 256);
    if (a.order == HILBERT && a.bands <= 384 && threads > 384 && (best_mode ? tsize <= 2 : tsize == 1)) {
        const uint32_t wide = threads * (best_mode ? 1 : 2), dense = segments(384) * (best_mode ? 2 : 3);
        if (dense < wide) segments(512);
        threads = (a.seg_blocks * a.bands + 31) & ~31u;
    }
    if (threads > 512) return QB3CU_ERR_PARAM;
    /* bulk copies (TMA) for the row staging when every segment's rows are whole 16 byte units at 16 byte addresses */
    a.bulk_stage = a.vec_stage && (((uintptr_t)d_src | src_tile_pitch | (a.stride * tsize)) & 15) == 0;
    for (uint32_t sg = 0; sg < a.segs && a.bulk_stage; sg++) {
        const uint32_t bx0 = sg * a.seg_blocks, nblk = a.seg_blocks < a.nbx - bx0 ? a.seg_blocks : a.nbx - bx0;
        const uint32_t xs = 4 * bx0 < a.vw - 4 ? 4 * bx0 : a.vw - 4, xe = 4 * (bx0 + nblk) < a.vw ? 4 * (bx0 + nblk) : a.vw;
        if (((xs * a.bands * tsize) | ((xe - xs) * a.bands * tsize)) & 15) a.bulk_stage = 0;
    }
    a.simd8 = a.bulk_stage && tsize == 1 && a.bands <= 4 && a.vw % 4 == 0 && (a.order == HILBERT || a.order == ZCURVE);
    a.rowpitch = ((a.seg_blocks * 4 * a.bands * tsize + 15) & ~15u) + 16;
    a.win_words = ((a.hdr_len * 8 + 128 + threads * max_group_bits(bits)) / 32 + 16 + 3) & ~3u;
    size_t smem = (size_t)a.win_words * 4 + 8 * (size_t)a.rowpitch + 2 * (size_t)a.bands * 8 + 36 * 4
                + 2 * (size_t)a.bands + threads;
    smem = (smem + 7) & ~(size_t)7;
    a.best_off = (uint32_t)smem;
    if (a.mode == M_CF_Z || a.mode == M_CF_H)
        smem += (size_t)threads * 8 + 2 * (size_t)a.bands * 8 + (size_t)threads * 4 + 8 + (size_t)a.bands * 8 + 2 * (size_t)a.bands;
    smem = (smem + 7) & ~(size_t)7;
    a.lut_off = (uint32_t)smem;
    smem += 508 * 4 + 64 * 2;
    if (smem > 200 * 1024) return QB3CU_ERR_PARAM;

    /* Few large tiles: one CTA per tile would leave most of the GPU idle (a 4096 x 4096 tile alone takes 1.5 ms), so a
       tile is cut into parts of whole block rows, a CTA each, and the parts' bits are joined afterwards. BEST needs a
       second look at the parts that depended on the factor the parts before them left behind (EncArgs::best_pass). */
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    uint8_t *tmp = nullptr;
    const bool best = a.mode == M_CF_Z || a.mode == M_CF_H;
    int dev = 0, nsm = 148;
    if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, dev);
    if (a.small == 0 && a.nby >= 16 && ntiles < (size_t)4 * nsm) {
        uint32_t parts = (uint32_t)(((size_t)6 * nsm + ntiles - 1) / ntiles); /* a few CTAs per SM to balance the load */
        if (parts > a.nby / 8) parts = a.nby / 8; /* eight block rows to a part at least */
        if (parts > 64) parts = 64;
        if (parts >= 3) { /* in two parts the joining pass costs more than the second CTA brings (measured) */
            a.part_rows = (a.nby + parts - 1) / parts;
            a.parts = (a.nby + a.part_rows - 1) / a.part_rows;
            const uint64_t groups = (uint64_t)a.part_rows * a.nbx * a.bands;
            a.tmp_slot = ((a.hdr_len + groups * max_group_bits(bits) / 8 + 64) + 15) & ~15ull;
            const size_t lens_bytes = (ntiles * a.parts * 8 + 15) & ~(size_t)15;
            const size_t pcf_bytes = best ? ntiles * a.parts * a.bands * 32 : 0;
            const size_t redo_bytes = best ? (ntiles * a.parts * 4 + 15) & ~(size_t)15 : 0;
            cudaMemPool_t pool = scratch_pool();
            const size_t tmp_bytes = lens_bytes + pcf_bytes + redo_bytes + (size_only ? 0 : a.tmp_slot * a.parts * ntiles);
            if (note_cuda(pool ? cudaMallocFromPoolAsync(reinterpret_cast<void **>(&tmp), tmp_bytes, pool, st)
                               : cudaMallocAsync(reinterpret_cast<void **>(&tmp), tmp_bytes, st)) != QB3CU_OK)
                return QB3CU_ERR_CUDA;
            a.part_bits = reinterpret_cast<unsigned long long *>(tmp);
            a.tmp = tmp + lens_bytes + pcf_bytes + redo_bytes;
            if (best) {
                a.best_pass = 1;
                a.part_pcf = reinterpret_cast<unsigned long long *>(tmp + lens_bytes);
                a.part_redo = reinterpret_cast<uint32_t *>(tmp + lens_bytes + pcf_bytes);
                if (note_cuda(cudaMemsetAsync(a.part_redo, 0, redo_bytes, st)) != QB3CU_OK) {
                    cudaFreeAsync(tmp, st);
                    return QB3CU_ERR_CUDA;
                }
            }
        }
    }
    cudaError_t err = launch_encode(a, tsize, ntiles, threads, smem, st);
    if (tmp) {
        const cudaError_t ferr = cudaFreeAsync(tmp, st);
        if (err == cudaSuccess) err = ferr;
    }
    if (err == cudaSuccess) count_launches((a.rle_mode ? 2 : 1) + (a.parts > 1 ? 1 : 0) + (a.parts > 1 && best ? 2 : 0));
    return note_cuda(err);
}

extern "C" {

int qb3cu_encode_batch(const qb3cu_config *cfg, const void *d_src, size_t src_tile_pitch, void *d_dst,
                       size_t dst_slot_bytes, uint64_t *d_sizes, uint32_t *d_status, uint64_t *d_state,
                       size_t ntiles, void *stream)
{
    return encode_impl(cfg, d_src, src_tile_pitch, d_dst, dst_slot_bytes, d_sizes, d_status, d_state, ntiles, stream, false);
}

int qb3cu_encoded_size_batch(const qb3cu_config *cfg, const void *d_src, size_t src_tile_pitch, uint64_t *d_sizes,
                             size_t ntiles, void *stream)
{
    if (!cfg || !rle_requested(cfg->mode))
        return encode_impl(cfg, d_src, src_tile_pitch, nullptr, 0, d_sizes, nullptr, nullptr, ntiles, stream, true);
    /* an RLE mode: whether the byte pass pays is only known once the bytes exist, so the streams are made, in scratch */
    if (!geometry_ok(cfg) || ntiles == 0) return ntiles == 0 && geometry_ok(cfg) ? QB3CU_OK : QB3CU_ERR_PARAM;
    const size_t slot = qb3cu_slot_bytes(cfg);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    void *scratch = nullptr;
    cudaMemPool_t pool = scratch_pool();
    if (note_cuda(pool ? cudaMallocFromPoolAsync(&scratch, slot * ntiles, pool, st) : cudaMallocAsync(&scratch, slot * ntiles, st)) != QB3CU_OK)
        return QB3CU_ERR_CUDA;
    const int rc = encode_impl(cfg, d_src, src_tile_pitch, scratch, slot, d_sizes, nullptr, nullptr, ntiles, stream, false);
    const int frc = note_cuda(cudaFreeAsync(scratch, st));
    return rc != QB3CU_OK ? rc : frc;
}

} /* extern "C" */

/* shared_sm: the host pipeline's call -- several batches and an encode are in flight beside this one, so the batch is
   packed onto as few SMs as hold its streams */
static int decode_impl(const qb3cu_config *cfg, const void *d_streams, const uint64_t *d_offsets, const uint64_t *d_lens,
                       void *d_dst, size_t dst_tile_pitch, uint32_t *d_status, int ref_compat, size_t ntiles, void *stream,
                       bool shared_sm)
{
    if (!geometry_ok(cfg) || !d_streams || !d_offsets || !d_lens || !d_dst || !d_status) return QB3CU_ERR_PARAM;
    if (ntiles == 0) return QB3CU_OK;
    if (ntiles > 0x7fffffffull) return QB3CU_ERR_PARAM;
    const uint32_t tsize = TYPESIZE[cfg->dtype];
    if (((uintptr_t)d_dst | dst_tile_pitch) % tsize) return QB3CU_ERR_PARAM;
    const uint64_t line = (uint64_t)cfg->width * cfg->bands;
    if (cfg->stride && cfg->stride < line) return QB3CU_ERR_PARAM;
    DecArgs a;
    memset(&a, 0, sizeof(a));
    a.streams = static_cast<const uint8_t *>(d_streams);
    a.offsets = reinterpret_cast<const unsigned long long *>(d_offsets);
    a.lens = reinterpret_cast<const unsigned long long *>(d_lens);
    a.dst = static_cast<uint8_t *>(d_dst);
    a.dst_pitch = dst_tile_pitch;
    a.status = d_status;
    a.stride = cfg->stride ? cfg->stride : line;
    a.w = cfg->width; a.h = cfg->height; a.bands = cfg->bands; a.dtype = cfg->dtype;
    a.ref_compat = ref_compat != 0;
    a.ntiles = (uint32_t)ntiles;
    a.rle_hint = rle_requested(cfg->mode);
    a.shared_sm = shared_sm;
    uint32_t launches = 0;
    cudaError_t err = launch_decode(a, tsize, static_cast<cudaStream_t>(stream), launches);
    count_launches(launches);
    return note_cuda(err);
}

int qb3::decode_batch_shared(const qb3cu_config *cfg, const void *d_streams, const uint64_t *d_offsets, const uint64_t *d_lens,
                             void *d_dst, size_t dst_tile_pitch, uint32_t *d_status, int ref_compat, size_t ntiles, void *stream)
{
    return decode_impl(cfg, d_streams, d_offsets, d_lens, d_dst, dst_tile_pitch, d_status, ref_compat, ntiles, stream, true);
}

extern "C" {

int qb3cu_decode_batch(const qb3cu_config *cfg, const void *d_streams, const uint64_t *d_offsets,
                       const uint64_t *d_lens, void *d_dst, size_t dst_tile_pitch, uint32_t *d_status,
                       int ref_compat, size_t ntiles, void *stream)
{
    return decode_impl(cfg, d_streams, d_offsets, d_lens, d_dst, dst_tile_pitch, d_status, ref_compat, ntiles, stream, false);
}

int qb3cu_pack_streams(const void *d_slots, size_t slot_bytes, const uint64_t *d_sizes, void *d_packed,
                       uint64_t *d_offsets, uint64_t *d_total, size_t ntiles, void *stream)
{
    if (!d_slots || !d_sizes || !d_packed || !d_offsets || !d_total) return QB3CU_ERR_PARAM;
    if (((uintptr_t)d_slots | (uintptr_t)d_packed | slot_bytes) % 16 || ntiles > 0x7fffffffull) return QB3CU_ERR_PARAM;
    if (ntiles == 0) return QB3CU_OK;
    cudaError_t err = launch_pack(static_cast<const uint8_t *>(d_slots), slot_bytes,
                                  reinterpret_cast<const unsigned long long *>(d_sizes), static_cast<uint8_t *>(d_packed),
                                  reinterpret_cast<unsigned long long *>(d_offsets),
                                  reinterpret_cast<unsigned long long *>(d_total), (uint32_t)ntiles,
                                  static_cast<cudaStream_t>(stream));
    if (err == cudaSuccess) count_launches(1 + (ntiles + 65534) / 65535);
    return note_cuda(err);
}

} /* extern "C" */
