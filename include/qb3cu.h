/*
 * qb3cu.h -- batched device-pointer C ABI of the B200-native QB3 codec.
 *
 * These are the entry points a host language binds (cgo / JNI / ctypes) when it already has
 * tiles in GPU memory; the QB3.h functions are batch-of-one wrappers over them. Plain C types
 * only: device pointers are void*, the stream is the cudaStream_t handle passed as void*
 * (NULL = the legacy default stream). Every call is asynchronous with respect to the host and
 * enqueues its kernels on that stream; results are in device memory.
 *
 * What each one replaces in the reference:
 *   qb3cu_encode_batch  <- one qb3_encode call per tile (QB3encode.cpp:488-574), i.e. the loops
 *                          QB3::encode_fast / encode_best (QB3encode.h:376-451, 617-724), the header
 *                          writer (QB3encode.cpp:189-268), quantize (:151-186), the small-image
 *                          reorder (:351-389), RLE0 (:280-332) and the stored fallback (:461-485)
 *   qb3cu_decode_batch  <- one qb3_read_start + qb3_read_info + qb3_read_data per tile
 *                          (QB3decode.cpp:130-264, 380-464), i.e. QB3::decode / decodeFTL
 *                          (QB3decode.h:293-741), deRLE0 (QB3decode.cpp:267-307), dequantize (:77-107)
 *   qb3cu_max_encoded_size <- qb3_max_encoded_size (QB3encode.cpp:112-118)
 *
 * There is no CPU implementation behind these: without a CUDA device they return QB3CU_ERR_CUDA.
 */
#ifndef QB3_B200_QB3CU_H
#define QB3_B200_QB3CU_H

#include <stddef.h>
#include <stdint.h>

#if !defined(LIBQB3_EXPORT)
#if defined(__GNUC__)
#define LIBQB3_EXPORT __attribute__((visibility("default")))
#else
#define LIBQB3_EXPORT
#endif
#endif

#if defined(__cplusplus)
extern "C" {
#endif

#define QB3CU_MAXBANDS 256

/* return codes of the host-side calls */
enum {
    QB3CU_OK = 0,
    QB3CU_ERR_PARAM = 1,   /* bad geometry, type, mode, band map, pitch or alignment */
    QB3CU_ERR_CUDA = 2     /* no device / CUDA runtime failure; see qb3cu_last_cuda_error */
};

/* per tile status words written by the kernels */
enum {
    QB3CU_TILE_OK = 0,
    QB3CU_TILE_BAD_HEADER = 1,   /* decode: not a QB3 stream, or it does not match the expected geometry */
    QB3CU_TILE_CORRUPT = 2,      /* decode: the reference's failure conditions (QB3decode.h:642,665,683,703,740) */
    QB3CU_TILE_RLE_TOO_BIG = 3   /* decode: expanded RLE payload larger than the raw size (QB3decode.cpp:401) */
};

/* Settings shared by all tiles of a batch: what the reference keeps in struct encs (QB3common.h:68-88). */
typedef struct qb3cu_config {
    uint32_t width, height, bands; /* 1..65536, 1..65536, 1..256 */
    uint32_t dtype;                /* qb3_dtype */
    uint32_t mode;                 /* qb3_mode as a caller would request it: 0..8 */
    uint32_t away;                 /* quantisation rounds half away from zero */
    uint64_t quanta;               /* >= 1; 1 = lossless */
    uint64_t order;                /* 4x4 scan curve, 0 = the mode's default (Hilbert, Z for modes 0..3) */
    uint64_t stride;               /* line to line distance in values, 0 = width * bands */
    uint8_t cband[QB3CU_MAXBANDS]; /* cband[c] = band subtracted from band c, c itself for none */
} qb3cu_config;

/* Fills width/height/bands/dtype and the reference defaults: FTL, quanta 1, identity band map or
   {1,1,1[,3]} for 3 or 4 bands (QB3encode.cpp:36-45). Returns QB3CU_ERR_PARAM for bad arguments. */
LIBQB3_EXPORT int qb3cu_config_init(qb3cu_config *cfg, uint32_t width, uint32_t height, uint32_t bands, uint32_t dtype);

/* Identical to qb3_max_encoded_size for the same geometry. */
LIBQB3_EXPORT size_t qb3cu_max_encoded_size(const qb3cu_config *cfg);

/* Bytes to reserve per tile in the destination of qb3cu_encode_batch: the value above rounded up
   for 16-byte vector stores plus one spare vector. Slots must start 16-byte aligned. */
LIBQB3_EXPORT size_t qb3cu_slot_bytes(const qb3cu_config *cfg);

/*
 * Encodes ntiles independent tiles, one QB3 stream each, byte-identical to qb3_encode.
 *   d_src            tile t starts at d_src + t * src_tile_pitch bytes; row major, band interleaved
 *   d_dst            stream t is written at d_dst + t * dst_slot_bytes (>= qb3cu_slot_bytes, multiple of 16)
 *   d_sizes[t]       stream length in bytes
 *   d_status[t]      QB3CU_TILE_OK, may be NULL
 *   d_state          NULL, or ntiles * 3 * bands uint64: per band {prev, runbits, cf} read before and written
 *                    after each tile -- the running state the reference keeps in the handle (QB3common.h:63)
 */
LIBQB3_EXPORT int qb3cu_encode_batch(const qb3cu_config *cfg, const void *d_src, size_t src_tile_pitch,
                                     void *d_dst, size_t dst_slot_bytes, uint64_t *d_sizes, uint32_t *d_status,
                                     uint64_t *d_state, size_t ntiles, void *stream);

/*
 * The sizes qb3cu_encode_batch would report, without the streams: d_sizes[t] = bytes of tile t's stream (stored
 * fallback included). Replaces the encodes a caller runs only to compare sizes -- cqb3 -m x tries ten band maps and
 * keeps the smallest (cqb3.cpp:561-586). The encode kernel runs with its packing and stores left out and needs no
 * destination; with an RLE mode (2, 3, 6, 7) the streams have to exist for the byte pass to be measured, so they are
 * made in scratch memory of the library's own.
 */
LIBQB3_EXPORT int qb3cu_encoded_size_batch(const qb3cu_config *cfg, const void *d_src, size_t src_tile_pitch,
                                           uint64_t *d_sizes, size_t ntiles, void *stream);

/*
 * Decodes ntiles streams whose headers must all describe width x height x bands of dtype (mode, quanta,
 * order and cband are read from each stream's own header, cfg->stride is the output stride). cfg->mode is
 * only a hint: when it names an RLE mode (2, 3, 6, 7 -- QB3M_BEST is 7), room is set aside to expand RLE
 * streams ahead of the parallel decode; without the hint RLE streams still decode, on a slow path.
 *   d_streams        base pointer; stream t occupies [d_offsets[t], d_offsets[t] + d_lens[t])
 *   d_dst            tile t is written at d_dst + t * dst_tile_pitch bytes
 *   d_status[t]      QB3CU_TILE_* ; the tile content is undefined unless QB3CU_TILE_OK
 *   ref_compat       nonzero: a stream without a CB chunk decodes with an all-zero band map, as the
 *                    reference decoder does; zero: identity map, as the format specifies
 */
LIBQB3_EXPORT int qb3cu_decode_batch(const qb3cu_config *cfg, const void *d_streams, const uint64_t *d_offsets,
                                     const uint64_t *d_lens, void *d_dst, size_t dst_tile_pitch,
                                     uint32_t *d_status, int ref_compat, size_t ntiles, void *stream);

/*
 * Packs the streams of qb3cu_encode_batch back to back, what a caller wants before moving them off the device or
 * into a tile index (MRF style): stream t goes to d_packed + d_offsets[t], every start rounded up to 16 bytes;
 * d_total[0] = bytes used. d_packed needs room for the sum of the sizes plus 16 bytes per tile (at most
 * ntiles * slot_bytes). The result feeds qb3cu_decode_batch directly (d_streams = d_packed, d_offsets, d_sizes).
 * (The reference hands one buffer per qb3_encode call back to its caller, cqb3.cpp:478-493; this is that step
 * for a batch.)
 */
LIBQB3_EXPORT int qb3cu_pack_streams(const void *d_slots, size_t slot_bytes, const uint64_t *d_sizes, void *d_packed,
                                     uint64_t *d_offsets, uint64_t *d_total, size_t ntiles, void *stream);

/*
 * ---- Host buffer pipeline: tiles in host memory in, packed streams in host memory out, and back ----
 *
 * What a caller of the reference does around qb3_encode / qb3_read_data for many tiles (the per image loops of
 * cqb3.cpp:405-493 and :276-323, the tile loop of GDAL's MRF driver): here one call takes the whole batch. It is cut
 * into chunks of tiles; host to device copies, kernels and device to host copies of different chunks overlap on
 * several CUDA streams; streams are packed on the device so that only the bytes produced cross PCIe. A pipe owns its
 * device staging memory, CUDA streams and pinned index buffers and keeps them between calls. One call at a time per
 * pipe; different pipes may be used from different threads at once (an encode and a decode then share both PCIe
 * directions). Host buffers should be page locked (qb3cu_host_alloc, cudaHostAlloc / cudaHostRegister, torch
 * pin_memory) -- pageable memory works but is copied synchronously by the driver.
 */
typedef struct qb3cu_pipe qb3cu_pipe;

/* chunk_tiles: tiles per chunk, 0 = chosen from the tile size (about 200 MB of pixels); depth: chunks in flight,
   0 = default (6: a chunk's decode takes about as long as five chunks' copies). Returns NULL for bad arguments or when
   there is no CUDA device. */
LIBQB3_EXPORT qb3cu_pipe *qb3cu_pipe_create(const qb3cu_config *cfg, size_t chunk_tiles, int depth);
LIBQB3_EXPORT void qb3cu_pipe_destroy(qb3cu_pipe *pipe);

/*
 * Encodes ntiles tiles from host memory (tile t at h_src + t * src_tile_pitch, laid out as for qb3cu_encode_batch).
 * Stream t, byte identical to qb3_encode's, lands at h_packed + h_offsets[t] (16 byte aligned starts, in tile order,
 * the first at 0) with length h_sizes[t]; *h_total = bytes of h_packed used. packed_capacity: bytes available in
 * h_packed; ntiles * qb3cu_slot_bytes() always suffices. Returns QB3CU_ERR_PARAM when the capacity is exceeded.
 * Synchronous: everything is in host memory on return.
 */
LIBQB3_EXPORT int qb3cu_pipe_encode(qb3cu_pipe *pipe, const void *h_src, size_t src_tile_pitch, void *h_packed,
                                    size_t packed_capacity, uint64_t *h_offsets, uint64_t *h_sizes, uint64_t *h_total,
                                    size_t ntiles);

/*
 * Decodes ntiles streams from host memory (stream t = h_lens[t] bytes at h_streams + h_offsets[t]; the layout
 * qb3cu_pipe_encode produces, or any other with the streams of a chunk of consecutive tiles reasonably close together)
 * into h_dst + t * dst_tile_pitch. h_status[t] = QB3CU_TILE_*. Synchronous.
 */
LIBQB3_EXPORT int qb3cu_pipe_decode(qb3cu_pipe *pipe, const void *h_streams, const uint64_t *h_offsets,
                                    const uint64_t *h_lens, void *h_dst, size_t dst_tile_pitch, uint32_t *h_status,
                                    int ref_compat, size_t ntiles);

/*
 * ---- Several devices from one process ----
 *
 * The same two calls over G devices: the batch is cut into G contiguous tile ranges, range g goes through a pipe on
 * devices[g] driven by its own host thread, nothing is exchanged between devices (tiles are independent QB3 streams,
 * SURVEY 8e). devices == NULL: devices 0 .. ndevices - 1; ndevices == 0: every device of the process. A device may
 * be named more than once (two pipes on it).
 * qb3cu_multi_encode: range g is packed into its own share of h_packed, the share starting at g * (capacity / G
 * rounded down to 16): within a range the streams are back to back as for qb3cu_pipe_encode, between ranges there is
 * a gap; h_offsets[t] / h_sizes[t] say where stream t is, *h_total = bytes of streams written (gaps not counted).
 * ntiles * qb3cu_slot_bytes() of capacity always suffices. The result feeds qb3cu_multi_decode or qb3cu_pipe_decode.
 */
typedef struct qb3cu_multi qb3cu_multi;
LIBQB3_EXPORT qb3cu_multi *qb3cu_multi_create(const qb3cu_config *cfg, const int *devices, int ndevices,
                                              size_t chunk_tiles, int depth);
LIBQB3_EXPORT void qb3cu_multi_destroy(qb3cu_multi *multi);
LIBQB3_EXPORT int qb3cu_multi_devices(const qb3cu_multi *multi);
LIBQB3_EXPORT int qb3cu_multi_encode(qb3cu_multi *multi, const void *h_src, size_t src_tile_pitch, void *h_packed,
                                     size_t packed_capacity, uint64_t *h_offsets, uint64_t *h_sizes, uint64_t *h_total,
                                     size_t ntiles);
LIBQB3_EXPORT int qb3cu_multi_decode(qb3cu_multi *multi, const void *h_streams, const uint64_t *h_offsets,
                                     const uint64_t *h_lens, void *h_dst, size_t dst_tile_pitch, uint32_t *h_status,
                                     int ref_compat, size_t ntiles);

/* Page locked host memory for the buffers above (cudaHostAlloc / cudaFreeHost). NULL on failure. */
LIBQB3_EXPORT void *qb3cu_host_alloc(size_t bytes);
LIBQB3_EXPORT void qb3cu_host_free(void *p);

/* Band limit of the QB3.h functions (qb3_create_encoder, qb3_read_start): 16 when the library is loaded, like the
   reference's QB3_MAXBANDS (QB3.h:34); a caller compiled with a larger QB3_MAXBANDS raises it here, up to
   QB3CU_MAXBANDS. Returns the limit now in force. The batched entry points above always take up to QB3CU_MAXBANDS. */
LIBQB3_EXPORT uint32_t qb3cu_api_max_bands(uint32_t bands);

/* cudaError_t of the most recent failing CUDA call made by this library on the calling thread. */
LIBQB3_EXPORT int qb3cu_last_cuda_error(void);

/* Number of kernels this library has launched since load (all threads); for benchmark bookkeeping. */
LIBQB3_EXPORT uint64_t qb3cu_kernel_launches(void);

#if defined(__cplusplus)
}
#endif
#endif /* QB3_B200_QB3CU_H */
