/*
 * QB3.h -- public C API of the B200-native QB3 codec.
 *
 * This header declares, unchanged in name, argument order, types and enum
 * values, the 21 functions of the reference library's public header
 * (reference: QB3lib/QB3.h:85-162) so that callers written against libQB3
 * (cqb3.cpp, GDAL frmts/mrf) compile and link against this library without
 * edits. The implementation behind it is hand-written CUDA for sm_100a; there
 * is no CPU codec in the library.
 *
 * Batched, device-pointer entry points for many independent tiles live in
 * qb3cu.h; the functions below are batch-of-one wrappers around them that
 * move caller (host) buffers to and from the GPU.
 */
#ifndef QB3_B200_QB3_H
#define QB3_B200_QB3_H

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

/*
 * Upper limit on the band count; sizes the band-map arrays callers pass to qb3_set_encoder_coreband /
 * qb3_get_coreband. 16 as in the reference (QB3.h:33-34, which allows "up to 256"): qb3_create_encoder and
 * qb3_read_start refuse more bands than that, exactly like the reference library, so that a caller built with
 * size_t bandmap[16] is never handed more. A program that wants more compiles with -DQB3_MAXBANDS=n (n <= 256,
 * the kernels' own limit) and tells the library once, before its first handle: qb3cu_api_max_bands(n) in qb3cu.h.
 */
#if !defined(QB3_MAXBANDS)
#define QB3_MAXBANDS 16
#endif

/* opaque handles (reference: QB3.h:36-37) */
typedef struct encs *encsp;
typedef struct decs *decsp;

/* pixel value types; signed types are coded through their unsigned bit pattern (reference: QB3.h:40) */
enum qb3_dtype {
    QB3_U8 = 0, QB3_I8, QB3_U16, QB3_I16, QB3_U32, QB3_I32, QB3_U64, QB3_I64
};

#define QB3_HAS_FTL 1

/* encoding modes (reference: QB3.h:50-74); values are part of the file format */
enum qb3_mode {
    QB3M_DEFAULT = 8,    /* alias of QB3M_FTL */
    QB3M_BASE = 4,       /* alias of QB3M_BASE_H */
    QB3M_BEST = 7,       /* alias of QB3M_CF_RLE_H */

    QB3M_BASE_Z = 0,     /* legacy Z-curve: step coding */
    QB3M_CF = 1,         /*   + common factor / index groups */
    QB3M_RLE = 2,        /*   + zero-run byte pass */
    QB3M_CF_RLE = 3,     /*   + both */

    QB3M_BASE_H = 4,     /* Hilbert curve: step coding */
    QB3M_CF_H = 5,       /*   + common factor / index groups */
    QB3M_RLE_H = 6,      /*   + zero-run byte pass */
    QB3M_CF_RLE_H = 7,   /*   + both */

    QB3M_FTL = 8,        /* Hilbert curve, no step coding: fastest */
    QB3M_END,            /* one past the last mode a caller may request */

    QB3M_STORED = 255,   /* raw pixels after the headers; chosen by the encoder only */
    QB3M_INVALID = -1
};

/* error codes reported by qb3_get_encoder_state (reference: QB3.h:77-83) */
enum qb3_error {
    QB3E_OK = 0,
    QB3E_EINV,           /* invalid parameter */
    QB3E_UNKN,           /* unknown chunk */
    QB3E_ERR,            /* unspecified */
    QB3E_LIBERR = 255    /* internal, includes CUDA runtime failures */
};

/* ---- encoder (reference: QB3.h:88-127, QB3encode.cpp) ---- */

/* Returns NULL for w or h outside 1..65536, bands outside 1..QB3_MAXBANDS or a bad type. */
LIBQB3_EXPORT encsp qb3_create_encoder(size_t width, size_t height, size_t bands, qb3_dtype dt);
LIBQB3_EXPORT void qb3_destroy_encoder(encsp p);
/* Clears the per-band running state and the error, keeps the settings. */
LIBQB3_EXPORT void qb3_reset_encoder(encsp p);
/* cband[i] is the band subtracted from band i (i itself = none). Default: identity,
   or {1,1,1[,3]} for 3 or 4 bands. The array is rewritten with the map actually used.
   Returns false only when bands differs from the encoder's band count. */
LIBQB3_EXPORT bool qb3_set_encoder_coreband(encsp p, size_t bands, size_t *cband);
/* Lossy: values are divided by q (rounded, away from zero if asked) before coding. */
LIBQB3_EXPORT bool qb3_set_encoder_quanta(encsp p, uint64_t q, bool away);
/* Size the destination of qb3_encode with this. */
LIBQB3_EXPORT size_t qb3_max_encoded_size(const encsp p);
/* Returns the mode in effect afterwards (unchanged when mode is out of range). */
LIBQB3_EXPORT qb3_mode qb3_set_encoder_mode(encsp p, qb3_mode mode);
/* Line to line distance of the source in values; 0 = width * bands. */
LIBQB3_EXPORT void qb3_set_encoder_stride(encsp p, size_t stride);
/* source: host pointer, row major, band interleaved. destination: host pointer with at
   least qb3_max_encoded_size bytes. Returns the stream length in bytes, 0 on error. */
LIBQB3_EXPORT size_t qb3_encode(encsp p, void *source, void *destination);
LIBQB3_EXPORT int qb3_get_encoder_state(encsp p);

/* ---- decoder (reference: QB3.h:132-162, QB3decode.cpp) ---- */

/* Parses the fixed header; image_size receives width, height, bands. source must stay
   valid until qb3_read_data returns and source_size must be the exact stream length. */
LIBQB3_EXPORT decsp qb3_read_start(void *source, size_t source_size, size_t *image_size);
/* Parses the chunks up to the data; false when the stream is malformed. */
LIBQB3_EXPORT bool qb3_read_info(decsp p);
/* Decodes into destination (host pointer, qb3_decoded_size bytes); returns that size or 0. */
LIBQB3_EXPORT size_t qb3_read_data(decsp p, void *destination);
LIBQB3_EXPORT void qb3_destroy_decoder(decsp p);
LIBQB3_EXPORT size_t qb3_decoded_size(const decsp p);
LIBQB3_EXPORT qb3_dtype qb3_get_type(const decsp p);
/* Line to line distance of the destination in values; 0 = width * bands. */
LIBQB3_EXPORT void qb3_set_decoder_stride(decsp p, size_t stride);
/* The following are valid after qb3_read_info. */
LIBQB3_EXPORT qb3_mode qb3_get_mode(const decsp p);
LIBQB3_EXPORT uint64_t qb3_get_quanta(const decsp p);
LIBQB3_EXPORT uint64_t qb3_get_order(const decsp p);
LIBQB3_EXPORT bool qb3_get_coreband(const decsp p, size_t *cband);

#if defined(__cplusplus)
}
#endif
#endif /* QB3_B200_QB3_H */
