/*
 * qb3_oracle.h -- TEST INFRASTRUCTURE ONLY.
 *
 * Plain-C, single-threaded restatement of the QB3 codec used as the parity
 * oracle for the CUDA path. Nothing in the product (qb3_b200/, include/)
 * may include, link or call this. Only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline leg use it, and only as the checker.
 *
 * Parity status: PINNED. The restatement is checked byte-for-byte against
 * the reference library compiled from /root/reference (oracle/_ref, see
 * oracle/Makefile) by tests/test_oracle_vs_ref.py, and against the committed
 * golden vectors in tests/golden/ that were generated from that library.
 */
#ifndef QB3_ORACLE_H
#define QB3_ORACLE_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define QB3O_MAXBANDS 256

/* Encoder settings + running state; mirrors struct encs (QB3common.h:68-88) */
typedef struct {
    size_t xsize, ysize, nbands;
    size_t stride;      /* line stride in values, 0 = xsize * nbands */
    uint64_t order;     /* 0 = Hilbert */
    uint64_t quanta;
    int away;
    int mode;           /* qb3_mode value 0..8 */
    int type;           /* qb3_dtype value 0..7 */
    int error;
    uint8_t cband[QB3O_MAXBANDS];
    /* running state per band (QB3common.h:63) */
    uint64_t prev[QB3O_MAXBANDS], runbits[QB3O_MAXBANDS], cf[QB3O_MAXBANDS];
} qb3o_enc;

/* Fill defaults the way qb3_create_encoder does (QB3encode.cpp:26-48); 0 on success */
int qb3o_init(qb3o_enc *e, size_t w, size_t h, size_t bands, int type);
/* qb3_set_encoder_mode semantics (QB3encode.cpp:120-134) */
int qb3o_set_mode(qb3o_enc *e, int mode);
/* qb3_set_encoder_coreband semantics (QB3encode.cpp:63-77) */
int qb3o_set_coreband(qb3o_enc *e, size_t bands, size_t *cband);
void qb3o_reset(qb3o_enc *e);
size_t qb3o_max_encoded_size(const qb3o_enc *e);
/* Full container encode, qb3_encode semantics (QB3encode.cpp:488-574) */
size_t qb3o_encode(qb3o_enc *e, const void *src, void *dst);

/* Parsed stream header; mirrors struct decs (QB3common.h:91-111) */
typedef struct {
    size_t xsize, ysize, nbands;
    uint64_t order, quanta;
    int mode, type;
    int has_cb;                 /* a CB chunk was present */
    uint8_t cband[QB3O_MAXBANDS];
    size_t data_offset;         /* byte offset of the payload after "DT" */
} qb3o_info;

/* Header + chunk parse (QB3decode.cpp:130-264). Returns 0 on success. */
int qb3o_read_info(const void *src, size_t len, qb3o_info *info);
/*
 * Full container decode (QB3decode.cpp:380-452). stride in values, 0 = default.
 * identity_default != 0: missing CB chunk means identity band map (the format
 * spec, doc/QB3.md:255); 0: replicate the reference decoder, which leaves the
 * map zeroed. Returns decoded byte count, 0 on failure.
 */
size_t qb3o_decode(const void *src, size_t len, void *dst, size_t stride, int identity_default);

/* table access for table tests: closed forms of CRG/DRG/csw/dsw (QB3encode.h:25-89, QB3decode.h:24-116) */
uint16_t qb3o_crg(unsigned rung, unsigned v);
uint16_t qb3o_drg(unsigned rung, unsigned x);
uint16_t qb3o_csw(unsigned ubits, unsigned d);
uint16_t qb3o_dsw(unsigned ubits, unsigned x);
uint16_t qb3o_signal(unsigned ubits);

#ifdef __cplusplus
}
#endif
#endif
