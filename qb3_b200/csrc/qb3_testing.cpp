/*
 * qb3_testing.cpp -- probes for the CPU-only tests (tests/test_host_logic.py): the closed forms the kernels use
 * (qb3_codes.h) and the header writer (qb3_host.h), callable without a GPU. Built into its own small library,
 * qb3_b200/libqb3cu_testing.so; nothing of it is part of libQB3.so.
 */
#include "qb3_host.h"

using namespace qb3;

extern "C" {


LIBQB3_EXPORT uint32_t qb3cu_debug_cs_entry(uint32_t U, uint32_t d) { return cs_entry(U, d); }
LIBQB3_EXPORT uint32_t qb3cu_debug_cs_signal(uint32_t U) { return cs_signal(U); }
LIBQB3_EXPORT uint32_t qb3cu_debug_ds_entry(uint32_t U, uint32_t x) { return ds_entry(U, x); }
/* (len << 12) | bits of a stand-alone value at a rung below 11, the reference's CRG table entry */
LIBQB3_EXPORT uint32_t qb3cu_debug_code(uint32_t rung, uint32_t v, int group)
{
    uint64_t lo; uint32_t hi;
    if (rung == 0) return 0x1000u | (v & 1);
    if (group ? group_swaps(rung) : single_swaps(rung)) v = mswap<uint32_t>(v, rung);
    const uint32_t len = code_bits<uint32_t>(v, rung, lo, hi);
    return (len << 12) | (uint32_t)lo;
}
LIBQB3_EXPORT uint32_t qb3cu_debug_decode(uint32_t rung, uint32_t x, int group)
{
    uint32_t len;
    if (rung == 0) return 0x1000u | (x & 1);
    uint32_t v = (uint32_t)decode_bits(x, 0, rung, len);
    if (group ? group_swaps(rung) : single_swaps(rung)) v = mswap<uint32_t>(v, rung);
    return (len << 12) | v;
}
LIBQB3_EXPORT int qb3cu_debug_step(uint32_t M, int decode) { return decode ? step_decode_index(M) : step_encode_index(M); }
LIBQB3_EXPORT uint32_t qb3cu_debug_headers(const qb3cu_config *cfg, uint32_t mode_byte, uint8_t *out)
{
    const uint64_t order = cfg->order ? cfg->order : (cfg->mode <= 3 ? ZCURVE : HILBERT);
    return build_headers(cfg, mode_byte, order, out);
}


} /* extern "C" */
