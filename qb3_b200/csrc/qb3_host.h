/*
 * qb3_host.h -- host side pieces shared by the C ABI (qb3_cabi.cu) and the test probes (qb3_testing.cpp):
 * the stream header writer.
 */
#ifndef QB3_B200_HOST_H
#define QB3_B200_HOST_H

#include <stdint.h>

#include "qb3_codes.h"
#include "../../include/qb3cu.h"

namespace qb3 {

/* Header bytes up to and including "DT" (reference: QB3encode.cpp:189-268, doc/QB3.md:228-259) */
inline uint32_t build_headers(const qb3cu_config *c, uint32_t mode_byte, uint64_t order, uint8_t *out)
{
    uint32_t n = 0;
    auto put = [&](uint64_t v, uint32_t bytes) { for (uint32_t i = 0; i < bytes; i++) out[n++] = (uint8_t)(v >> (8 * i)); };
    put(0x80334251u, 4); /* "QB3\200" */
    put(c->width - 1, 2);
    put(c->height - 1, 2);
    put(c->bands - 1, 1);
    put(c->dtype, 1);
    put(mode_byte, 1);
    bool banddiff = false;
    for (uint32_t b = 0; b < c->bands; b++) banddiff |= c->cband[b] != b;
    if (mode_byte != M_STORED && banddiff) {
        put('C' | ('B' << 8), 2);
        put(c->bands, 2);
        for (uint32_t b = 0; b < c->bands; b++) put(c->cband[b], 1);
    }
    if (c->quanta >= 2) {
        const uint32_t qbytes = 1 + topbit64(c->quanta) / 8;
        put('Q' | ('V' << 8), 2);
        put(qbytes, 2);
        put(c->quanta, qbytes);
    }
    if (order != ZCURVE && mode_byte != M_STORED) {
        put('S' | ('C' << 8), 2);
        put(8, 2);
        put(order, 8);
    }
    put('D' | ('T' << 8), 2);
    return n;
}
} // namespace qb3
#endif
