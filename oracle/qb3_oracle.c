/*
 * qb3_oracle.c -- TEST INFRASTRUCTURE ONLY (see qb3_oracle.h).
 *
 * CPU restatement of the QB3 codec of lucianpls/QB3, written from the format
 * description (doc/QB3.md) and the behaviour of QB3lib, group by group from the
 * closed forms of the code tables (attic/tables.py) instead of the reference's
 * table-driven accumulator loops. Each function cites the reference file:line
 * it follows. All values are carried as uint64_t masked to the type width.
 *
 * Parity status: PINNED against oracle/_ref (reference compiled from
 * /root/reference) and tests/golden/ -- see tests/test_oracle_vs_ref.py.
 */
#include "qb3_oracle.h"
#include <stdlib.h>
#include <string.h>

#define ZCURVE 0x0145236789cdabefull   /* QB3common.h:185 */
#define HILBERT 0x01548cd9aefb7623ull  /* QB3common.h:193 */
#define MODE_STORED 255
#define MODE_FTL 8

static const int TYPESIZE[8] = {1, 1, 2, 2, 4, 4, 8, 8}; /* QB3decode.cpp:25 */

/* ------------------------------------------------------------------ bits */

static unsigned topbit(uint64_t v) { return 63u - (unsigned)__builtin_clzll(v); } /* QB3common.h:42-61 */
static uint64_t wmask(unsigned bits) { return bits >= 64 ? ~0ull : ((1ull << bits) - 1); }
static unsigned ubits_of(unsigned bits) { return bits == 8 ? 3 : bits == 16 ? 4 : bits == 32 ? 5 : 6; } /* QB3encode.h:380 */

/* LSB-first bit writer; bitstream.h:66-126 (oBits). Bytes are cleared as they are entered. */
typedef struct { uint8_t *p; size_t pos; } obits;

static void put(obits *s, uint64_t v, unsigned n)
{
    while (n) {
        unsigned o = (unsigned)(s->pos & 7);
        unsigned k = 8 - o < n ? 8 - o : n;
        if (!o) s->p[s->pos >> 3] = 0;
        s->p[s->pos >> 3] |= (uint8_t)((v & ((1u << k) - 1)) << o);
        v >>= k;
        n -= k;
        s->pos += k;
    }
}

/* LSB-first bit reader, zero fill past the end, saturating advance; bitstream.h:25-63 (iBits) */
typedef struct { const uint8_t *p; size_t len, pos; } ibits;

static uint64_t peek(const ibits *s)
{
    uint64_t v = 0;
    for (unsigned i = 0; i < 9; i++) {
        size_t byte = (s->pos >> 3) + i;
        if (byte * 8 >= s->len) break;
        unsigned sh = i * 8;
        unsigned o = (unsigned)(s->pos & 7);
        if (sh >= o) { if (sh - o < 64) v |= (uint64_t)s->p[byte] << (sh - o); }
        else v |= (uint64_t)s->p[byte] >> (o - sh);
    }
    return v;
}
static void advance(ibits *s, size_t d) { s->pos = (s->pos + d < s->len) ? s->pos + d : s->len; }
static uint64_t get(ibits *s, unsigned n) { uint64_t v = peek(s) & wmask(n); advance(s, n); return v; }

/* ------------------------------------------------------------- codewords */

/* One value at a rung: short / nominal / long suffix code (QB3encode.h:132-141, tables.py:37-45).
   Up to 65 bits at rung 63: lo holds the low 64, hi the 65th. */
typedef struct { unsigned len; uint64_t lo; unsigned hi; } cw;

static cw code(uint64_t v, unsigned r)
{
    cw c = {0, 0, 0};
    if (r == 0) { c.len = 1; c.lo = v & 1; return c; }
    uint64_t half = 1ull << (r - 1), full = 1ull << r;
    if (v < half) { c.len = r; c.lo = v << 1; }
    else if (v < full) { c.len = r + 1; c.lo = ((v - half) << 2) | 1; }
    else {
        uint64_t x = v - full;
        c.len = r + 2;
        c.lo = (x << 2) | 3;
        c.hi = (unsigned)((x >> 62) & 1);
    }
    return c;
}

static void put_cw(obits *s, cw c)
{
    if (c.len <= 64) put(s, c.lo, c.len);
    else { put(s, c.lo, 64); put(s, c.hi, 1); } /* QB3encode.h:267-275 */
}

/* "middle swap": 2^r - 1 <-> 2^r (QB3encode.h:30-33) */
static uint64_t mswap(uint64_t v, unsigned r)
{
    uint64_t full = 1ull << r;
    if (v == full) return full - 1;
    if (v == full - 1) return full;
    return v;
}

/* value inside a group: rungs 1..7 swapped (inline LUTs QB3encode.h:185-197 and crg3..7), others not */
static cw gcode(uint64_t v, unsigned r) { return code((r >= 1 && r <= 7) ? mswap(v, r) : v, r); }
/* stand-alone value (cf, index table): qb3csztbl, QB3encode.h:144-150: crg0..2 unswapped, crg3..7 swapped */
static cw scode(uint64_t v, unsigned r) { return code((r >= 3 && r <= 7) ? mswap(v, r) : v, r); }

/* Inverse of code() from the low bits of x (QB3decode.h:119-129, tables.py:48-63); hi65 = bit 64 of the window */
static uint64_t dcode(uint64_t x, unsigned hi65, unsigned r, unsigned *len)
{
    if (r == 0) { *len = 1; return x & 1; }
    if (!(x & 1)) { *len = r; return (x & wmask(r)) >> 1; }
    if (!(x & 2)) { *len = r + 1; return ((x & wmask(r + 1)) >> 2) + (1ull << (r - 1)); }
    *len = r + 2;
    uint64_t v = (r + 2 <= 64) ? (x & wmask(r + 2)) >> 2 : (x >> 2) | ((uint64_t)hi65 << 62);
    return v + (1ull << r); /* at r == 63, v < 2^63 so this is v | 2^63 */
}
static uint64_t gdcode(uint64_t x, unsigned hi65, unsigned r, unsigned *len)
{
    uint64_t v = dcode(x, hi65, r, len);
    return (r >= 1 && r <= 7) ? mswap(v, r) : v;
}
static uint64_t sdcode(uint64_t x, unsigned r, unsigned *len)
{
    uint64_t v = dcode(x, 0, r, len);
    return (r >= 3 && r <= 7) ? mswap(v, r) : v;
}

/* mag-sign folding in 8 bits, used by the code switch construction (tables.py:5-7) */
static unsigned mags8(int v) { unsigned u = (unsigned)v & 0xff; return 0xff & ((0xff * (u >> 7)) ^ (u << 1)); }
static int smag8(unsigned v) { return (v & 1) ? -(int)((v >> 1) + 1) : (int)(v >> 1); }

/* Code switch entry for a rung delta d, (len << 12) | bits including the change flag; tables.py:115-133 */
uint16_t qb3o_csw(unsigned U, unsigned d)
{
    if (d == 0) return 0x1000;
    unsigned sb = 1u << (U - 1);
    unsigned m = (d & sb) ? mags8((int)d - (int)(2 * sb)) : mags8((int)((d - 1) & (sb - 1)));
    cw c = code(m, U - 1);
    return (uint16_t)(((c.len + 1) << 12) | ((c.lo << 1) & 0xfff) | 1);
}

/* The otherwise unused "+max" delta, tables.py:179-187, QB3encode.h:286 */
uint16_t qb3o_signal(unsigned U)
{
    unsigned sb = 1u << (U - 1);
    cw c = code(mags8((int)(sb - 1)), U - 1);
    return (uint16_t)(((c.len + 1) << 12) | ((c.lo << 1) & 0xfff) | 1);
}

/* Decoded code switch, indexed by the U+1 bits after the change flag; tables.py:137-151 */
uint16_t qb3o_dsw(unsigned U, unsigned x)
{
    unsigned len;
    uint64_t v = dcode(x, 0, U - 1, &len);
    int s = smag8((unsigned)v & 0xff);
    unsigned d = s >= 0 ? ((unsigned)(s + 1) & ((1u << (U - 1)) - 1)) : ((unsigned)s & ((1u << U) - 1));
    return (uint16_t)(((len + 1) << 12) | d);
}

/* The reference's CRG/DRG tables from the closed form (QB3encode.h:25-74, QB3decode.h:24-95) */
uint16_t qb3o_crg(unsigned r, unsigned v)
{
    cw c = scode(v, r);
    return (uint16_t)((c.len << 12) | c.lo);
}
uint16_t qb3o_drg(unsigned r, unsigned x)
{
    unsigned len;
    uint64_t v = sdcode(x & wmask(r + 2), r, &len);
    return (uint16_t)((len << 12) | v);
}

/* Emit a rung change. with_flag: normal switch; otherwise the flag bit is dropped and
   "no change" is spelled as SIGNAL (QB3encode.h:300-305, 581-592) */
static void put_cs(obits *s, unsigned U, unsigned d, int with_flag)
{
    uint16_t e = qb3o_csw(U, d & ((1u << U) - 1));
    if (with_flag) { put(s, e & 0xfff, e >> 12); return; }
    if ((e >> 12) == 1) e = qb3o_signal(U);
    put(s, (e & 0xfff) >> 1, (e >> 12) - 1);
}

/* ---------------------------------------------------------- group encode */

static uint64_t mags_w(uint64_t v, unsigned bits) /* QB3common.h:127-129 */
{
    v &= wmask(bits);
    return ((v << 1) ^ (0 - (v >> (bits - 1)))) & wmask(bits);
}
static uint64_t smag_w(uint64_t v, unsigned bits) /* QB3common.h:133-135 */
{
    return ((v >> 1) ^ (0 - (v & 1))) & wmask(bits);
}
static uint64_t magsabs(uint64_t v) { return (v >> 1) + (v & 1); } /* QB3encode.h:92 */

/* The 16 values of a group, after the rung switch (QB3encode.h:155-280). */
static void put_body(obits *s, const uint64_t m[16], uint64_t bitsused, int use_step)
{
    if (bitsused <= 1) { /* QB3encode.h:159-166 */
        put(s, bitsused, 1);
        if (bitsused)
            for (int i = 0; i < 16; i++) put(s, m[i], 1);
        return;
    }
    unsigned r = topbit(bitsused);
    uint64_t g[16];
    memcpy(g, m, sizeof(g));
    if (use_step) { /* step-down of the rung bits: QB3common.h:141-150, QB3encode.h:169-176 */
        unsigned M = 0;
        for (int i = 0; i < 16; i++) M |= (unsigned)((g[i] >> r) & 1) << i;
        if ((M & (M + 1)) == 0)
            g[__builtin_popcount(M) - 1] ^= 1ull << r;
    }
    for (int i = 0; i < 16; i++) put_cw(s, gcode(g[i], r));
}

/* gcd of the non-zero magnitudes; QB3encode.h:98-126 */
static uint64_t gcf(const uint64_t m[16])
{
    uint64_t g = 0;
    for (int i = 0; i < 16; i++) {
        uint64_t a = magsabs(m[i]);
        while (a) { uint64_t t = g % a; g = a; a = t; }
        if (g == 1) return 1;
    }
    return g;
}

/* Common factor group; QB3encode.h:283-361 */
static void put_cf(obits *s, const uint64_t m[16], uint64_t cf, uint64_t pcf, unsigned oldrung, unsigned U, unsigned bits)
{
    uint16_t sig = qb3o_signal(U);
    uint64_t q[16], qbits = 0;
    put(s, sig & 0xfff, sig >> 12);
    for (int i = 0; i < 16; i++)
        qbits |= q[i] = (((magsabs(m[i]) / cf) << 1) - (m[i] & 1)) & wmask(bits); /* magsdiv, QB3encode.h:95 */
    uint64_t c = (cf - 2) & wmask(bits);
    unsigned trung = topbit(qbits | 1), cfrung = topbit(c | 1);
    put_cs(s, U, trung - oldrung, 0);
    if (c != pcf) {
        put(s, 1, 1);
        if (trung >= cfrung && (trung < cfrung + U || cfrung == 0)) {
            put(s, 0, 1);
            if (trung == 0) { /* QB3encode.h:312-318 */
                put(s, c, 1);
                for (int i = 0; i < 16; i++) put(s, q[i], 1);
                return;
            }
            put_cw(s, scode(c, trung));
        }
        else { /* QB3encode.h:324-347 */
            put_cs(s, U, cfrung - trung, 1);
            put_cw(s, scode(c ^ (1ull << cfrung), cfrung - 1));
            if (trung == 0) {
                for (int i = 0; i < 16; i++) put(s, q[i], 1);
                return;
            }
        }
    }
    else {
        put(s, 0, 1);
        if (trung == 0) {
            for (int i = 0; i < 16; i++) put(s, q[i], 1);
            return;
        }
    }
    put_body(s, q, qbits, 1);
}

/* Index group; QB3encode.h:557-613. Returns the bit length, 800 if more than 8 distinct values */
static size_t put_index(obits *s, const uint64_t m[16], unsigned rung, unsigned oldrung, unsigned U)
{
    uint64_t val[8];
    unsigned cnt[8], n = 0;
    for (int i = 0; i < 16; i++) {
        unsigned j = 0;
        while (j < n && val[j] != m[i]) j++;
        if (j == n) {
            if (n == 8) return 800;
            val[n] = m[i]; cnt[n++] = 1;
        }
        else cnt[j]++;
    }
    /* stable, descending count; QB3encode.h:546-554 */
    for (unsigned i = 1; i < n; i++)
        for (unsigned j = i; j > 0 && cnt[j] > cnt[j - 1]; j--) {
            uint64_t tv = val[j]; val[j] = val[j - 1]; val[j - 1] = tv;
            unsigned tc = cnt[j]; cnt[j] = cnt[j - 1]; cnt[j - 1] = tc;
        }
    size_t start = s->pos;
    uint16_t sig = qb3o_signal(U);
    unsigned MASK = (1u << U) - 1;
    put(s, sig & 0xfff, sig >> 12);
    put_cs(s, U, MASK - oldrung, 0);
    put_cs(s, U, rung - oldrung, 0);
    for (int i = 0; i < 16; i++) {
        unsigned j = 0;
        while (val[j] != m[i]) j++;
        put_cw(s, code(j, 2)); /* no middle swap, QB3encode.h:599-601 */
    }
    for (unsigned j = 0; j < n; j++) put_cw(s, scode(val[j], rung));
    return s->pos - start;
}

typedef struct {
    size_t w, h, bands;
    const uint8_t *cband;
    uint64_t order;
    unsigned bits;
    int use_step, best;
    uint64_t *prev, *runbits, *pcf; /* per band running state, in/out */
} stream_cfg;

/* Encode a whole (w, h >= 4) image already converted to contiguous uint64 values.
   Restates encode_fast (QB3encode.h:376-451) and encode_best (QB3encode.h:617-724). */
static void encode_stream(const uint64_t *img, const stream_cfg *k, obits *s)
{
    const unsigned bits = k->bits, U = ubits_of(bits);
    const uint64_t WM = wmask(bits);
    size_t nbx = (k->w + 3) / 4, nby = (k->h + 3) / 4;
    uint8_t scratch_a[256], scratch_b[256];
    for (size_t by = 0; by < nby; by++) {
        size_t y0 = by * 4 + 4 > k->h ? k->h - 4 : by * 4;
        for (size_t bx = 0; bx < nbx; bx++) {
            size_t x0 = bx * 4 + 4 > k->w ? k->w - 4 : bx * 4;
            for (size_t c = 0; c < k->bands; c++) {
                uint64_t m[16], bitsused = 0, prv = k->prev[c];
                size_t cb = k->cband[c];
                for (int i = 0; i < 16; i++) {
                    unsigned n = (unsigned)(k->order >> (4 * (15 - i))) & 15;
                    size_t at = ((y0 + (n >> 2)) * k->w + x0 + (n & 3)) * k->bands;
                    uint64_t p = img[at + c];
                    if (cb != c) p = (p - img[at + cb]) & WM;
                    m[i] = mags_w(p - prv, bits);
                    prv = p;
                    bitsused |= m[i];
                }
                k->prev[c] = prv;
                unsigned rung = topbit(bitsused | 1), oldrung = (unsigned)k->runbits[c];
                k->runbits[c] = rung;
                if (!k->best || bitsused <= 1) {
                    put_cs(s, U, rung - oldrung, 1);
                    put_body(s, m, bitsused, k->use_step);
                    continue;
                }
                /* BEST: QB3encode.h:691-713 */
                uint64_t cf = gcf(m);
                obits e = {scratch_a, 0};
                if (cf >= 2) put_cf(&e, m, cf, k->pcf[c], oldrung, U, bits);
                else { put_cs(&e, U, rung - oldrung, 1); put_body(&e, m, bitsused, 1); }
                int use_idx = 0;
                obits x = {scratch_b, 0};
                if (rung > 3 && rung < 63 && e.pos >= 36 + 3 * U + 2 * rung)
                    use_idx = put_index(&x, m, rung, oldrung, U) < e.pos;
                if (use_idx) e = x;
                else if (cf > 1) k->pcf[c] = (cf - 2) & WM;
                for (size_t b = 0; b < e.pos; b += 8) {
                    unsigned n = e.pos - b < 8 ? (unsigned)(e.pos - b) : 8;
                    put(s, e.p[b >> 3], n);
                }
            }
        }
    }
}

/* ---------------------------------------------------------- group decode */

/* 16 values at a rung (QB3decode.h:142-290), optional step undo (QB3decode.h:285-289) */
static void get_body(ibits *s, unsigned r, uint64_t g[16], int use_step)
{
    if (r == 0) {
        if (get(s, 1)) for (int i = 0; i < 16; i++) g[i] = get(s, 1);
        else memset(g, 0, 16 * sizeof(uint64_t));
        return;
    }
    for (int i = 0; i < 16; i++) {
        unsigned len, hi = 0;
        uint64_t x = peek(s);
        if (r == 63) { ibits t = *s; advance(&t, 64); hi = (unsigned)(peek(&t) & 1); }
        g[i] = gdcode(x, hi, r, &len);
        advance(s, len);
    }
    if (use_step) {
        unsigned M = 0;
        for (int i = 0; i < 16; i++) M |= (unsigned)((g[i] >> r) & 1) << i;
        if ((M & (M + 1)) == 0 && M != 0xffff)
            g[__builtin_popcount(M)] ^= 1ull << r;
    }
}

/* Restates decodeFTL (QB3decode.h:293-412) and decode (QB3decode.h:578-741). Returns nonzero on failure. */
static int decode_stream(const uint8_t *src, size_t len, uint64_t *img, const stream_cfg *k, int ftl)
{
    const unsigned bits = k->bits, U = ubits_of(bits), MASK = (1u << U) - 1, LONG = 2 * MASK + 1;
    const uint64_t WM = wmask(bits);
    size_t nbx = (k->w + 3) / 4, nby = (k->h + 3) / 4;
    ibits s = {src, len * 8, 0};
    int failed = 0;
    for (size_t by = 0; by < nby && !failed; by++) {
        size_t y0 = by * 4 + 4 > k->h ? k->h - 4 : by * 4;
        for (size_t bx = 0; bx < nbx && !failed; bx++) {
            size_t x0 = bx * 4 + 4 > k->w ? k->w - 4 : bx * 4;
            for (size_t c = 0; c < k->bands; c++) {
                uint64_t g[16];
                unsigned cs = 0;
                if (get(&s, 1)) {
                    cs = qb3o_dsw(U, (unsigned)peek(&s) & LONG);
                    advance(&s, (cs >> 12) - 1);
                }
                if (ftl || (cs & 0xfff) != 0 || cs == 0) {
                    unsigned rung = (unsigned)(k->runbits[c] + cs) & MASK;
                    k->runbits[c] = rung;
                    get_body(&s, rung, g, !ftl);
                }
                else { /* SIGNAL: QB3decode.h:624-716 */
                    unsigned l;
                    cs = qb3o_dsw(U, (unsigned)peek(&s) & LONG);
                    unsigned rung = (unsigned)(k->runbits[c] + cs) & MASK;
                    advance(&s, (cs >> 12) - 1);
                    if (rung != MASK) { /* common factor */
                        unsigned cfrung = rung;
                        uint64_t cf = k->pcf[c];
                        if (get(&s, 1)) {
                            unsigned own = (unsigned)get(&s, 1);
                            if (own) {
                                cs = qb3o_dsw(U, (unsigned)peek(&s) & LONG);
                                cfrung = (rung + cs) & MASK;
                                failed |= cfrung == rung;
                                advance(&s, (cs >> 12) - 1);
                            }
                            if (own && cfrung == 0) { failed = 1; break; } /* reference indexes DRG[-1] here */
                            cf = sdcode(peek(&s), cfrung - own, &l) + ((uint64_t)own << cfrung);
                            cf &= WM;
                            k->pcf[c] = cf;
                            advance(&s, l);
                        }
                        cf = (cf + 2) & WM;
                        if (rung) {
                            uint64_t used = 0;
                            get_body(&s, rung, g, 1);
                            for (int i = 0; i < 16; i++) /* magsmul, QB3decode.h:575 */
                                used |= g[i] = (magsabs(g[i]) * (cf << 1) - (g[i] & 1)) & WM;
                            k->runbits[c] = topbit(used | 1);
                            failed |= cf > used;
                        }
                        else {
                            uint64_t v = (((cf - 1) << 1) | 1) & WM;
                            uint64_t b = get(&s, 16);
                            for (int i = 0; i < 16; i++) g[i] = ((b >> i) & 1) ? v : 0;
                            k->runbits[c] = topbit(v);
                        }
                    }
                    else { /* index group */
                        uint64_t tbl[8] = {0};
                        unsigned maxidx = 0, used = 0;
                        cs = qb3o_dsw(U, (unsigned)peek(&s) & LONG);
                        rung = (unsigned)(k->runbits[c] + cs) & MASK;
                        k->runbits[c] = rung;
                        failed |= rung == 63;
                        advance(&s, (cs >> 12) - 1);
                        for (int i = 0; i < 16; i++) {
                            g[i] = dcode(peek(&s), 0, 2, &l);
                            advance(&s, l);
                            used += l;
                            if (g[i] > maxidx) maxidx = (unsigned)g[i];
                        }
                        failed |= used > 52;
                        for (unsigned i = 0; i <= maxidx; i++) {
                            tbl[i] = sdcode(peek(&s), rung, &l);
                            advance(&s, l);
                        }
                        for (int i = 0; i < 16; i++) g[i] = tbl[g[i]];
                    }
                }
                uint64_t prv = k->prev[c];
                for (int i = 0; i < 16; i++) {
                    unsigned n = (unsigned)(k->order >> (4 * (15 - i))) & 15;
                    prv = (prv + smag_w(g[i], bits)) & WM;
                    img[((y0 + (n >> 2)) * k->w + x0 + (n & 3)) * k->bands + c] = prv;
                }
                k->prev[c] = prv;
            }
        }
        if (failed) break;
        /* band delta undo per block row, QB3decode.h:730-737 */
        for (size_t j = 0; j < 4; j++)
            for (size_t c = 0; c < k->bands; c++)
                if (k->cband[c] != c)
                    for (size_t x = 0; x < k->w; x++) {
                        size_t at = ((y0 + j) * k->w + x) * k->bands;
                        img[at + c] = (img[at + c] + img[at + k->cband[c]]) & WM;
                    }
    }
    return failed || (s.len - s.pos) > 7;
}

/* -------------------------------------------------------- container: encode */

static uint64_t load_val(const void *p, size_t i, int type)
{
    switch (type >> 1) {
    case 0: return ((const uint8_t *)p)[i];
    case 1: return ((const uint16_t *)p)[i];
    case 2: return ((const uint32_t *)p)[i];
    default: return ((const uint64_t *)p)[i];
    }
}
static void store_val(void *p, size_t i, int type, uint64_t v)
{
    switch (type >> 1) {
    case 0: ((uint8_t *)p)[i] = (uint8_t)v; break;
    case 1: ((uint16_t *)p)[i] = (uint16_t)v; break;
    case 2: ((uint32_t *)p)[i] = (uint32_t)v; break;
    default: ((uint64_t *)p)[i] = v;
    }
}
static int64_t sext(uint64_t v, unsigned bits) { return bits >= 64 ? (int64_t)v : (int64_t)(v << (64 - bits)) >> (64 - bits); }

/* quantize(), QB3encode.cpp:137-186: C truncating / and % in the (signed or unsigned) type */
static uint64_t quantize_val(uint64_t v, uint64_t q, int away, int is_signed, unsigned bits)
{
    if (is_signed) {
        int64_t n = sext(v, bits), d = (int64_t)q, r;
        if (q == 2) r = away ? n / 2 + n % 2 : n / 2;
        else if (q == 3) r = n / 3 + (n % 3) / 2;
        else if (q == 4) r = away ? n / 4 + (n % 4) / 2 : n / 4 + (n % 4) / 3;
        else if (away) {
            int64_t m = n % d, h = d / 2 + d % 2;
            r = n / d + (!(n < 0) & (m >= h)) - ((n < 0) & ((m + h) <= 0));
        }
        else {
            int64_t m = n % d, h = d / 2;
            r = n / d + (!(n < 0) & (m > h)) - ((n < 0) & ((m + h) < 0));
        }
        return (uint64_t)r & wmask(bits);
    }
    uint64_t n = v, d = q, r;
    if (q == 2) r = away ? n / 2 + n % 2 : n / 2;
    else if (q == 3) r = n / 3 + (n % 3) / 2;
    else if (q == 4) r = away ? n / 4 + (n % 4) / 2 : n / 4 + (n % 4) / 3;
    else if (away) { uint64_t m = n % d, h = d / 2 + d % 2; r = n / d + (m >= h); }
    else { uint64_t m = n % d, h = d / 2; r = n / d + (m > h); }
    return r & wmask(bits);
}

static int has_banddiff(const qb3o_enc *e)
{
    for (size_t c = 0; c < e->nbands; c++) if (e->cband[c] != c) return 1;
    return 0;
}

/* write_headers, QB3encode.cpp:189-268 */
static void put_headers(const qb3o_enc *e, int mode, obits *s)
{
    put(s, 0x80334251u, 32); /* "QB3\200" */
    put(s, e->xsize - 1, 16);
    put(s, e->ysize - 1, 16);
    put(s, e->nbands - 1, 8);
    put(s, (unsigned)e->type, 8);
    put(s, (unsigned)mode & 0xff, 8);
    if (mode != MODE_STORED && has_banddiff(e)) {
        put(s, 'C' | ('B' << 8), 16);
        put(s, e->nbands, 16);
        for (size_t c = 0; c < e->nbands; c++) put(s, e->cband[c], 8);
    }
    if (e->quanta >= 2) {
        unsigned qbytes = 1 + topbit(e->quanta) / 8;
        put(s, 'Q' | ('V' << 8), 16);
        put(s, qbytes, 16);
        put(s, e->quanta, qbytes * 8);
    }
    if (e->order != ZCURVE && mode != MODE_STORED) {
        put(s, 'S' | ('C' << 8), 16);
        put(s, 8, 16);
        put(s, e->order ? e->order : HILBERT, 64);
    }
    put(s, 'D' | ('T' << 8), 16);
}

/* RLE0 / RLE0Size, QB3encode.cpp:271-332. dst == NULL only counts. */
static size_t rle0(const uint8_t *src, size_t len, uint8_t *dst)
{
    size_t i = 0, n = 0;
    uint8_t last = 0;
#define OUT(b) do { uint8_t b_ = (uint8_t)(b); if (dst) dst[n] = b_; n++; } while (0)
    while (i + 2 < len) {
        uint8_t c = src[i++];
        size_t left = len - i; /* bytes after c */
        int pair = (c == 0 || c == 0xff) && c == src[i];
        if (!pair || (c == 0 && (last == 0xff || left < 3 || src[i + 1] || src[i + 2]))) {
            OUT(c); last = c;
            continue;
        }
        i++;
        if (c == 0) {
            size_t run = 0;
            i += 2;
            while (run < 0xfe && i + run < len && src[i + run] == 0) run++;
            i += run;
            c = (uint8_t)run;
        }
        last = 0;
        OUT(0xff); OUT(0xff); OUT(c);
    }
    while (i < len) OUT(src[i++]);
#undef OUT
    return n;
}

static size_t raw_size(const qb3o_enc *e) { return e->xsize * e->ysize * e->nbands * TYPESIZE[e->type]; }

/* stored_encode, QB3encode.cpp:461-485 (stride in values) */
static size_t stored_encode(qb3o_enc *e, const void *src, uint8_t *dst)
{
    obits s = {dst, 0};
    /* the reference also overwrites p->mode here for good (QB3encode.cpp:464); not replicated */
    put_headers(e, MODE_STORED, &s);
    size_t hdr = s.pos / 8, ts = TYPESIZE[e->type], line = e->xsize * e->nbands * ts;
    size_t stride = (e->stride ? e->stride : e->xsize * e->nbands) * ts;
    for (size_t y = 0; y < e->ysize; y++)
        memcpy(dst + hdr + y * line, (const uint8_t *)src + y * stride, line);
    return hdr + raw_size(e);
}

int qb3o_init(qb3o_enc *e, size_t w, size_t h, size_t bands, int type)
{
    if (w == 0 || w > 0x10000 || h == 0 || h > 0x10000 || bands == 0 || bands > QB3O_MAXBANDS || type < 0 || type > 7)
        return 1;
    memset(e, 0, sizeof(*e));
    e->xsize = w; e->ysize = h; e->nbands = bands; e->type = type;
    e->quanta = 1; e->mode = MODE_FTL;
    for (size_t c = 0; c < bands; c++) e->cband[c] = (uint8_t)c;
    if (bands == 3 || bands == 4) e->cband[0] = e->cband[2] = 1;
    return 0;
}

void qb3o_reset(qb3o_enc *e)
{
    memset(e->prev, 0, sizeof(e->prev));
    memset(e->runbits, 0, sizeof(e->runbits));
    memset(e->cf, 0, sizeof(e->cf));
    e->error = 0;
}

int qb3o_set_mode(qb3o_enc *e, int mode)
{
    if (mode >= 0 && mode < 9) e->mode = mode;
    if (e->mode >= 0 && e->mode <= 3) e->order = ZCURVE;
    return e->mode;
}

int qb3o_set_coreband(qb3o_enc *e, size_t bands, size_t *cband)
{
    if (bands != e->nbands) return 0;
    for (size_t i = 0; i < bands; i++) e->cband[i] = (uint8_t)(cband[i] < bands ? cband[i] : i);
    for (size_t i = 0; i < bands; i++) if (e->cband[i] != i) e->cband[e->cband[i]] = e->cband[i];
    for (size_t i = 0; i < bands; i++) cband[i] = e->cband[i];
    return 1;
}

size_t qb3o_max_encoded_size(const qb3o_enc *e) /* QB3encode.cpp:112-118 */
{
    size_t n = 16 * ((e->xsize + 3) / 4) * ((e->ysize + 3) / 4) * e->nbands;
    double bits_per_value = 17.0 / 16.0 + 8 * TYPESIZE[e->type];
    return 1024 + (size_t)(bits_per_value * n / 8);
}

size_t qb3o_encode(qb3o_enc *e, const void *src, void *destination)
{
    uint8_t *dst = (uint8_t *)destination;
    if (e->xsize * e->ysize <= 16) return stored_encode(e, src, dst);
    const int user_mode = e->mode;
    const int rle = user_mode == 2 || user_mode == 3 || user_mode == 6 || user_mode == 7;
    const int mode = rle ? user_mode - 2 : user_mode; /* QB3encode.cpp:494-506 */
    if (mode < 0 || mode > 8) { e->error = 1; return 0; }
    const unsigned bits = 8 * TYPESIZE[e->type];
    const size_t w = e->xsize, h = e->ysize, nb = e->nbands;
    for (size_t c = 0; c < nb; c++) if (e->cband[c] >= nb) { e->error = 2; return 0; }

    obits s = {dst, 0};
    put_headers(e, mode, &s);
    const size_t hdr = s.pos / 8;

    /* virtual image: the image itself, or the small-image reorder of QB3encode.cpp:351-389 */
    size_t vw = w, vh = h;
    if (w < 4 || h < 4) {
        size_t ng = (w * h + 15) / 16;
        if (w < 4) { vw = 4; vh = ng * 4; } else { vw = ng * 4; vh = 4; }
    }
    uint64_t *img = (uint64_t *)calloc(vw * vh * nb, sizeof(uint64_t));
    const size_t stride = e->stride ? e->stride : w * nb;
    for (size_t y = 0; y < h; y++)
        for (size_t x = 0; x < w; x++) {
            size_t pix = (w >= 4 && h >= 4) ? y * w + x : (w < 4 ? y * w + x : x * h + y);
            for (size_t c = 0; c < nb; c++) {
                uint64_t v = load_val(src, y * stride + x * nb + c, e->type);
                if (e->quanta >= 2) v = quantize_val(v, e->quanta, e->away, e->type & 1, bits);
                img[pix * nb + c] = v;
            }
        }
    /* Quantised images and images with a side under 4 are coded through a copy of the control structure
       (encs subimg(*p), QB3encode.cpp:405; smallimg, :352): the running state is read, never written back. */
    uint64_t (*st)[QB3O_MAXBANDS] = NULL;
    uint64_t *prev = e->prev, *runbits = e->runbits, *cf = e->cf;
    if (e->quanta >= 2 || w < 4 || h < 4) {
        st = malloc(3 * sizeof(*st));
        memcpy(st[0], e->prev, sizeof(*st)); memcpy(st[1], e->runbits, sizeof(*st)); memcpy(st[2], e->cf, sizeof(*st));
        prev = st[0]; runbits = st[1]; cf = st[2];
    }
    stream_cfg k = {vw, vh, nb, e->cband, e->order ? e->order : HILBERT, bits,
                    mode != MODE_FTL, mode == 1 || mode == 5, prev, runbits, cf};
    encode_stream(img, &k, &s);
    free(img);
    free(st);

    size_t len = (s.pos + 7) / 8;
    s.pos = len * 8;
    if (rle && len <= qb3o_max_encoded_size(e) / 2) { /* QB3encode.cpp:536-566 */
        size_t data = len - hdr, avail = qb3o_max_encoded_size(e) - len;
        size_t rsz = rle0(dst + hdr, data, NULL);
        if (rsz <= avail && rsz < data) {
            uint8_t *tmp = (uint8_t *)malloc(rsz);
            rle0(dst + hdr, data, tmp);
            obits r = {dst, 0};
            put_headers(e, user_mode, &r);
            memcpy(dst + r.pos / 8, tmp, rsz);
            free(tmp);
            return r.pos / 8 + rsz;
        }
    }
    if (raw_size(e) > len) return len;
    return stored_encode(e, src, dst);
}

/* -------------------------------------------------------- container: decode */

static int valid_curve(uint64_t v)
{
    unsigned m = 0;
    for (int i = 0; i < 16; i++, v >>= 4) m |= 1u << (v & 15);
    return m == 0xffff;
}

int qb3o_read_info(const void *source, size_t len, qb3o_info *o)
{
    const uint8_t *p = (const uint8_t *)source;
    memset(o, 0, sizeof(*o));
    if (len < 15 || p[0] != 'Q' || p[1] != 'B' || p[2] != '3' || p[3] != 0x80) return 1;
    o->xsize = 1 + (p[4] | (p[5] << 8));
    o->ysize = 1 + (p[6] | (p[7] << 8));
    o->nbands = 1 + p[8];
    o->type = p[9];
    o->mode = p[10];
    if ((o->mode >= 9 && o->mode != MODE_STORED) || ((p[11] | p[12]) & 0x80) || o->type > 7) return 1;
    if (o->mode <= 3) o->order = ZCURVE;
    ibits s = {p + 11, (len - 11) * 8, 0};
    if (len - 11 < 4) return 1;
    for (;;) { /* QB3decode.cpp:186-260 */
        uint64_t v = peek(&s);
        unsigned sig = (unsigned)(v & 0xffff), clen = (unsigned)((v >> 16) & 0xffff);
        if (sig == ('Q' | ('V' << 8))) {
            if (clen > 4 || clen < 1) return 1;
            advance(&s, 32);
            o->quanta = get(&s, clen * 8);
            if (o->quanta < 2) return 1;
        }
        else if (sig == ('C' | ('B' << 8))) {
            if (clen != o->nbands) return 1;
            advance(&s, 32);
            int bad = 0;
            for (size_t i = 0; i < o->nbands; i++) {
                o->cband[i] = (uint8_t)get(&s, 8);
                bad |= o->cband[i] >= o->nbands;
            }
            if (bad) return 1;
            o->has_cb = 1;
        }
        else if (sig == ('D' | ('T' << 8))) {
            advance(&s, 16);
            size_t used = s.pos / 8;
            if (len - 11 <= used) return 1;
            o->data_offset = 11 + used;
            return 0;
        }
        else if (sig == ('S' | ('C' << 8))) {
            if (clen != 8) return 1;
            if (o->mode < 4 || o->mode == MODE_STORED) return 1;
            advance(&s, 32);
            o->order = get(&s, 64);
            if (!valid_curve(o->order)) return 1;
        }
        else {
            if (sig & 0x20) advance(&s, (size_t)clen * 8);
            else return 2;
        }
        if (s.pos >= s.len) return 1;
    }
}

/* deRLE0Size / deRLE0, QB3decode.cpp:267-307 */
static size_t derle0_size(const uint8_t *src, size_t len)
{
    size_t i = 0, n = 0;
    while (i + 2 < len) {
        if (src[i] != 0xff || src[i + 1] != 0xff) { n++; i++; continue; }
        n += src[i + 2] == 0xff ? 2 : 4 + (size_t)src[i + 2];
        i += 3;
    }
    return n + (len - i);
}
static int derle0(const uint8_t *src, size_t len, uint8_t *d, size_t dlen)
{
    size_t i = 0, n = 0;
    while (n < dlen && i + 2 < len) {
        uint8_t c = src[i++];
        if (c != 0xff || src[i] != 0xff) { d[n++] = c; continue; }
        size_t count = 2;
        if (src[i + 1] != 0xff) { c = 0; count = 4 + (size_t)src[i + 1]; }
        if (dlen - n < count) return 1;
        i += 2;
        while (count--) d[n++] = c;
    }
    while (i < len && n < dlen) d[n++] = src[i++];
    return (dlen - n) != (len - i);
}

size_t qb3o_decode(const void *source, size_t len, void *dst, size_t stride_in, int identity_default)
{
    qb3o_info o;
    if (qb3o_read_info(source, len, &o)) return 0;
    const uint8_t *src = (const uint8_t *)source + o.data_offset;
    size_t n = len - o.data_offset;
    const size_t w = o.xsize, h = o.ysize, nb = o.nbands, ts = TYPESIZE[o.type];
    const size_t outsz = w * h * nb * ts;
    const size_t stride = stride_in ? stride_in : w * nb;
    const unsigned bits = 8 * (unsigned)ts;
    if (o.mode == MODE_STORED) { /* QB3decode.cpp:356-375 */
        if (n != outsz) return 0;
        for (size_t y = 0; y < h; y++)
            memcpy((uint8_t *)dst + y * stride * ts, src + y * w * nb * ts, w * nb * ts);
        return outsz;
    }
    if (w * h < 16) return 0;
    uint8_t *unrle = NULL;
    if (o.mode == 2 || o.mode == 3 || o.mode == 6 || o.mode == 7) {
        size_t sz = derle0_size(src, n);
        if (sz > outsz) return 0;
        unrle = (uint8_t *)malloc(sz ? sz : 1);
        if (derle0(src, n, unrle, sz)) { free(unrle); return 0; }
        src = unrle; n = sz;
    }
    if (!o.has_cb && identity_default)
        for (size_t c = 0; c < nb; c++) o.cband[c] = (uint8_t)c;
    size_t vw = w, vh = h;
    if (w < 4 || h < 4) { /* QB3decode.cpp:321-329 */
        size_t ng = (w * h + 15) / 16;
        if (w < 4) { vw = 4; vh = ng * 4; } else { vw = ng * 4; vh = 4; }
    }
    uint64_t *img = (uint64_t *)calloc(vw * vh * nb, sizeof(uint64_t));
    uint64_t prev[QB3O_MAXBANDS] = {0}, runbits[QB3O_MAXBANDS] = {0}, pcf[QB3O_MAXBANDS] = {0};
    stream_cfg k = {vw, vh, nb, o.cband, o.order ? o.order : HILBERT, bits, 0, 0, prev, runbits, pcf};
    int failed = decode_stream(src, n, img, &k, o.mode == MODE_FTL);
    free(unrle);
    if (failed) { free(img); return 0; }
    const int is_signed = o.type & 1;
    for (size_t y = 0; y < h; y++)
        for (size_t x = 0; x < w; x++) {
            size_t pix = (w >= 4 && h >= 4) ? y * w + x : (w < 4 ? y * w + x : x * h + y);
            for (size_t c = 0; c < nb; c++) {
                uint64_t v = img[pix * nb + c];
                if (o.quanta > 1) { /* dequantize, QB3decode.cpp:77-107 */
                    uint64_t q = o.quanta;
                    if (is_signed) {
                        int64_t d = sext(v, bits), mx = (int64_t)(wmask(bits) >> 1), mn = -mx - 1;
                        int64_t r = d <= mx / (int64_t)q ? (int64_t)((uint64_t)d * q) : mx;
                        if (q > 2 && d < mn / (int64_t)q) r = mn;
                        v = (uint64_t)r & wmask(bits);
                    }
                    else v = v <= wmask(bits) / q ? v * q : wmask(bits);
                }
                store_val(dst, y * stride + x * nb + c, o.type, v);
            }
        }
    free(img);
    return outsz;
}
