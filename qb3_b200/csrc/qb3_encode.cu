/*
 * qb3_encode.cu -- the QB3 encode kernel for sm_100a.
 *
 * Replaces the serial per-group loops QB3::encode_fast / QB3::encode_best of the reference
 * (QB3encode.h:376-451, 617-724) and the byte-at-a-time writer oBits (bitstream.h:66-126).
 *
 * One CTA encodes one tile, walking it in stream order a segment (a run of 4x4 blocks of one block
 * row, all bands) at a time, one thread per group (block, band):
 *   1. the four image rows of the segment are staged in shared memory with 16 byte loads
 *   2. every thread gathers its 16 values in curve order, subtracts the core band, takes the running
 *      delta and folds the sign. The two pieces of state the reference carries serially are neighbour
 *      lookups: the predictor of a group's first value is the previous block's last value of the same
 *      band, the previous rung is the previous block's rung -- no scan is needed for either
 *   3. the group's bit length is computed, a CTA wide exclusive scan turns lengths into bit offsets
 *   4. every thread packs its codes at its offset into a shared memory bit window. Whole 32 bit words
 *      are plain stores; the words two neighbouring threads share are merged with a segmented
 *      warp-shuffle OR, and only the two words a warp shares with its neighbours use atomicOr
 *   5. complete 16 byte units of the window go to global memory with coalesced vector stores, the
 *      partial unit is carried to the next segment
 * Input is read once, output written once. Headers, the stored fallback, the small-image reorder and
 * quantisation (QB3encode.cpp:151-268, 351-389, 461-485) run on the device too.
 */
#include <cstdlib>

#include "qb3_device.cuh"

namespace qb3 {

/* ------------------------------------------------------------------ staging */

/* quantize(), reference QB3encode.cpp:137-186: C++ truncating / and % in the signed or unsigned type */
template <int BITS> __device__ __forceinline__ uint64_t quantize_value(uint64_t v, uint64_t q, bool away, bool is_signed)
{
    const uint64_t M = lowmask64(BITS);
    if (is_signed) {
        const long long n = (long long)(v << (64 - BITS)) >> (64 - BITS), d = (long long)q;
        long long r;
        if (q == 2) r = away ? n / 2 + n % 2 : n / 2;
        else if (q == 3) r = n / 3 + (n % 3) / 2;
        else if (q == 4) r = away ? n / 4 + (n % 4) / 2 : n / 4 + (n % 4) / 3;
        else {
            const long long m = n % d, h = away ? d / 2 + d % 2 : d / 2;
            if (away) r = n / d + (long long)(n >= 0 && m >= h) - (long long)(n < 0 && m + h <= 0);
            else r = n / d + (long long)(n >= 0 && m > h) - (long long)(n < 0 && m + h < 0);
        }
        return (uint64_t)r & M;
    }
    const uint64_t n = v & M, d = q;
    uint64_t r;
    if (q == 2) r = away ? n / 2 + n % 2 : n / 2;
    else if (q == 3) r = n / 3 + (n % 3) / 2;
    else if (q == 4) r = away ? n / 4 + (n % 4) / 2 : n / 4 + (n % 4) / 3;
    else {
        const uint64_t m = n % d, h = away ? d / 2 + d % 2 : d / 2;
        r = n / d + (uint64_t)(away ? m >= h : m > h);
    }
    return r & M;
}

/*
 * Stage rows y0..y0+3, pixels xs..xs+npix-1 (all bands) of the coded geometry into shared memory.
 * Fast path: 16 byte copies at the source's own alignment (row r lands at stage + r * rowpitch + (addr & 15)).
 * General path: element by element through the small-image reorder (QB3encode.cpp:351-389) and the quantiser.
 */

/* Issues the copies; whole 16 byte units travel as cp.async (LDGSTS, L2 only), so the rows of the next segment
   are in flight while the current one is being coded. The caller commits and waits. */
template <typename T>
__device__ __forceinline__ void stage_rows(const EncArgs &a, const uint8_t *src, uint8_t *stage,
                                           uint32_t y0, uint32_t xs, uint32_t npix)
{
    constexpr int BITS = traits<T>::BITS;
    const uint32_t rowvals = npix * a.bands;
    if (a.vec_stage) {
        const uint32_t rowbytes = rowvals * (uint32_t)sizeof(T), upr = a.rowpitch >> 4;
        const uint64_t lpitch = a.stride * sizeof(T);
        const uint8_t *g0 = src + ((uint64_t)y0 * a.stride + (uint64_t)xs * a.bands) * sizeof(T);
        for (uint32_t idx = threadIdx.x; idx < 4 * upr; idx += blockDim.x) {
            const uint32_t r = (idx >= upr) + (idx >= 2 * upr) + (idx >= 3 * upr), u = idx - r * upr;
            const uint8_t *g = g0 + r * lpitch;
            const uint32_t mis = (uint32_t)((uintptr_t)g & 15), first = 16 * u; /* first: offset in the row's aligned span */
            if (first >= mis + rowbytes) continue;
            const uint8_t *ga = g - mis + first;
            uint8_t *sa = stage + r * a.rowpitch + first;
            if (first >= mis && first + 16 <= mis + rowbytes)
                cp_async16(sa, ga);
            else
                for (uint32_t b = 0; b < 16; b++)
                    if (first + b >= mis && first + b < mis + rowbytes) sa[b] = ga[b];
        }
        return;
    }
    const T *s = reinterpret_cast<const T *>(src);
    const uint64_t npixels = (uint64_t)a.w * a.h;
    for (uint32_t e = threadIdx.x; e < 4 * rowvals; e += blockDim.x) {
        const uint32_t r = e / rowvals, j = e - r * rowvals;
        const uint32_t px = xs + j / a.bands, c = j % a.bands, py = y0 + r;
        uint64_t x = px, y = py;
        bool inside = true;
        if (a.small == 1) { /* narrow: rows concatenated into a 4 wide image */
            const uint64_t p = (uint64_t)py * 4 + px;
            inside = p < npixels;
            y = p / a.w; x = p % a.w;
        }
        else if (a.small == 2) { /* short: column major pixels into a 4 high image */
            const uint64_t p = (uint64_t)py * a.vw + px;
            inside = p < npixels;
            x = p / a.h; y = p % a.h;
        }
        uint64_t v = 0;
        if (inside) {
            v = s[y * a.stride + x * a.bands + c];
            if (a.quanta > 1) v = quantize_value<BITS>(v, a.quanta, a.away != 0, a.is_signed != 0);
        }
        reinterpret_cast<T *>(stage + r * a.rowpitch)[j] = (T)v;
    }
}

/*
 * Bulk copy engine (TMA, one dimension): one thread moves a whole row global -> shared with a single instruction
 * (cp.async.bulk, SASS UBLKCP) and an mbarrier counts the bytes in. Source, destination and size are multiples of 16.
 */
__device__ __forceinline__ void bulk_row(uint32_t smem_dst, const void *gmem_src, uint32_t bytes, uint32_t mbar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 :: "r"(smem_dst), "l"(gmem_src), "r"(bytes), "r"(mbar) : "memory");
}
__device__ __forceinline__ void mbar_init(uint32_t mbar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(mbar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect(uint32_t mbar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(mbar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t mbar, uint32_t parity)
{
    uint32_t ok;
    do {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok) : "r"(mbar), "r"(parity) : "memory");
    } while (!ok);
}

/* ------------------------------------------------------------------ bit packing */

/* Per thread writer into the shared 32 bit word window. The thread owns bits [s, e). */
struct Packer {
    uint32_t *win;
    uint64_t acc;
    uint32_t n;      /* valid bits in acc, < 32 between calls */
    uint32_t w;      /* index of the word acc starts at */
    uint32_t head;   /* first completed word, kept back: it may be shared with earlier threads */
    bool crossed;    /* at least one word completed */

    __device__ __forceinline__ void start(uint32_t *window, uint32_t s)
    {
        win = window; acc = 0; n = s & 31; w = s >> 5; head = 0; crossed = false;
    }
    __device__ __forceinline__ void put32(uint32_t bits, uint32_t len) /* len <= 32, bits < 2^len */
    {
        acc |= (uint64_t)bits << n;
        n += len;
        if (n >= 32) {
            if (crossed) win[w] = (uint32_t)acc;
            else { head = (uint32_t)acc; crossed = true; }
            w++;
            acc >>= 32;
            n -= 32;
        }
    }
    __device__ __forceinline__ void put64(uint64_t bits, uint32_t len) /* len <= 64 */
    {
        if (len > 32) { put32((uint32_t)bits, 32); put32((uint32_t)(bits >> 32), len - 32); }
        else put32((uint32_t)bits, len);
    }
    /*
     * Merge the partial words of the warp's threads and store them. s / e are this thread's bit range
     * (threads without work pass s == e == end of the previous thread). Must be called by all 32 lanes.
     */
    __device__ __forceinline__ void finish(uint32_t s, uint32_t e)
    {
        const uint32_t lane = lane_id();
        uint32_t v = (uint32_t)acc; /* bits of the last, incomplete word; 0 when n == 0 */
        uint32_t f = crossed ? 1u : 0u;
        /* segmented inclusive OR scan: a thread that completed a word starts a new segment */
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t vo = __shfl_up_sync(0xffffffffu, v, d), fo = __shfl_up_sync(0xffffffffu, f, d);
            if (lane >= d) {
                if (!f) v |= vo;
                f |= fo;
            }
        }
        uint32_t cin = __shfl_up_sync(0xffffffffu, v, 1);
        if (lane == 0) cin = 0;
        const uint32_t w0 = __shfl_sync(0xffffffffu, s >> 5, 0);
        if (crossed) {
            const uint32_t wi = s >> 5;
            if (wi == w0) atomicOr(&win[wi], head | cin); /* may hold bits of the previous warp / segment */
            else win[wi] = head | cin;
        }
        if (lane == 31 && (e & 31) != 0 && v != 0) atomicOr(&win[e >> 5], v);
    }
};

/* ------------------------------------------------------------------ group coding */

/* Rung switch with its change flag (reference: QB3encode.h:439-440) */
template <int U> __device__ __forceinline__ uint32_t switch_entry(uint32_t rung, uint32_t oldrung)
{
    return cs_entry(U, (rung - oldrung) & ((1u << U) - 1));
}

/* In place: step-down flip (BASE) and the middle swap, so that code_len / code_bits apply directly.
   (reference: QB3encode.h:169-197 and the crg3..7 tables) */
template <typename W> __device__ __forceinline__ void prepare_group(W (&m)[16], uint32_t rung, bool use_step)
{
    if (use_step) {
        uint32_t M = 0;
#pragma unroll
        for (int i = 0; i < 16; i++) M |= ((uint32_t)(m[i] >> rung) & 1u) << i;
        const int k = step_encode_index(M);
#pragma unroll
        for (int i = 0; i < 16; i++) if (i == k) m[i] ^= (W)1 << rung;
    }
    if (group_swaps(rung)) {
#pragma unroll
        for (int i = 0; i < 16; i++) m[i] = mswap(m[i], rung);
    }
}


/* ------------------------------------------------------------------ BEST mode helpers */

/* a % b; for magnitudes below 2^24 (8 and 16 bit data) through the float unit, which beats the integer divide
   sequence by a factor of three: the quotient estimate is off by at most one, fixed up exactly */
template <typename W> __device__ __forceinline__ W small_mod(W a, W b) { return a % b; }
template <> __device__ __forceinline__ uint32_t small_mod<uint32_t>(uint32_t a, uint32_t b)
{
    if ((a | b) >> 24) return a % b;
    const uint32_t q = (uint32_t)__fdividef((float)a, (float)b);
    int32_t r = (int32_t)(a - q * b);
    if (r < 0) r += (int32_t)b;
    else if ((uint32_t)r >= b) r -= (int32_t)b;
    return (uint32_t)r;
}

/* gcd of the non-zero magnitudes of a group (reference: gcf, QB3encode.h:98-126). Starting from the smallest
   magnitude the remainders collapse at once, and nearly every group of real data is known to be 1 after a step or two. */
template <typename W> __device__ __forceinline__ W group_gcd(const W (&m)[16])
{
    W g = ~(W)0;
#pragma unroll
    for (int i = 0; i < 16; i++) {
        const W a = magsabs(m[i]);
        if (a != 0 && a < g) g = a;
    }
#pragma unroll
    for (int i = 0; i < 16; i++) {
        if (g > 1) {
            W a = small_mod<W>(magsabs(m[i]), g);
            while (a) { const W t = small_mod<W>(g, a); g = a; a = t; }
        }
    }
    return g;
}

/* value i of a group as it is coded: step-down flip of value k, then the middle swap */
template <typename W> __device__ __forceinline__ W coded_value(W v, int i, int k, uint32_t rung)
{
    if (i == k) v ^= (W)1 << rung;
    return group_swaps(rung) ? mswap(v, rung) : v;
}
template <typename W> __device__ __forceinline__ int step_index(const W (&m)[16], uint32_t rung)
{
    uint32_t M = 0;
#pragma unroll
    for (int i = 0; i < 16; i++) M |= ((uint32_t)(m[i] >> rung) & 1u) << i;
    return step_encode_index(M);
}
/* bits of the 16 values of a group at rung >= 1, with step coding */
template <typename W> __device__ __forceinline__ uint32_t body_len(const W (&m)[16], uint32_t rung)
{
    const int k = step_index(m, rung);
    uint32_t len = 0;
#pragma unroll
    for (int i = 0; i < 16; i++) len += code_len<W>(coded_value(m[i], i, k, rung), rung);
    return len;
}
/* stand-alone value: qb3csztbl (reference: QB3encode.h:144-150) */
template <typename W> __device__ __forceinline__ uint32_t single_len(W v, uint32_t rung)
{
    if (rung == 0) return 1;
    return code_len<W>(single_swaps(rung) ? mswap(v, rung) : v, rung);
}
/* length of a rung switch written without its change flag; "no change" is spelled as the signal */
__device__ __forceinline__ uint32_t cs_noflag(uint32_t U, uint32_t delta)
{
    uint32_t e = cs_entry(U, delta & ((1u << U) - 1));
    if ((e >> 12) == 1) e = cs_signal(U);
    return ((e >> 12) - 1) << 12 | ((e & 0xfff) >> 1);
}

template <typename W, int BITS> struct ValuePut {
    __device__ static __forceinline__ void put(Packer &pk, W v, uint32_t rung) /* rung >= 1 */
    {
        uint64_t lo; uint32_t hi;
        const uint32_t l = code_bits<W>(v, rung, lo, hi);
        if (BITS <= 16) pk.put32((uint32_t)lo, l);
        else if (BITS == 32 || l <= 64) pk.put64(lo, l);
        else { pk.put64(lo, 64); pk.put32(hi, 1); } /* 65 bits at rung 63, reference: QB3encode.h:267-275 */
    }
    __device__ static __forceinline__ void put_single(Packer &pk, W v, uint32_t rung)
    {
        if (rung == 0) { pk.put32((uint32_t)v & 1, 1); return; }
        put(pk, single_swaps(rung) ? mswap(v, rung) : v, rung);
    }
    __device__ static __forceinline__ void put_body(Packer &pk, const W (&m)[16], uint32_t rung)
    {
        const int k = step_index(m, rung);
#pragma unroll
        for (int i = 0; i < 16; i++) put(pk, coded_value(m[i], i, k, rung), rung);
    }
    __device__ static __forceinline__ void put_raw16(Packer &pk, const W (&m)[16]) /* 16 one bit values */
    {
        uint32_t b = 0;
#pragma unroll
        for (int i = 0; i < 16; i++) b |= ((uint32_t)m[i] & 1u) << i;
        pk.put32(b, 16);
    }
};

/* Index group analysis (reference: ienc, QB3encode.h:557-613): up to 8 distinct values in first-seen order,
   stable sorted by descending count; slot[i] = table index of value i. Returns false for more than 8. */
template <typename W> struct IndexTable {
    W val[8];
    uint32_t cnt[8];
    uint32_t n;
    uint64_t slots; /* 16 x 3 bits */
    __device__ bool build(const W (&m)[16])
    {
        /* cheap exact rejection first: values hashed into 64 buckets, more than 8 buckets hit means more than 8 values */
        uint64_t seen = 0;
#pragma unroll
        for (int i = 0; i < 16; i++) {
            const uint32_t h = (uint32_t)m[i] ^ (uint32_t)((uint64_t)m[i] >> 32);
            seen |= 1ull << ((h * 0x9E3779B1u) >> 26);
        }
        if (__popcll(seen) > 8) return false;
        /* Everything below indexes the table with compile time constants only (unrolled, predicated), so that it
           lives in registers: with run time indices it sat in local memory and the threads that came this far kept
           the rest of their CTA waiting at the next barrier (35 % of the stall samples of the BEST encoder). */
#pragma unroll
        for (int k = 0; k < 8; k++) { val[k] = 0; cnt[k] = 0; }
        n = 0;
        uint64_t first = 0; /* 16 x 3 bits: the first-seen index of value i */
#pragma unroll
        for (int i = 0; i < 16; i++) {
            uint32_t hit = 0;
#pragma unroll
            for (int k = 0; k < 8; k++) hit |= (uint32_t)(cnt[k] != 0 && val[k] == m[i]) << k;
            if (hit == 0) {
                if (n == 8) return false;
#pragma unroll
                for (int k = 0; k < 8; k++) if (k == (int)n) { val[k] = m[i]; cnt[k] = 1; }
                first |= (uint64_t)n << (3 * i);
                n++;
            }
            else {
#pragma unroll
                for (int k = 0; k < 8; k++) cnt[k] += (hit >> k) & 1;
                first |= (uint64_t)((uint32_t)__ffs((int)hit) - 1) << (3 * i);
            }
        }
        /* stable, by descending count: adjacent exchanges on strictly greater only (unused entries count 0 and stay
           behind); id[k] = first-seen index of the entry now at k */
        uint32_t id[8];
#pragma unroll
        for (int k = 0; k < 8; k++) id[k] = k;
#pragma unroll
        for (int pass = 0; pass < 7; pass++) {
#pragma unroll
            for (int j = 7; j > pass; j--) {
                if (cnt[j] > cnt[j - 1]) {
                    const W tv = val[j]; val[j] = val[j - 1]; val[j - 1] = tv;
                    const uint32_t tc = cnt[j]; cnt[j] = cnt[j - 1]; cnt[j - 1] = tc;
                    const uint32_t ti = id[j]; id[j] = id[j - 1]; id[j - 1] = ti;
                }
            }
        }
        uint32_t where = 0; /* 8 x 3 bits: the slot of the entry first seen as number f */
#pragma unroll
        for (int k = 0; k < 8; k++) where |= (uint32_t)k << (3 * id[k]);
        slots = 0;
#pragma unroll
        for (int i = 0; i < 16; i++) slots |= (uint64_t)((where >> (3 * ((uint32_t)(first >> (3 * i)) & 7))) & 7) << (3 * i);
        return true;
    }
    /* bits after the three prefix fields */
    __device__ uint32_t payload_len(uint32_t rung) const
    {
        uint32_t len = 0;
#pragma unroll
        for (int i = 0; i < 16; i++) len += code_len<uint32_t>((uint32_t)(slots >> (3 * i)) & 7, 2);
#pragma unroll
        for (int j = 0; j < 8; j++) if ((uint32_t)j < n) len += single_len<W>(val[j], rung);
        return len;
    }
};

/* mags of a difference held in the low BITS bits of d, whatever is above them (reference: QB3common.h:127-131). For
   8 and 16 bit data in 32 bit registers: sign extend, then (s << 1) ^ (s >> 31) needs no mask afterwards. */
template <int BITS, typename W> __device__ __forceinline__ W mags_of_delta(W d)
{
    if (BITS <= 16 && sizeof(W) == 4) {
        const int32_t s = BITS == 8 ? (int32_t)(int8_t)(uint32_t)d : (int32_t)(int16_t)(uint32_t)d;
        return (W)(uint32_t)((s << 1) ^ (s >> 31));
    }
    return mags<BITS, W>(d);
}

/* nibble i of the scan curve; a compile time constant for the two curves the encoder itself uses */
template <int CURVE> __device__ __forceinline__ uint32_t curve_pos(uint64_t order, int i)
{
    if (CURVE == 1) return (uint32_t)(HILBERT >> (4 * (15 - i))) & 15;
    if (CURVE == 2) return (uint32_t)(ZCURVE >> (4 * (15 - i))) & 15;
    return (uint32_t)(order >> (4 * (15 - i))) & 15;
}

/* rung >= 1 code of a value below 2^17 in 32 bit arithmetic, packed (len << 20) | bits; no middle swap */
__device__ __forceinline__ uint32_t packed_code32(uint32_t v, uint32_t r)
{
    const uint32_t top = v >> r, nxt = (v >> (r - 1)) & 1, tn = top | nxt;
    const uint32_t payload = v & (((1u << (r - 1)) << top) - 1);
    return ((r + tn + top) << 20) | (payload << (1 + tn)) | tn | (top << 1);
}

/* first entry of rung r (1..7) in the shared code table: 2^(r+1) entries per rung */
__device__ __forceinline__ uint32_t lut_base(uint32_t r) { return (2u << r) - 4; }
constexpr uint32_t LUT_ENTRIES = 508;

/* 8 / 16 bit data: the 16 values of a group as packed codes (len << 20) | bits, step flip first when asked for; the
   rung 1..7 codes (middle swap included) come from the shared table, higher rungs are computed */
__device__ __forceinline__ void group_codes(uint32_t (&m)[16], uint32_t rung, bool use_step, const uint32_t *lut)
{
    if (use_step) {
        const int k = step_index<uint32_t>(m, rung);
#pragma unroll
        for (int i = 0; i < 16; i++) if (i == k) m[i] ^= 1u << rung;
    }
    if (rung < 8) {
        const uint32_t *t = lut + lut_base(rung);
#pragma unroll
        for (int i = 0; i < 16; i++) m[i] = t[m[i]];
    }
    else {
#pragma unroll
        for (int i = 0; i < 16; i++) m[i] = packed_code32(m[i], rung);
    }
}

/* the packed codes of a group into the bit window: 8 bit data three codes to a word, 16 bit data two */
template <int BITS> __device__ __forceinline__ void put_codes(Packer &pk, const uint32_t (&c)[16])
{
    if (BITS == 8) {
#pragma unroll
        for (int j = 0; j < 6; j++) {
            const uint32_t a0 = c[3 * j], l0 = a0 >> 20;
            uint32_t cw = a0 & 0xfffffu, cl = l0;
            if (j < 5) {
                const uint32_t a1 = c[3 * j + 1], a2 = c[3 * j + 2], l01 = l0 + (a1 >> 20);
                cw |= ((a1 & 0xfffffu) << l0) | ((a2 & 0xfffffu) << l01);
                cl = l01 + (a2 >> 20);
            }
            pk.put32(cw, cl);
        }
    }
    else {
#pragma unroll
        for (int i = 0; i < 16; i += 2) {
            const uint32_t a0 = c[i], a1 = c[i + 1], l0 = a0 >> 20;
            pk.put64((uint64_t)(a0 & 0xfffffu) | ((uint64_t)(a1 & 0xfffffu) << l0), l0 + (a1 >> 20));
        }
    }
}

/* DENSE: built for CTAs of at most 384 threads, three to an SM (56 registers): the kernel is bound by instruction issue
   and a third CTA gives the schedulers more warps to pick from. Used for 8 bit FTL / BASE on the Hilbert curve. */
template <typename T, bool BEST, int CURVE, bool DENSE = false>
__global__ void __launch_bounds__(DENSE ? 384 : 512, DENSE ? (BEST ? 2 : 3) : BEST ? 1 : 2) encode_kernel(const __grid_constant__ EncArgs a)
{
    typedef typename traits<T>::W W;
    constexpr int BITS = traits<T>::BITS, U = traits<T>::U;
    constexpr uint32_t UMASK = (1u << U) - 1;
    (void)UMASK;

    extern __shared__ __align__(16) uint8_t smem[];
    uint32_t *win = reinterpret_cast<uint32_t *>(smem);
    uint8_t *stage = smem + (size_t)a.win_words * 4;                                                   /* [2][4][rowpitch] */
    unsigned long long *carry_prev = reinterpret_cast<unsigned long long *>(stage + 8 * (size_t)a.rowpitch); /* [2][bands] */
    uint32_t *scan_scratch = reinterpret_cast<uint32_t *>(carry_prev + 2 * a.bands);                   /* [33] */
    uint8_t *carry_rung = reinterpret_cast<uint8_t *>(scan_scratch + 36);                              /* [2][bands] */
    uint8_t *rung_s = carry_rung + 2 * a.bands;                                                        /* [blockDim] */
    /* BEST only: the common factor that was last written for each band is the one piece of cross-group state
       that is not a neighbour lookup; it is propagated with a strided max-scan over the committing groups */
    unsigned long long *cfm2_s = reinterpret_cast<unsigned long long *>(smem + a.best_off);            /* [blockDim] */
    unsigned long long *carry_pcf = cfm2_s + blockDim.x;                                               /* [2][bands] */
    int *commit_s = reinterpret_cast<int *>(carry_pcf + 2 * a.bands);                                  /* [blockDim] */
    /* BEST in parts (EncArgs::best_pass): the factors met while the incoming one is unknown, and whether it still is */
    unsigned long long *dep_mask = reinterpret_cast<unsigned long long *>(commit_s + blockDim.x + (blockDim.x & 1)); /* [bands] */
    uint8_t *carry_unk = reinterpret_cast<uint8_t *>(dep_mask + a.bands);                              /* [2][bands] */
    /* 8 / 16 bit FTL and BASE: the rung 1..7 group codes (middle swap included) and the rung switches as shared
       tables, generated here from the closed forms; (len << 20) | bits and (len << 12) | bits */
    constexpr bool USE_LUT = !BEST && BITS <= 16;
    constexpr bool HAVE_LUT = BITS <= 16; /* BEST codes its plain groups, the bulk, through the same tables */
    uint32_t *lut = reinterpret_cast<uint32_t *>(smem + a.lut_off);                                     /* [508] */
    uint16_t *cs_lut = reinterpret_cast<uint16_t *>(lut + LUT_ENTRIES);                                 /* [2^U] */

    const uint32_t tid = threadIdx.x, NT = blockDim.x;
    const uint32_t multi = a.parts > 1, tile = multi ? blockIdx.x / a.parts : blockIdx.x, part = multi ? blockIdx.x % a.parts : 0;
    const uint32_t by_lo = part * a.part_rows, by_hi = multi ? min(a.nby, by_lo + a.part_rows) : a.nby;
    const uint8_t *src = a.src + (uint64_t)tile * a.src_pitch;
    uint8_t *dst = multi ? a.tmp + ((uint64_t)tile * a.parts + part) * a.tmp_slot : a.dst + (uint64_t)tile * a.slot;
    const uint64_t room = multi ? a.tmp_slot : a.slot;
    const uint32_t hdr_len = part == 0 ? a.hdr_len : 0;
    const bool use_step = a.mode != M_FTL;

    if (BEST && a.best_pass == 2 && !a.part_redo[(uint64_t)tile * a.parts + part]) return; /* nothing depended on the factor coming in */
    /* running state in, zero unless the caller keeps it across calls (reference: QB3encode.h:391-394) */
    for (uint32_t c = tid; c < a.bands; c += NT) {
        const unsigned long long *st = a.state ? a.state + (uint64_t)tile * 3 * a.bands : nullptr;
        carry_prev[c] = st ? st[c] : 0ull;
        carry_rung[c] = st ? (uint8_t)st[a.bands + c] : (uint8_t)0;
        if (BEST) {
            carry_pcf[c] = st ? st[2 * a.bands + c] : 0ull;
            if (a.best_pass) {
                unsigned long long *pp = a.part_pcf + (((uint64_t)tile * a.parts + part) * a.bands + c) * 4;
                if (a.best_pass == 2) carry_pcf[c] = pp[3];
                if (a.best_pass == 1 && part == 0) pp[3] = carry_pcf[c];
                carry_unk[c] = a.best_pass == 1 && part > 0;
                dep_mask[c] = 0;
            }
        }
    }
    for (uint32_t i = tid; i < a.win_words; i += NT) win[i] = 0;
    if (HAVE_LUT) {
        for (uint32_t i = tid; i < LUT_ENTRIES; i += NT) {
            const uint32_t r = topbit32(i + 4) - 1, v = i - lut_base(r);
            lut[i] = packed_code32(mswap<uint32_t>(v, r), r);
        }
        for (uint32_t i = tid; i < (1u << U); i += NT) cs_lut[i] = (uint16_t)cs_entry(U, i);
    }
    __syncthreads();
    for (uint32_t i = tid; i < hdr_len; i += NT) reinterpret_cast<uint8_t *>(win)[i] = a.hdr[i];
    if (part > 0 && tid < a.bands) {
        /* A later part starts from the state the blocks before it leave behind, and both pieces of it are neighbour
           look-ups: the last value and the rung of the last block of the block row above (for its rung, the last value
           of the block before that one). Straight from global memory: 17 pixels per band. */
        const T *img = reinterpret_cast<const T *>(src);
        const W TMp = (W)lowmask64(BITS);
        const uint32_t cc = tid, cb = a.cband[cc];
        auto pixel = [&](uint32_t x, uint32_t y, uint32_t k) -> W {
            uint64_t v = img[(uint64_t)y * a.stride + (uint64_t)x * a.bands + k];
            if (a.quanta > 1) v = quantize_value<BITS>(v, a.quanta, a.away != 0, a.is_signed != 0);
            return (W)v;
        };
        auto value = [&](uint32_t x, uint32_t y) -> W {
            W v = pixel(x, y, cc);
            if (cb != cc) v -= pixel(x, y, cb);
            return v & TMp;
        };
        const uint32_t n15 = curve_pos<CURVE>(a.order, 15);
        const uint32_t by = by_lo - 1, bx = a.nbx - 1, x0 = min(4 * bx, a.vw - 4), y0 = 4 * by; /* not the last block row */
        W before = (W)carry_prev[cc] & TMp; /* last value of the block ahead of that one; the caller's state at the very start */
        if (bx > 0) before = value(min(4 * (bx - 1), a.vw - 4) + (n15 & 3), y0 + (n15 >> 2));
        else if (by > 0) before = value(x0 + (n15 & 3), 4 * (by - 1) + (n15 >> 2));
        W used = 0, last = before;
#pragma unroll
        for (int i = 0; i < 16; i++) {
            const uint32_t n = curve_pos<CURVE>(a.order, i);
            const W v = value(x0 + (n & 3), y0 + (n >> 2));
            used |= mags<BITS, W>((v - last) & TMp);
            last = v;
        }
        carry_prev[cc] = (unsigned long long)last;
        carry_rung[cc] = (uint8_t)topbit((W)(used | 1));
    }

    uint32_t wbits = hdr_len * 8;      /* bits waiting in the window */
    uint64_t flushed = 0;              /* 16 byte units already in global memory */
    bool overflow = false;             /* output would not fit the slot: the tile ends up stored */
    uint32_t it = 0;

    const uint32_t blk = tid / a.bands, c = tid - blk * a.bands; /* this thread's block in the segment and band */
    /* byte SIMD front end: pixel j of band k sits at byte j * bands + k of a block's row; selectors that pick the four
       bytes of this thread's band, and of its core band, out of the row's four words */
    uint32_t sel_own_lo = 0, sel_own_hi = 0, sel_own_m = 0, sel_core_lo = 0, sel_core_hi = 0, sel_core_m = 0;
    if (BITS == 8 && a.simd8) {
        const uint32_t cbk = a.cband[c < a.bands ? c : 0];
        for (uint32_t j = 0; j < 4; j++) {
            const uint32_t oo = j * a.bands + c, ok = j * a.bands + cbk;
            sel_own_lo |= (oo & 7) << (4 * j); sel_own_hi |= (oo & 7) << (4 * j); sel_own_m |= (oo < 8 ? j : 4 + j) << (4 * j);
            sel_core_lo |= (ok & 7) << (4 * j); sel_core_hi |= (ok & 7) << (4 * j); sel_core_m |= (ok < 8 ? j : 4 + j) << (4 * j);
        }
    }
    const uint32_t stage_bytes = 4 * a.rowpitch;
    /* segment (by, sg) -> staged rows; issued one segment ahead */
    /* Rows that are whole 16 byte units at 16 byte addresses (a.bulk_stage, decided by the host for the launch) travel
       by the bulk copy engine: thread 0 issues four copies per segment and an mbarrier per staging buffer counts the
       bytes in -- no per thread address arithmetic, no issue slots (it was a sixth of the kernel's instructions as
       16 byte cp.async). Anything else goes through stage_rows. */
    __shared__ __align__(8) unsigned long long stage_bar[2];
    const uint32_t bar_sa = (uint32_t)__cvta_generic_to_shared(stage_bar);
    if (a.bulk_stage && tid == 0) {
        mbar_init(bar_sa, 1);
        mbar_init(bar_sa + 8, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    auto issue = [&](uint32_t by, uint32_t sg, uint32_t buf) {
        const uint32_t bx0 = sg * a.seg_blocks, nblk = min(a.seg_blocks, a.nbx - bx0);
        const uint32_t xs = min(4 * bx0, a.vw - 4), xe = min(4 * (bx0 + nblk), a.vw);
        if (a.bulk_stage) {
            if (tid == 0) {
                const uint32_t rowbytes = (xe - xs) * a.bands * (uint32_t)sizeof(T);
                const uint64_t lpitch = a.stride * sizeof(T);
                const uint8_t *g0 = src + ((uint64_t)min(4 * by, a.vh - 4) * a.stride + (uint64_t)xs * a.bands) * sizeof(T);
                const uint32_t d0 = (uint32_t)__cvta_generic_to_shared(stage + buf * stage_bytes);
                mbar_expect(bar_sa + 8 * buf, 4 * rowbytes);
#pragma unroll
                for (int r = 0; r < 4; r++) bulk_row(d0 + r * a.rowpitch, g0 + r * lpitch, rowbytes, bar_sa + 8 * buf);
            }
            return;
        }
        stage_rows<T>(a, src, stage + buf * stage_bytes, min(4 * by, a.vh - 4), xs, xe - xs);
    };
    if (a.bulk_stage) __syncthreads(); /* the barriers exist before anybody waits on them */
    if (a.small != 3) issue(by_lo, 0, 0);
    cp_async_commit();

    /* a.small == 3: sixteen pixels or fewer are stored outright (reference: QB3encode.cpp:490-491) */
    for (uint32_t by = by_lo; by < by_hi && a.small != 3; by++) {
        const uint32_t y0 = min(4 * by, a.vh - 4);
        for (uint32_t sg = 0; sg < a.segs; sg++, it++) {
            const uint32_t bx0 = sg * a.seg_blocks, nblk = min(a.seg_blocks, a.nbx - bx0), ng = nblk * a.bands;
            const uint32_t xs = min(4 * bx0, a.vw - 4);
            const uint32_t par = it & 1;
            {   /* next segment's rows start travelling now; this segment's have had a whole iteration to land */
                const uint32_t nsg = sg + 1 < a.segs ? sg + 1 : 0, nby = nsg ? by : by + 1;
                if (nby < by_hi) issue(nby, nsg, par ^ 1);
                cp_async_commit();
                cp_async_wait<1>();
                if (a.bulk_stage) mbar_wait(bar_sa + 8 * par, (it >> 1) & 1);
            }
            __syncthreads(); /* also orders the header / carry / table writes before their first use */
            const uint8_t *sbuf = stage + par * stage_bytes;

            const bool active = tid < ng;
            W m[16];
            uint32_t mw[6], ml[6]; /* 8 bit data, table path: the sixteen codes merged three to a word, and the words' lengths */
            W bitsused = 0;
            uint32_t rung = 0;
            if (active) {
                const uint32_t bx = bx0 + blk, x0 = min(4 * bx, a.vw - 4), cb = a.cband[c];
                /* where row r of the segment starts in the staging buffer: rows keep their source alignment, so the
                   offset moves with the row's address modulo 16. One 64 bit address is worked out, the rest follows. */
                uint32_t rowoff[4];
                {
                    uint32_t mis = 0, step = 0;
                    if (a.vec_stage) {
                        mis = (uint32_t)((uintptr_t)(src + ((uint64_t)y0 * a.stride + (uint64_t)xs * a.bands) * sizeof(T)) & 15);
                        step = (uint32_t)(a.stride * sizeof(T)) & 15;
                    }
#pragma unroll
                    for (int r = 0; r < 4; r++) rowoff[r] = r * a.rowpitch + ((mis + r * step) & 15);
                }
                const W TM = (W)lowmask64(BITS);
                /* the core band is always read and masked away for a band that is its own core: the lanes of a warp
                   hold different bands, and a branch here would run the whole gather twice */
                const W dmask = cb != c ? TM : (W)0;
                /* per row: this block's first pixel, own band; the core band is a fixed distance away */
                const T *own[4];
                const int core_d = (int)cb - (int)c;
                uint32_t coloff[4];
#pragma unroll
                for (int r = 0; r < 4; r++) {
                    own[r] = reinterpret_cast<const T *>(sbuf + rowoff[r]) + (x0 - xs) * a.bands + c;
                    coloff[r] = r * a.bands;
                }
                W prv;
                if (blk > 0) { /* last value of the previous block: curve position 15 */
                    const uint32_t n15 = curve_pos<CURVE>(a.order, 15);
                    const int back = (int)((4 * (bx - 1) + (n15 & 3)) - x0) * (int)a.bands;
                    const T *q15 = own[n15 >> 2] + back;
                    prv = (W)q15[0] - ((W)q15[core_d] & dmask); /* only the low BITS bits matter from here on */
                }
                else prv = (W)carry_prev[par * a.bands + c] & TM;
                if constexpr (BITS == 8 && CURVE != 0) {
                  if (a.simd8) {
                    /*
                     * Byte SIMD front end (8 bit data, up to four bands, rows staged at 16 byte addresses, width a
                     * multiple of four): a block's row is at most four words; the four pixels of this band are picked
                     * out of them with three byte permutes (the selectors depend on band and band count only), the
                     * core band likewise, and from there on four values travel per register: core band subtraction,
                     * the curve order (one permute per four positions on the Hilbert and Z curves), the running delta
                     * (the predecessor vector is the same words moved by a byte) and the sign folding.
                     */
                    uint32_t d[4];
#pragma unroll
                    for (int r = 0; r < 4; r++) {
                        const uint32_t *wp = reinterpret_cast<const uint32_t *>(sbuf + rowoff[r] + (x0 - xs) * a.bands);
                        const uint32_t w0 = wp[0], w1 = wp[1], w2 = wp[2], w3 = wp[3];
                        const uint32_t o = __byte_perm(__byte_perm(w0, w1, sel_own_lo), __byte_perm(w2, w3, sel_own_hi), sel_own_m);
                        const uint32_t k = __byte_perm(__byte_perm(w0, w1, sel_core_lo), __byte_perm(w2, w3, sel_core_hi), sel_core_m);
                        d[r] = __vsub4(o, k & (uint32_t)(0 - (uint32_t)(cb != c)));
                    }
                    uint32_t cv[4];
                    if (CURVE == 1) { /* Hilbert 0x01548cd9aefb7623: positions (x, y) by fours */
                        cv[0] = __byte_perm(d[0], d[1], 0x4510); cv[1] = __byte_perm(d[2], d[3], 0x1540);
                        cv[2] = __byte_perm(d[2], d[3], 0x3762); cv[3] = __byte_perm(d[0], d[1], 0x3267);
                    }
                    else {            /* Z 0x0145236789cdabef */
                        cv[0] = __byte_perm(d[0], d[1], 0x5410); cv[1] = __byte_perm(d[0], d[1], 0x7632);
                        cv[2] = __byte_perm(d[2], d[3], 0x5410); cv[3] = __byte_perm(d[2], d[3], 0x7632);
                    }
                    uint32_t before = (uint32_t)prv << 24, mm[4], used4 = 0;
#pragma unroll
                    for (int k = 0; k < 4; k++) {
                        const uint32_t pv = __byte_perm(before, cv[k], 0x6543); /* the value ahead of each of the four */
                        const uint32_t dl = __vsub4(cv[k], pv);
                        /* mags per byte (QB3common.h:127-131): (d << 1) ^ (0xff where d is negative) */
                        mm[k] = ((dl & 0x7f7f7f7fu) << 1) ^ (((dl >> 7) & 0x01010101u) * 0xffu);
                        used4 |= mm[k];
                        before = cv[k];
                    }
                    prv = cv[3] >> 24;
                    used4 |= used4 >> 16;
                    bitsused = (used4 | (used4 >> 8)) & 0xff;
#pragma unroll
                    for (int i = 0; i < 16; i++) m[i] = __byte_perm(mm[i >> 2], 0u, 0x4440 + (i & 3));
                  }
                }
                if (!(BITS == 8 && CURVE != 0 && a.simd8)) {
#pragma unroll
                for (int i = 0; i < 16; i++) {
                    const uint32_t n = curve_pos<CURVE>(a.order, i);
                    const T *q = own[n >> 2] + coloff[n & 3];
                    const W v = (W)q[0] - ((W)q[core_d] & dmask);
                    m[i] = mags_of_delta<BITS, W>(v - prv);
                    prv = v;
                    bitsused |= m[i];
                }
                }
                rung = topbit((W)(bitsused | 1));
                rung_s[tid] = (uint8_t)rung;
                if (blk == nblk - 1) { /* becomes the neighbour of the next segment's first block */
                    carry_prev[(par ^ 1) * a.bands + c] = (unsigned long long)prv;
                    carry_rung[(par ^ 1) * a.bands + c] = (uint8_t)rung;
                }
            }
            __syncthreads();

            uint32_t len = 0, cs = 0, oldrung = 0;
            /* BEST: 0 plain, 1 common factor, 2 index, 3 nothing (the reference's empty index group, see below) */
            uint32_t kind = 0, trung = 0, cfrung = 0;
            bool same = false;
            W q[16]; /* common factor quotients, BEST only */
            W cf = 1, cm2 = 0;
            if (active) {
                oldrung = blk > 0 ? rung_s[tid - a.bands] : carry_rung[par * a.bands + c];
                cs = USE_LUT ? cs_lut[(rung - oldrung) & UMASK] : switch_entry<U>(rung, oldrung);
            }
            if (!BEST) {
                if (active) {
                    len = cs >> 12;
                    if (bitsused <= 1) len += 1 + (bitsused ? 16 : 0); /* reference: QB3encode.h:159-166 */
                    else if (USE_LUT) {
                        /* every value becomes its packed code, (len << 20) | bits: the table holds the swapped rung 1..7
                           codes, higher rungs (16 bit data) are computed */
                        if (use_step) {
                            const int k = step_index<W>(m, rung);
#pragma unroll
                            for (int i = 0; i < 16; i++) if (i == k) m[i] ^= (W)1 << rung;
                        }
                        if (rung < 8) {
                            const uint32_t *t = lut + lut_base(rung);
#pragma unroll
                            for (int i = 0; i < 16; i++) m[i] = (W)t[(uint32_t)m[i]];
                        }
                        else {
#pragma unroll
                            for (int i = 0; i < 16; i++) m[i] = (W)packed_code32((uint32_t)m[i], rung);
                        }
                        if (BITS == 8) {
                            /* codes are 9 bits at most: three of them travel as one word of up to 27 bits. m[0..5] become
                               the merged words, m[8..13] their lengths */
#pragma unroll
                            for (int j = 0; j < 6; j++) {
                                const uint32_t a0 = (uint32_t)m[3 * j], l0 = a0 >> 20;
                                uint32_t cw = a0 & 0xfffffu, cl = l0;
                                if (j < 5) {
                                    const uint32_t a1 = (uint32_t)m[3 * j + 1], a2 = (uint32_t)m[3 * j + 2], l01 = l0 + (a1 >> 20);
                                    cw |= ((a1 & 0xfffffu) << l0) | ((a2 & 0xfffffu) << l01);
                                    cl = l01 + (a2 >> 20);
                                }
                                mw[j] = cw; ml[j] = cl;
                                len += cl;
                            }
                        }
                        else {
#pragma unroll
                            for (int i = 0; i < 16; i++) len += (uint32_t)m[i] >> 20;
                        }
                    }
                    else {
                        prepare_group<W>(m, rung, use_step);
#pragma unroll
                        for (int i = 0; i < 16; i++) len += code_len<W>(m[i], rung);
                    }
                }
            }
            else {
                /* reference: encode_best, QB3encode.h:691-713 */
                const W TM = (W)lowmask64(BITS);
                uint32_t l_plain = 0, l_same = 0, l_diff = 0, l_idx = 800;
                bool idx_elig = false, commit = false;
                const uint32_t idx_min = 36 + 3 * U;
                if (active) {
                    len = cs >> 12;
                    if (bitsused <= 1) len += 1 + (bitsused ? 16 : 0);
                    else {
                        cf = group_gcd<W>(m);
                        if (cf >= 2) { /* reference: cfgenc, QB3encode.h:283-361 */
                            W qbits = 0;
#pragma unroll
                            for (int i = 0; i < 16; i++) qbits |= q[i] = (((magsabs(m[i]) / cf) << 1) - (m[i] & 1)) & TM;
                            cm2 = (cf - 2) & TM;
                            trung = topbit((W)(qbits | 1));
                            cfrung = topbit((W)(cm2 | 1));
                            const uint32_t body = trung == 0 ? 16 : body_len<W>(q, trung);
                            l_same = (U + 2) + (cs_noflag(U, trung - oldrung) >> 12) + 1 + body;
                            uint32_t cfl;
                            if (trung >= cfrung && (trung < cfrung + U || cfrung == 0))
                                cfl = 1 + (trung == 0 ? 1 : single_len<W>(cm2, trung));
                            else
                                cfl = (cs_entry(U, (cfrung - trung) & UMASK) >> 12) + single_len<W>(cm2 ^ ((W)1 << cfrung), cfrung - 1);
                            l_diff = l_same + cfl;
                        }
                        else if constexpr (HAVE_LUT) { /* lengths only: the values are still needed for the index candidate */
                            uint32_t cc[16];
#pragma unroll
                            for (int i = 0; i < 16; i++) cc[i] = (uint32_t)m[i];
                            group_codes(cc, rung, true, lut);
                            l_plain = cs >> 12;
#pragma unroll
                            for (int i = 0; i < 16; i++) l_plain += cc[i] >> 20;
                        }
                        else l_plain = (cs >> 12) + body_len<W>(m, rung);
                        idx_elig = rung > 3 && rung < 63;
                        if (idx_elig) {
                            IndexTable<W> tb;
                            if (tb.build(m))
                                l_idx = (U + 2) + (cs_noflag(U, UMASK - oldrung) >> 12) + (cs_noflag(U, rung - oldrung) >> 12)
                                      + tb.payload_len(rung);
                        }
                        const uint32_t sz = cf >= 2 ? l_diff : l_plain;
                        commit = cf >= 2 && !(idx_elig && sz >= idx_min + 2 * rung && l_idx < sz);
                    }
                }
                /* previous common factor of this band: the latest earlier group of the band that committed one.
                   Inclusive max-scan of the committing thread index with stride 'bands', in shared memory. */
                int last = (active && commit) ? (int)tid : -1;
                cfm2_s[tid] = (unsigned long long)cm2;
                if (__syncthreads_or(last >= 0)) { /* nobody commits a factor in most segments of real data */
                    for (uint32_t d = a.bands; d < ng; d <<= 1) {
                        commit_s[tid] = last;
                        __syncthreads();
                        if (tid >= d && tid < ng) last = max(last, commit_s[tid - d]);
                        __syncthreads();
                    }
                }
                commit_s[tid] = last;
                __syncthreads();
                if (active) {
                    const int before = blk > 0 ? commit_s[tid - a.bands] : -1;
                    const W pcf = (W)(before >= 0 ? cfm2_s[before] : carry_pcf[par * a.bands + c]);
                    /* a later part of a tile, first pass: nobody has written a factor for this band yet in this part and
                       the one the parts before leave behind is not known */
                    const bool unk = a.best_pass && before < 0 && carry_unk[par * a.bands + c];
                    if (unk && cf >= 2) atomicOr(&dep_mask[c], 1ull << ((uint32_t)cm2 & 63));
                    if (blk == nblk - 1) {
                        carry_pcf[(par ^ 1) * a.bands + c] = last >= 0 ? cfm2_s[last] : carry_pcf[par * a.bands + c];
                        if (a.best_pass) carry_unk[(par ^ 1) * a.bands + c] = last >= 0 ? (uint8_t)0 : carry_unk[par * a.bands + c];
                    }
                    if (bitsused > 1) {
                        same = cf >= 2 && pcf == cm2 && !unk;
                        const uint32_t sz = cf >= 2 ? (same ? l_same : l_diff) : l_plain;
                        kind = cf >= 2 ? 1 : 0;
                        len = sz;
                        if (idx_elig && sz >= idx_min + 2 * rung && l_idx < sz) {
                            /* more than 8 distinct values report 800 bits; when the group is longer than that (64 bit
                               data only) the reference replaces it by the empty side buffer (QB3encode.h:704-708) */
                            kind = l_idx == 800 ? 3 : 2;
                            len = l_idx == 800 ? 0 : l_idx;
                        }
                    }
                }
            }
            uint32_t total;
            const uint32_t off = block_exclusive_scan(len, scan_scratch, total);

            const uint32_t s = wbits + off, e = s + len;
            Packer pk;
            pk.start(win, s);
            if (a.size_only) { /* the lengths are all that is wanted */
                const uint32_t Bq = wbits + total, nuq = Bq >> 7;
                if ((flushed + nuq) * 16 > room) overflow = true;
                flushed += nuq;
                wbits = Bq & 127;
                continue;
            }
            if (BEST && active && kind != 0) {
                typedef ValuePut<W, BITS> VP;
                const uint32_t sig = cs_signal(U);
                if (kind == 1) { /* common factor group */
                    uint32_t e = cs_noflag(U, trung - oldrung);
                    pk.put32(sig & 0xfff, sig >> 12);
                    pk.put32(e & 0xfff, e >> 12);
                    bool done = false;
                    if (!same) {
                        pk.put32(1, 1);
                        if (trung >= cfrung && (trung < cfrung + U || cfrung == 0)) {
                            pk.put32(0, 1);
                            if (trung == 0) { pk.put32((uint32_t)cm2 & 1, 1); VP::put_raw16(pk, q); done = true; }
                            else VP::put_single(pk, cm2, trung);
                        }
                        else {
                            e = cs_entry(U, (cfrung - trung) & UMASK);
                            pk.put32(e & 0xfff, e >> 12);
                            VP::put_single(pk, cm2 ^ ((W)1 << cfrung), cfrung - 1);
                            if (trung == 0) { VP::put_raw16(pk, q); done = true; }
                        }
                    }
                    else {
                        pk.put32(0, 1);
                        if (trung == 0) { VP::put_raw16(pk, q); done = true; }
                    }
                    if (!done) VP::put_body(pk, q, trung);
                }
                else if (kind == 2) { /* index group */
                    IndexTable<W> tb;
                    tb.build(m);
                    uint32_t e = cs_noflag(U, UMASK - oldrung);
                    pk.put32(sig & 0xfff, sig >> 12);
                    pk.put32(e & 0xfff, e >> 12);
                    e = cs_noflag(U, rung - oldrung);
                    pk.put32(e & 0xfff, e >> 12);
                    for (int i = 0; i < 16; i++) {
                        uint64_t lo; uint32_t hi;
                        const uint32_t l = code_bits<uint32_t>((uint32_t)(tb.slots >> (3 * i)) & 7, 2, lo, hi);
                        pk.put32((uint32_t)lo, l); /* no middle swap, reference: QB3encode.h:599-601 */
                    }
#pragma unroll
                    for (int j = 0; j < 8; j++) if ((uint32_t)j < tb.n) VP::put_single(pk, tb.val[j], rung);
                }
            }
            else if (BEST && active && bitsused > 1) { /* plain group, always with step coding */
                pk.put32(cs & 0xfff, cs >> 12);
                if constexpr (HAVE_LUT) {
                    uint32_t cc[16];
#pragma unroll
                    for (int i = 0; i < 16; i++) cc[i] = (uint32_t)m[i];
                    group_codes(cc, rung, true, lut);
                    put_codes<BITS>(pk, cc);
                }
                else ValuePut<W, BITS>::put_body(pk, m, rung);
            }
            else if (active) {
                pk.put32(cs & 0xfff, cs >> 12);
                if (bitsused <= 1) {
                    uint32_t b = (uint32_t)bitsused;
                    if (bitsused) {
#pragma unroll
                        for (int i = 0; i < 16; i++) b |= (uint32_t)m[i] << (i + 1);
                    }
                    pk.put32(b, bitsused ? 17 : 1);
                }
                else if (USE_LUT && BITS == 8) {
#pragma unroll
                    for (int j = 0; j < 6; j++) pk.put32(mw[j], ml[j]);
                }
                else if (USE_LUT) { /* 16 bit data: codes of up to 17 bits go in pairs */
#pragma unroll
                    for (int i = 0; i < 16; i += 2) {
                        const uint32_t a0 = (uint32_t)m[i], a1 = (uint32_t)m[i + 1], l0 = a0 >> 20;
                        pk.put64((uint64_t)(a0 & 0xfffffu) | ((uint64_t)(a1 & 0xfffffu) << l0), l0 + (a1 >> 20));
                    }
                }
                else {
#pragma unroll
                    for (int i = 0; i < 16; i++) {
                        uint64_t lo; uint32_t hi;
                        const uint32_t l = code_bits<W>(m[i], rung, lo, hi);
                        if (BITS <= 16) pk.put32((uint32_t)lo, l);
                        else if (BITS == 32 || l <= 64) pk.put64(lo, l);
                        else { pk.put64(lo, 64); pk.put32(hi, 1); } /* 65 bits at rung 63, reference: QB3encode.h:267-275 */
                    }
                }
            }
            pk.finish(s, e);
            __syncthreads();

            /* complete 16 byte units leave for global memory, the rest is carried */
            const uint32_t B = wbits + total, nu = B >> 7;
            for (uint32_t k = tid; k < nu; k += NT) {
                if ((flushed + k + 1) * 16 <= room)
                    st_stream16(dst + (flushed + k) * 16, reinterpret_cast<const uint4 *>(win)[k]);
            }
            if ((flushed + nu) * 16 > room) overflow = true;
            uint32_t cw = 0;
            if (tid < 4) cw = win[nu * 4 + tid];
            __syncthreads();
            const uint32_t used = (B + 31) >> 5;
            for (uint32_t i = 4 + tid; i <= used; i += NT) win[i] = 0;
            if (tid < 4) win[tid] = cw;
            flushed += nu;
            wbits = B & 127;
        }
    }
    __syncthreads();

    /* running state out (reference: QB3encode.h:446-449); of a tile in parts, the last part's */
    if (a.state && (!multi || part + 1 == a.parts)) {
        unsigned long long *st = a.state + (uint64_t)tile * 3 * a.bands;
        for (uint32_t c = tid; c < a.bands; c += NT) {
            st[c] = carry_prev[(it & 1) * a.bands + c] & lowmask64(BITS);
            st[a.bands + c] = carry_rung[(it & 1) * a.bands + c];
            if (BEST) st[2 * a.bands + c] = carry_pcf[(it & 1) * a.bands + c];
        }
    }

    uint64_t len_bytes = flushed * 16 + ((wbits + 7) >> 3);
    if ((flushed + 1) * 16 > room) overflow = true;
    if (wbits && !overflow && tid == 0 && !a.size_only)
        st_stream16(dst + flushed * 16, reinterpret_cast<const uint4 *>(win)[0]);
    if (multi) { /* stitch_kernel puts the parts together and settles size, status and the stored fallback */
        if (tid == 0) a.part_bits[(uint64_t)tile * a.parts + part] = overflow ? ~0ull : flushed * 128 + wbits;
        if (BEST && a.best_pass == 1) {
            for (uint32_t c = tid; c < a.bands; c += NT) {
                unsigned long long *pp = a.part_pcf + (((uint64_t)tile * a.parts + part) * a.bands + c) * 4;
                pp[0] = carry_pcf[(it & 1) * a.bands + c];
                pp[1] = carry_unk[(it & 1) * a.bands + c] ? 0ull : 1ull;
                pp[2] = dep_mask[c];
            }
        }
        return;
    }

    /* stored fallback when coding did not shrink the tile (reference: QB3encode.cpp:570-573, 461-485) */
    /* with an RLE mode the choice is left to rle_kernel: the reference tries RLE first (QB3encode.cpp:536-573) */
    if (a.small == 3 || overflow || (!a.rle_mode && a.raw_size <= len_bytes)) {
        __syncthreads();
        const uint64_t line = (uint64_t)a.w * a.bands * sizeof(T), pitch = a.stride * sizeof(T);
        if (!a.size_only) {
            for (uint32_t i = tid; i < a.hdr_stored_len; i += NT) dst[i] = a.hdr_stored[i];
            for (uint64_t i = tid; i < a.raw_size; i += NT) {
                const uint64_t y = i / line, x = i - y * line;
                dst[a.hdr_stored_len + i] = src[y * pitch + x];
            }
        }
        len_bytes = a.hdr_stored_len + a.raw_size;
    }
    if (tid == 0) {
        a.sizes[tile] = len_bytes;
        if (a.status) a.status[tile] = 0;
    }
}


/* ------------------------------------------------------------------ BEST in parts: the factors handed down */

/* One CTA per tile, a thread per band: walks the parts in order (EncArgs::part_pcf), gives every part the factor its
   band starts from, and marks the parts that met that very factor while they did not know it. The tile's running
   state gets the factor the last part leaves behind. */
__global__ void __launch_bounds__(256) best_resolve_kernel(const __grid_constant__ EncArgs a)
{
    const uint32_t tile = blockIdx.x;
    for (uint32_t c = threadIdx.x; c < a.bands; c += blockDim.x) {
        unsigned long long *pp = a.part_pcf + ((uint64_t)tile * a.parts * a.bands + c) * 4;
        unsigned long long cur = pp[3]; /* the first part knew where it started */
        for (uint32_t q = 0; q < a.parts; q++, pp += 4 * (uint64_t)a.bands) {
            if (q > 0) {
                pp[3] = cur;
                if ((pp[2] >> (cur & 63)) & 1) a.part_redo[(uint64_t)tile * a.parts + q] = 1;
            }
            if (pp[1]) cur = pp[0];
        }
        if (a.state) a.state[(uint64_t)tile * 3 * a.bands + 2 * a.bands + c] = cur;
    }
}

/* ------------------------------------------------------------------ joining the parts of a tile */

/*
 * A tile coded in parts (encode_kernel with a.parts > 1): part p holds len_p bits from bit 0 of its region. The stream is
 * their concatenation; CTA (p, tile) writes the 32 bit words of the slot whose first bit lies in part p, taking the
 * low bits of its first word from the tail of part p - 1. Then, as at the end of encode_kernel: size, status, and
 * the stored fallback when coding did not pay or a part did not fit (reference: QB3encode.cpp:570-573, 461-485).
 */
__global__ void __launch_bounds__(256) stitch_kernel(const __grid_constant__ EncArgs a, uint32_t tsize)
{
    const uint32_t part = blockIdx.x, tile = blockIdx.y, tid = threadIdx.x, NT = blockDim.x;
    const unsigned long long *lens = a.part_bits + (uint64_t)tile * a.parts;
    unsigned long long start = 0, total = 0;
    bool overflow = false;
    for (uint32_t q = 0; q < a.parts; q++) {
        const unsigned long long l = lens[q];
        overflow |= l == ~0ull;
        if (q < part) start += l;
        total += l;
    }
    uint8_t *dst = a.dst + (uint64_t)tile * a.slot;
    const uint8_t *src = a.src + (uint64_t)tile * a.src_pitch;
    const uint64_t len_bytes = (total + 7) >> 3;
    overflow |= ((len_bytes + 15) & ~15ull) > a.slot;
    if (a.size_only) {
        if (part == 0 && tid == 0)
            a.sizes[tile] = overflow || (!a.rle_mode && a.raw_size <= len_bytes) ? a.hdr_stored_len + a.raw_size : len_bytes;
        return;
    }
    if (overflow || (!a.rle_mode && a.raw_size <= len_bytes)) {
        if (part == 0)
            for (uint32_t i = tid; i < a.hdr_stored_len; i += NT) dst[i] = a.hdr_stored[i];
        const uint64_t line = (uint64_t)a.w * a.bands * tsize, pitch = a.stride * tsize;
        const uint64_t per = (a.raw_size + a.parts - 1) / a.parts, lo = per * part, hi = min(a.raw_size, lo + per);
        for (uint64_t i = lo + tid; i < hi; i += NT) {
            const uint64_t y = i / line, x = i - y * line;
            dst[a.hdr_stored_len + i] = src[y * pitch + x];
        }
        if (part == 0 && tid == 0) {
            a.sizes[tile] = a.hdr_stored_len + a.raw_size;
            if (a.status) a.status[tile] = 0;
        }
        return;
    }
    const unsigned long long len = lens[part], end = start + len;
    const uint32_t *in = reinterpret_cast<const uint32_t *>(a.tmp + ((uint64_t)tile * a.parts + part) * a.tmp_slot);
    uint32_t *out = reinterpret_cast<uint32_t *>(dst);
    const uint32_t sh = (uint32_t)start & 31;           /* bits of the first word that belong to the parts before */
    const unsigned long long w0 = start >> 5;
    /* words whose first bit is in [start, end), and the word start falls into when it begins inside it */
    const unsigned long long wfirst = sh ? w0 + 1 : w0, wend = (end + 31) >> 5, nin = (len + 31) >> 5;
    auto word_at = [&](long long o) -> uint32_t { /* 32 bits of this part from local bit offset o, zeros outside */
        const long long i = o >> 5;
        const uint32_t r = (uint32_t)o & 31;
        const uint32_t lo = i >= 0 && (unsigned long long)i < nin ? in[i] : 0u;
        const uint32_t hi = r && i + 1 >= 0 && (unsigned long long)(i + 1) < nin ? in[i + 1] : 0u;
        return r ? (lo >> r) | (hi << (32 - r)) : lo;
    };
    if (len) {
        if (sh && tid == 0) {
            /* The shared word, from whatever parts have bits in it: usually the tail of the part before and the head
               of this one, but a part can be a few bits or none at all (a BEST part whose groups the reference's
               encoder drops, QB3encode.h:704-708), so every part is asked. */
            uint32_t v = 0;
            unsigned long long qs = 0;
            for (uint32_t q = 0; q < a.parts; q++) {
                const unsigned long long ql = lens[q], lo = 32 * w0 > qs ? 32 * w0 : qs, hi = 32 * w0 + 32 < qs + ql ? 32 * w0 + 32 : qs + ql;
                if (lo < hi) {
                    const uint32_t *pin = reinterpret_cast<const uint32_t *>(a.tmp + ((uint64_t)tile * a.parts + q) * a.tmp_slot);
                    const unsigned long long o = lo - qs, i = o >> 5, pn = (ql + 31) >> 5;
                    const uint32_t r = (uint32_t)o & 31, n = (uint32_t)(hi - lo);
                    const uint32_t x0 = pin[i], x1 = r && i + 1 < pn ? pin[i + 1] : 0u;
                    const uint32_t bits = (r ? (x0 >> r) | (x1 << (32 - r)) : x0) & (n < 32 ? (1u << n) - 1 : ~0u);
                    v |= bits << (uint32_t)(lo - 32 * w0);
                }
                qs += ql;
            }
            out[w0] = v;
        }
        /* a word that ends beyond this part belongs to the next part that has bits, unless there is none */
        bool more = false;
        for (uint32_t q = part + 1; q < a.parts; q++) more |= lens[q] != 0;
        const unsigned long long wstop = (more && (end & 31)) ? end >> 5 : wend;
        for (unsigned long long w = wfirst + tid; w < wstop; w += NT)
            out[w] = word_at((long long)(32 * w) - (long long)start);
    }
    if (part == 0 && tid == 0) {
        a.sizes[tile] = len_bytes;
        if (a.status) a.status[tile] = 0;
    }
}

/* ------------------------------------------------------------------ RLE0 byte pass */

/*
 * RLE0 / RLE0Size (reference: QB3encode.cpp:271-332): "FF FF" becomes "FF FF FF", four or more zero bytes become
 * "FF FF n" (n = count - 4, at most 0xfe) unless the byte emitted just before was a literal FF; the last two bytes are
 * always literal. The transducer is serial, but its state is forgotten at every byte that is neither 00 nor FF: such a
 * byte can only leave as a literal, and after it the scanner stands at a token start knowing all it needs (the last
 * byte out was not FF). So the data is cut into chunks that begin right after such a byte (rle_sync), the chunks are
 * measured and written independently by the warps of a CTA, and a scan of the chunk sizes places them.
 *
 * rle_range works through [i0, e) of the n bytes at p, e a chunk start or n. 256 bytes a step, eight to a lane, as
 * aligned words; a step without an equal pair of 00 or FF (nearly all of them, in a compressed stream) is found out
 * with a few word operations and leaves as literals at once; otherwise the literals in front of the first candidate
 * go together and the candidate is resolved on its own. out == nullptr only measures. Returns the output size
 * (warp uniform).
 */
__device__ static uint64_t rle_range(const uint8_t *p, uint64_t n, uint64_t i0, uint64_t e, uint8_t *out)
{
    const uint32_t FULL = 0xffffffffu, lane = lane_id();
    constexpr int BPL = 8;
    const uint64_t lim = n >= 2 ? n - 2 : 0; /* a pair or run cannot start in the last two bytes */
    const uint64_t limr = lim < e ? lim : e;
    uint64_t i = i0, o = 0;
    uint32_t last = 0;
    while (i < limr) {
        /* the lane's eight bytes from position i + 8 * lane: three aligned words, moved into place. Bytes past n may be
           anything (they are inside the tile's slot): no position from limr on is ever a candidate, and literals are
           counted, not taken on trust */
        const uint64_t pos = i + BPL * lane;
        const uint8_t *q = p + pos;
        const uint32_t sh = 8 * (uint32_t)((uintptr_t)q & 3);
        const uint32_t *qa = reinterpret_cast<const uint32_t *>(q - (sh >> 3));
        const bool in = pos < n;
        const uint32_t w0 = in ? qa[0] : 0x01010101u, w1 = in ? qa[1] : 0x01010101u, w2 = in ? qa[2] : 0x01010101u;
        const uint32_t x0 = __funnelshift_r(w0, w1, sh), x1 = __funnelshift_r(w1, w2, sh);
        uint32_t nx = __shfl_down_sync(FULL, x0, 1); /* the byte after the lane's eight */
        if (lane == 31) nx = __funnelshift_r(w2, 0u, sh);
        uint32_t k = 32 * BPL; /* first candidate of the step */
        {
            const uint32_t y0 = __funnelshift_r(x0, x1, 8), y1 = __funnelshift_r(x1, nx, 8); /* each byte's successor */
            auto has_zero = [](uint32_t v) { return (v - 0x01010101u) & ~v & 0x80808080u; };
            const uint32_t any = has_zero(x0 | y0) | has_zero(x1 | y1) | has_zero(~(x0 & y0)) | has_zero(~(x1 & y1));
            if (__any_sync(FULL, any != 0) || i + 32 * BPL > limr) {
                uint32_t b[BPL + 1];
#pragma unroll
                for (int j = 0; j < 4; j++) { b[j] = (x0 >> (8 * j)) & 0xff; b[4 + j] = (x1 >> (8 * j)) & 0xff; }
                b[BPL] = nx & 0xff;
#pragma unroll
                for (int j = 0; j < BPL; j++) {
                    const bool cand = pos + j < limr && b[j] == b[j + 1] && (b[j] == 0 || b[j] == 0xff);
                    const uint32_t mask = __ballot_sync(FULL, cand);
                    if (mask) k = min(k, BPL * ((uint32_t)__ffs((int)mask) - 1) + j);
                }
            }
        }
        auto byte_at = [&](uint32_t idx) -> uint32_t { /* byte idx of the step (warp uniform idx) */
            const uint32_t j = idx % BPL;
            const uint32_t v = ((j < 4 ? x0 : x1) >> (8 * (j & 3))) & 0xff;
            return __shfl_sync(FULL, v, idx / BPL);
        };
        const uint32_t nlit = (uint32_t)min((uint64_t)k, limr - i);
        if (nlit) {
            if (out) {
#pragma unroll
                for (int j = 0; j < BPL; j++)
                    if (BPL * lane + j < nlit) out[o + BPL * lane + j] = (uint8_t)(((j < 4 ? x0 : x1) >> (8 * (j & 3))) & 0xff);
            }
            last = byte_at(nlit - 1);
            o += nlit;
            i += nlit;
        }
        if (k == 32 * BPL || i >= limr) continue;
        const uint32_t c = byte_at(k); /* the candidate, now at position i */
        if (c == 0xff) {
            if (out && lane < 3) out[o + lane] = 0xff;
            o += 3; i += 2; last = 0;
            continue;
        }
        const bool run = last != 0xff && n - (i + 1) >= 3 && p[i + 2] == 0 && p[i + 3] == 0;
        if (!run) {
            if (out && lane == 0) out[o] = 0;
            o += 1; i += 1; last = 0;
            continue;
        }
        i += 4;
        uint32_t r = 0; /* zeros beyond the first four */
        for (;;) {
            const uint64_t qq = i + r + lane;
            const bool z = qq < n && r + lane < 0xfe && p[qq] == 0;
            const uint32_t zm = __ballot_sync(FULL, z);
            const uint32_t here = zm == FULL ? 32 : (uint32_t)__ffs((int)~zm) - 1;
            r += here;
            if (here < 32) break;
        }
        if (out && lane < 3) out[o + lane] = lane < 2 ? 0xff : (uint8_t)r;
        o += 3; i += r; last = 0;
    }
    while (i < e) { /* the literal tail */
        const uint64_t pos = i + lane;
        const uint32_t cnt = (uint32_t)min((uint64_t)32, e - i);
        if (out && pos < e) out[o + lane] = p[pos];
        o += cnt; i += cnt;
    }
    return o;
}

/* the first chunk start at or after position from: a position whose predecessor is neither 00 nor FF; n when there
   is none (warp uniform) */
__device__ static uint64_t rle_sync(const uint8_t *p, uint64_t n, uint64_t from)
{
    if (from == 0) return 0;
    for (uint64_t q = from; q < n; q += 32) {
        const uint64_t at = q + lane_id();
        const uint32_t v = at <= n ? p[at - 1] : 0u;
        const uint32_t m = __ballot_sync(0xffffffffu, at <= n && v != 0 && v != 0xff);
        if (m) return q + (uint32_t)__ffs((int)m) - 1;
    }
    return n;
}

/* One CTA per tile, after encode_kernel, for the RLE modes only: RLE when it pays, else the stored check
   (reference: QB3encode.cpp:536-573). */
constexpr uint32_t RLE_THREADS = 512, RLE_MAXCHUNKS = 2048;
/* blockDim and maxch (the chunk tables in dynamic shared memory) follow the size a stream can have: a batch of 64 x 64
   tiles gets two warps and a dozen table entries per tile, not sixteen warps and 16 KB */
__global__ void __launch_bounds__(RLE_THREADS) rle_kernel(const __grid_constant__ EncArgs a, uint32_t ntiles, uint32_t maxch)
{
    extern __shared__ __align__(16) uint32_t rle_smem[];
    uint32_t *sync_s = rle_smem;                 /* [maxch + 1] chunk starts; streams are far below 4 GB */
    uint32_t *size_s = sync_s + maxch + 1;       /* [maxch] */
    uint32_t *scratch = size_s + maxch;          /* [36] */
    const uint32_t tile = blockIdx.x, tid = threadIdx.x, NT = blockDim.x, lane = tid & 31, warp = tid >> 5, nwarps = NT / 32;
    uint8_t *dst = a.dst + (uint64_t)tile * a.slot;
    const uint8_t *src = a.src + (uint64_t)tile * a.src_pitch;
    if (dst[10] == M_STORED) return; /* already final */
    const uint64_t len = a.sizes[tile], hdr = a.hdr_len, data = len - hdr;
    if (len <= a.max_size / 2 && data < 0xffffffffull) { /* "a vague limit", but it decides the bytes */
        const uint64_t avail = a.max_size - len;
        const uint8_t *p = dst + hdr;
        /* chunks of at least 1 KB, at most maxch of them */
        uint64_t csize = 1024;
        while ((data + csize - 1) / csize > maxch) csize *= 2;
        const uint32_t nch = (uint32_t)((data + csize - 1) / csize);
        for (uint32_t k = warp; k < nch; k += nwarps) {
            const uint64_t s = rle_sync(p, data, k * csize);
            if (lane == 0) sync_s[k] = (uint32_t)s;
        }
        if (tid == 0) sync_s[nch] = (uint32_t)data;
        __syncthreads();
        for (uint32_t k = warp; k < nch; k += nwarps) {
            const uint64_t s = sync_s[k], e = sync_s[k + 1];
            const uint64_t z = s < e ? rle_range(p, data, s, e, nullptr) : 0;
            if (lane == 0) size_s[k] = (uint32_t)z;
        }
        __syncthreads();
        /* exclusive scan of the chunk sizes, a CTA's worth at a time */
        uint64_t rsz = 0;
        for (uint32_t k0 = 0; k0 < nch; k0 += NT) {
            const uint32_t k = k0 + tid, v = k < nch ? size_s[k] : 0u;
            uint32_t total;
            const uint32_t off = block_exclusive_scan(v, scratch, total);
            __syncthreads();
            if (k < nch) size_s[k] = (uint32_t)rsz + off; /* now the chunk's place in the output */
            rsz += total;
        }
        __syncthreads();
        if (rsz <= avail && rsz < data) {
            for (uint32_t k = warp; k < nch; k += nwarps) {
                const uint64_t s = sync_s[k], e = sync_s[k + 1];
                if (s < e) rle_range(p, data, s, e, dst + len + size_s[k]); /* into the free tail of the slot */
            }
            __syncthreads();
            /* down over the data: the output is shorter than the data, so the two ranges do not overlap */
            for (uint64_t k = tid; k < rsz; k += NT) dst[hdr + k] = dst[len + k];
            if (tid == 0) {
                dst[10] = (uint8_t)a.rle_mode;
                a.sizes[tile] = hdr + rsz;
            }
            return;
        }
    }
    if (a.raw_size <= len) { /* stored fallback, reference: QB3encode.cpp:461-485 */
        const uint32_t ts = a.raw_size / ((uint64_t)a.w * a.h * a.bands);
        const uint64_t line = (uint64_t)a.w * a.bands * ts, pitch = a.stride * ts;
        __syncthreads();
        for (uint32_t i = tid; i < a.hdr_stored_len; i += NT) dst[i] = a.hdr_stored[i];
        for (uint64_t i = tid; i < a.raw_size; i += NT) {
            const uint64_t y = i / line, x = i - y * line;
            dst[a.hdr_stored_len + i] = src[y * pitch + x];
        }
        if (tid == 0) a.sizes[tile] = a.hdr_stored_len + a.raw_size;
    }
}

/*
 * The same for few, large tiles: a CTA would walk tens of megabytes alone (19 streams of 27 MB took 32 ms), so S CTAs share
 * a tile and the steps of rle_kernel become launches: 0 measures the chunks, 1 (one CTA per tile) places them and
 * decides, 2 writes them into the free tail of the slot, 3 moves them down or stores the tile. What the later steps
 * need to know about a tile is settled in step 1 (RleWide::meta), for the last step changes the tile's size and mode
 * byte while its other CTAs are still at work.
 */
struct RleWide {
    uint32_t *sync;             /* [tile][RLE_MAXCHUNKS + 1] chunk starts */
    uint32_t *size;             /* [tile][RLE_MAXCHUNKS] chunk sizes, then their places in the output */
    unsigned long long *meta;   /* [tile][4]: what to do (0 nothing, 1 RLE, 2 store), stream length, RLE size, unused */
    uint32_t S;
};
__device__ static uint64_t rle_chunk_size(uint64_t data, uint32_t &nch)
{
    uint64_t csize = 1024; /* chunks of at least 1 KB, at most RLE_MAXCHUNKS of them */
    while ((data + csize - 1) / csize > RLE_MAXCHUNKS) csize *= 2;
    nch = (uint32_t)((data + csize - 1) / csize);
    return csize;
}
__global__ void __launch_bounds__(RLE_THREADS) rle_wide_kernel(const __grid_constant__ EncArgs a, const RleWide w, uint32_t phase)
{
    __shared__ uint32_t scratch[36];
    const uint32_t tile = blockIdx.y, part = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = RLE_THREADS / 32;
    uint8_t *dst = a.dst + (uint64_t)tile * a.slot;
    const uint8_t *src = a.src + (uint64_t)tile * a.src_pitch;
    const uint64_t hdr = a.hdr_len;
    uint32_t *sync_g = w.sync + (uint64_t)tile * (RLE_MAXCHUNKS + 1), *size_g = w.size + (uint64_t)tile * RLE_MAXCHUNKS;
    unsigned long long *meta = w.meta + 4 * (uint64_t)tile;
    if (phase <= 1) {
        const uint64_t len = a.sizes[tile], data = len - hdr;
        const bool stored = dst[10] == M_STORED, try_rle = !stored && len <= a.max_size / 2 && data < 0xffffffffull;
        uint32_t nch = 0;
        const uint64_t csize = try_rle ? rle_chunk_size(data, nch) : 0;
        const uint8_t *p = dst + hdr;
        if (phase == 0) {
            const uint32_t per = (nch + w.S - 1) / w.S, k1 = min(nch, (part + 1) * per);
            for (uint32_t k = part * per + warp; k < k1; k += nwarps) {
                const uint64_t s = rle_sync(p, data, k * csize), e = k + 1 < nch ? rle_sync(p, data, (k + 1) * csize) : data;
                const uint64_t z = s < e ? rle_range(p, data, s, e, nullptr) : 0;
                if (lane == 0) { sync_g[k] = (uint32_t)s; size_g[k] = (uint32_t)z; }
            }
            return;
        }
        uint64_t rsz = 0; /* phase 1: the chunks' places, and what becomes of the tile */
        for (uint32_t k0 = 0; k0 < nch; k0 += RLE_THREADS) {
            const uint32_t k = k0 + tid, v = k < nch ? size_g[k] : 0u;
            uint32_t total;
            const uint32_t off = block_exclusive_scan(v, scratch, total);
            __syncthreads();
            if (k < nch) size_g[k] = (uint32_t)rsz + off;
            rsz += total;
        }
        if (tid == 0) {
            sync_g[nch] = (uint32_t)data;
            const bool rle = try_rle && rsz <= a.max_size - len && rsz < data;
            meta[0] = rle ? 1 : (!stored && a.raw_size <= len) ? 2 : 0;
            meta[1] = len;
            meta[2] = rsz;
        }
        return;
    }
    const uint64_t what = meta[0], len = meta[1], rsz = meta[2], data = len - hdr;
    if (phase == 2) {
        if (what != 1) return;
        uint32_t nch = 0;
        rle_chunk_size(data, nch);
        const uint32_t per = (nch + w.S - 1) / w.S, k1 = min(nch, (part + 1) * per);
        for (uint32_t k = part * per + warp; k < k1; k += nwarps) {
            const uint64_t s = sync_g[k], e = sync_g[k + 1];
            if (s < e) rle_range(dst + hdr, data, s, e, dst + len + size_g[k]);
        }
        return;
    }
    if (what == 1) { /* down over the data: the output is shorter than the data, so the two ranges do not overlap */
        const uint64_t per = (rsz + w.S - 1) / w.S, lo = per * part, hi = min(rsz, lo + per);
        for (uint64_t k = lo + tid; k < hi; k += RLE_THREADS) dst[hdr + k] = dst[len + k];
        if (part == 0 && tid == 0) {
            dst[10] = (uint8_t)a.rle_mode;
            a.sizes[tile] = hdr + rsz;
        }
    }
    else if (what == 2) { /* stored fallback, reference: QB3encode.cpp:461-485 */
        const uint32_t ts = a.raw_size / ((uint64_t)a.w * a.h * a.bands);
        const uint64_t line = (uint64_t)a.w * a.bands * ts, pitch = a.stride * ts;
        const uint64_t per = (a.raw_size + w.S - 1) / w.S, lo = per * part, hi = min(a.raw_size, lo + per);
        if (part == 0)
            for (uint32_t i = tid; i < a.hdr_stored_len; i += RLE_THREADS) dst[i] = a.hdr_stored[i];
        for (uint64_t i = lo + tid; i < hi; i += RLE_THREADS) {
            const uint64_t y = i / line, x = i - y * line;
            dst[a.hdr_stored_len + i] = src[y * pitch + x];
        }
        if (part == 0 && tid == 0) a.sizes[tile] = a.hdr_stored_len + a.raw_size;
    }
}

cudaMemPool_t scratch_pool(); /* qb3_decode.cu */
void count_launches(uint64_t n); /* qb3_cabi.cu */

/* the RLE pass of a batch: one CTA per tile, or for few tiles several (rle_wide_kernel) */
static cudaError_t launch_rle(const EncArgs &a, size_t ntiles, cudaStream_t st)
{
    int dev = 0, nsm = 148;
    if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, dev);
    /* streams of less than 64 KB are a handful of chunks: nothing to share */
    if (ntiles >= (size_t)2 * nsm || a.max_size / 2 < 65536) {
        /* a warp for every 4 KB a stream can have when RLE is tried at all, a table entry for every 1 KB */
        const uint64_t most = a.max_size / 2;
        uint32_t maxch = (uint32_t)(most / 1024 + 2 < RLE_MAXCHUNKS ? most / 1024 + 2 : RLE_MAXCHUNKS);
        uint32_t threads = (uint32_t)(32 * ((most + 4095) / 4096) < RLE_THREADS ? 32 * ((most + 4095) / 4096) : RLE_THREADS);
        if (threads < 32) threads = 32;
        rle_kernel<<<(unsigned)ntiles, threads, (2 * maxch + 1 + 36) * 4, st>>>(a, (uint32_t)ntiles, maxch);
        return cudaGetLastError();
    }
    RleWide w;
    w.S = (uint32_t)(((size_t)4 * nsm + ntiles - 1) / ntiles);
    if (w.S > 128) w.S = 128;
    const size_t sync_bytes = ntiles * (RLE_MAXCHUNKS + 1) * 4, size_bytes = ntiles * RLE_MAXCHUNKS * 4;
    uint8_t *tmp = nullptr;
    cudaMemPool_t pool = scratch_pool();
    const size_t bytes = ((sync_bytes + size_bytes + 15) & ~(size_t)15) + ntiles * 32;
    cudaError_t err = pool ? cudaMallocFromPoolAsync(reinterpret_cast<void **>(&tmp), bytes, pool, st)
                           : cudaMallocAsync(reinterpret_cast<void **>(&tmp), bytes, st);
    if (err != cudaSuccess) return err;
    w.sync = reinterpret_cast<uint32_t *>(tmp);
    w.size = reinterpret_cast<uint32_t *>(tmp + sync_bytes);
    w.meta = reinterpret_cast<unsigned long long *>(tmp + ((sync_bytes + size_bytes + 15) & ~(size_t)15));
    for (uint32_t phase = 0; phase < 4 && err == cudaSuccess; phase++) {
        rle_wide_kernel<<<dim3(phase == 1 ? 1 : w.S, (unsigned)ntiles), RLE_THREADS, 0, st>>>(a, w, phase);
        err = cudaGetLastError();
    }
    const cudaError_t ferr = cudaFreeAsync(tmp, st);
    if (err == cudaSuccess) count_launches(3); /* the caller counts one launch for the pass */
    return err != cudaSuccess ? err : ferr;
}

/* ------------------------------------------------------------------ stream packing */

/* offsets[t] = sum of the 16 byte rounded sizes before t, total[0] = their sum. One CTA; tiles in chunks of blockDim. */
__global__ void __launch_bounds__(1024) pack_offsets_kernel(const unsigned long long *sizes, unsigned long long *offsets,
                                                            unsigned long long *total, uint32_t ntiles)
{
    __shared__ unsigned long long warp_sum[32];
    __shared__ unsigned long long run_s;
    const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) run_s = 0;
    __syncthreads();
    for (uint32_t t0 = 0; t0 < ntiles; t0 += blockDim.x) {
        const uint32_t t = t0 + tid;
        const unsigned long long v = t < ntiles ? (sizes[t] + 15) & ~15ull : 0ull;
        unsigned long long inc = v;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const unsigned long long o = __shfl_up_sync(0xffffffffu, inc, d);
            if (lane >= d) inc += o;
        }
        if (lane == 31) warp_sum[warp] = inc;
        __syncthreads();
        if (warp == 0) {
            unsigned long long w = lane < (blockDim.x >> 5) ? warp_sum[lane] : 0ull, wi = w;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const unsigned long long o = __shfl_up_sync(0xffffffffu, wi, d);
                if (lane >= d) wi += o;
            }
            warp_sum[lane] = wi - w;
        }
        __syncthreads();
        const unsigned long long run = run_s;
        if (t < ntiles) offsets[t] = run + warp_sum[warp] + inc - v;
        __syncthreads();
        if (tid == blockDim.x - 1) run_s = run + warp_sum[warp] + inc;
        __syncthreads();
    }
    if (tid == 0) total[0] = run_s;
}

/* stream t moves from its slot to packed + offsets[t], 16 bytes per thread; grid (chunks, tiles) */
__global__ void __launch_bounds__(256) pack_copy_kernel(const uint8_t *slots, uint64_t slot, const unsigned long long *sizes,
                                                        const unsigned long long *offsets, uint8_t *packed)
{
    const uint32_t tile = blockIdx.y;
    const uint64_t units = (sizes[tile] + 15) >> 4;
    const uint4 *src = reinterpret_cast<const uint4 *>(slots + (uint64_t)tile * slot);
    uint4 *dst = reinterpret_cast<uint4 *>(packed + offsets[tile]);
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < units; i += (uint64_t)gridDim.x * blockDim.x)
        st_stream16(dst + i, ld_stream16(src + i));
}

cudaError_t launch_pack(const uint8_t *slots, uint64_t slot, const unsigned long long *sizes, uint8_t *packed,
                        unsigned long long *offsets, unsigned long long *total, uint32_t ntiles, cudaStream_t st)
{
    pack_offsets_kernel<<<1, 1024, 0, st>>>(sizes, offsets, total, ntiles);
    cudaError_t err = cudaGetLastError();
    if (err != cudaSuccess) return err;
    /* enough CTAs per tile to cover a slot's worth of 16 byte units with a few iterations each */
    uint32_t chunks = (uint32_t)((slot / 16 + 256 * 8 - 1) / (256 * 8));
    if (chunks < 1) chunks = 1;
    if (chunks > 64) chunks = 64;
    for (uint32_t t0 = 0; t0 < ntiles; t0 += 65535) { /* gridDim.y limit */
        const uint32_t n = ntiles - t0 < 65535 ? ntiles - t0 : 65535;
        pack_copy_kernel<<<dim3(chunks, n), 256, 0, st>>>(slots + (uint64_t)t0 * slot, slot, sizes + t0, offsets + t0, packed);
        err = cudaGetLastError();
        if (err != cudaSuccess) return err;
    }
    return cudaSuccess;
}

/* ------------------------------------------------------------------ launch */

template <typename T> static cudaError_t launch_encode_t(const EncArgs &a, size_t ntiles, uint32_t threads, size_t smem, cudaStream_t st)
{
    const bool best = a.mode == M_CF_Z || a.mode == M_CF_H;
    const int curve = a.order == HILBERT ? 1 : a.order == ZCURVE ? 2 : 0;
    auto kern = best ? (curve == 1 ? encode_kernel<T, true, 1> : curve == 2 ? encode_kernel<T, true, 2> : encode_kernel<T, true, 0>)
                     : (curve == 1 ? encode_kernel<T, false, 1> : curve == 2 ? encode_kernel<T, false, 2> : encode_kernel<T, false, 0>);
    if constexpr (sizeof(T) == 1) {
        if (!best && curve == 1 && threads <= 384) kern = encode_kernel<T, false, 1, true>;
    }
    if constexpr (sizeof(T) <= 2) { /* BEST: two CTAs to an SM (85 registers) hide the wait for the threads with index groups */
        if (best && curve == 1 && threads <= 384) kern = encode_kernel<T, true, 1, true>;
    }
    cudaError_t err = allow_max_smem_of(reinterpret_cast<const void *>(kern));
    if (err != cudaSuccess) return err;
    kern<<<(unsigned)(ntiles * (a.parts > 1 ? a.parts : 1)), threads, smem, st>>>(a);
    err = cudaGetLastError();
    if (err == cudaSuccess && a.parts > 1 && best) { /* EncArgs::best_pass */
        best_resolve_kernel<<<(unsigned)ntiles, 256, 0, st>>>(a);
        EncArgs b = a;
        b.best_pass = 2;
        kern<<<(unsigned)(ntiles * a.parts), threads, smem, st>>>(b);
        err = cudaGetLastError();
    }
    if (err == cudaSuccess && a.parts > 1) {
        stitch_kernel<<<dim3(a.parts, (unsigned)ntiles), 256, 0, st>>>(a, (uint32_t)sizeof(T));
        err = cudaGetLastError();
    }
    if (err != cudaSuccess || !a.rle_mode) return err;
    return launch_rle(a, ntiles, st);
}

cudaError_t launch_encode(const EncArgs &a, uint32_t tsize, size_t ntiles, uint32_t threads, size_t smem, cudaStream_t st)
{
    switch (tsize) {
    case 1: return launch_encode_t<uint8_t>(a, ntiles, threads, smem, st);
    case 2: return launch_encode_t<uint16_t>(a, ntiles, threads, smem, st);
    case 4: return launch_encode_t<uint32_t>(a, ntiles, threads, smem, st);
    default: return launch_encode_t<uint64_t>(a, ntiles, threads, smem, st);
    }
}

} // namespace qb3
