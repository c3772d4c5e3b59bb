/*
 * qb3_decode.cu -- the QB3 decode kernels for sm_100a.
 *
 * Replaces QB3::decodeFTL / QB3::decode / gdecode (QB3decode.h:142-741), the header and chunk parser
 * (QB3decode.cpp:130-264), deRLE0 (:267-307), the small-image scatter (:321-353), stored_decode (:356-375)
 * and dequantize (:77-107).
 *
 * A QB3 stream has no block index and every code's position depends on the two low bits of the code
 * before it, so one stream is one serial bit parse; throughput comes from decoding many streams at
 * once. parse_kernel gives every stream a thread (32 streams per warp, all lanes walking the same
 * rung-uniform arithmetic decode, per band state in shared memory laid out [band][lane]); the values
 * it reconstructs are still band-differenced. finish_kernel then undoes the band difference and the
 * quantisation element-wise over the tiles with coalesced accesses, which is what the reference does
 * per block row (QB3decode.h:730-737) and at the end (QB3decode.cpp:434-450).
 */
#include <cstdlib>
#include <mutex>
#include <type_traits>

#include "qb3_device.cuh"

namespace qb3 {

/* ------------------------------------------------------------------ stream header */

__device__ __forceinline__ uint32_t rd16(const uint8_t *p) { return p[0] | ((uint32_t)p[1] << 8); }

/*
 * Fixed header and chunks (reference: QB3decode.cpp:130-264, doc/QB3.md:228-259). cband_out gets one entry per
 * band with the given element stride. Geometry and type must match what the batch was declared with.
 */
__device__ static void parse_header(const uint8_t *p, uint64_t len, const DecArgs &a, StreamInfo &o,
                                    uint8_t *cband_out, uint32_t cband_stride)
{
    o.order = 0; o.quanta = 1; o.mode = 0; o.data_off = 0; o.has_cb = 0; o.bad = 1;
    for (uint32_t c = 0; c < a.bands; c++) cband_out[c * cband_stride] = a.ref_compat ? 0 : (uint8_t)c;
    if (len < 15 || p[0] != 'Q' || p[1] != 'B' || p[2] != '3' || p[3] != 0x80) return;
    if (rd16(p + 4) + 1 != a.w || rd16(p + 6) + 1 != a.h || (uint32_t)p[8] + 1 != a.bands || p[9] != a.dtype) return;
    o.mode = p[10];
    if ((o.mode > M_FTL && o.mode != M_STORED) || ((p[11] | p[12]) & 0x80)) return;
    if (o.mode <= 3) o.order = ZCURVE;
    uint64_t at = 11;
    for (;;) {
        if (at + 2 > len) return;
        const uint32_t sig = rd16(p + at);
        if (sig == ('D' | ('T' << 8))) {
            at += 2;
            if (at >= len) return;
            o.data_off = (uint32_t)at;
            o.bad = 0;
            return;
        }
        if (at + 4 > len) return;
        const uint32_t clen = rd16(p + at + 2);
        if (sig == ('Q' | ('V' << 8))) {
            if (clen > 4 || clen < 1 || at + 4 + clen > len) return;
            uint64_t q = 0;
            for (uint32_t i = 0; i < clen; i++) q |= (uint64_t)p[at + 4 + i] << (8 * i);
            if (q < 2) return;
            o.quanta = q;
            at += 4 + clen;
        }
        else if (sig == ('C' | ('B' << 8))) {
            if (clen != a.bands || at + 4 + clen > len) return;
            for (uint32_t c = 0; c < a.bands; c++) {
                const uint8_t b = p[at + 4 + c];
                if (b >= a.bands) return;
                cband_out[c * cband_stride] = b;
            }
            o.has_cb = 1;
            at += 4 + clen;
        }
        else if (sig == ('S' | ('C' << 8))) {
            if (clen != 8 || o.mode < 4 || o.mode == M_STORED || at + 12 > len) return;
            uint64_t v = 0;
            for (int i = 0; i < 8; i++) v |= (uint64_t)p[at + 4 + i] << (8 * i);
            uint32_t seen = 0;
            for (int i = 0; i < 16; i++) seen |= 1u << ((v >> (4 * i)) & 15);
            if (seen != 0xffff) return;
            o.order = v;
            at += 12;
        }
        else {
            /* unknown chunk: the reference skips lower case ones by their length only and so never gets past
               them (QB3decode.cpp:254-255); treat every unknown chunk as a bad header */
            return;
        }
    }
}

/* ------------------------------------------------------------------ bit reader */

/*
 * Forward-only LSB-first reader with the reference's semantics (bitstream.h:25-63): zero fill past the end,
 * position saturates at the end. In RLE mode the bytes are expanded on the fly (deRLE0, QB3decode.cpp:267-291)
 * so no scratch copy of the stream is needed.
 */
struct Reader {
    const uint8_t *base; /* 8 byte aligned in direct mode */
    uint64_t phys_len;   /* bytes available from base */
    uint64_t src;        /* next physical byte */
    uint64_t end, pos;   /* logical bit positions */
    uint64_t w0, w1, widx;
    uint32_t run;
    uint8_t runbyte;
    bool rle;

    __device__ __forceinline__ uint64_t next_word()
    {
        if (!rle) {
            uint64_t v = 0;
            if (src + 8 <= phys_len) v = *reinterpret_cast<const uint64_t *>(base + src);
            else for (uint32_t i = 0; src + i < phys_len; i++) v |= (uint64_t)base[src + i] << (8 * i);
            src += 8;
            return v;
        }
        uint64_t v = 0;
        for (uint32_t i = 0; i < 8; i++) {
            uint32_t b = 0;
            if (run) { b = runbyte; run--; }
            else if (src < phys_len) {
                b = base[src++];
                if (b == 0xff && src + 1 < phys_len && base[src] == 0xff) { /* reference: QB3decode.cpp:271-287 */
                    const uint8_t n = base[src + 1];
                    src += 2;
                    if (n == 0xff) { runbyte = 0xff; run = 1; }
                    else { b = 0; runbyte = 0; run = 3 + n; }
                }
            }
            v |= (uint64_t)b << (8 * i);
        }
        return v;
    }
    /* p, len: the payload. For RLE streams logical_len is the expanded size. */
    __device__ __forceinline__ void open(const uint8_t *p, uint64_t len, bool is_rle, uint64_t logical_len)
    {
        rle = is_rle; run = 0; runbyte = 0; src = 0;
        uint32_t mis = 0;
        if (!rle) mis = (uint32_t)((uintptr_t)p & 7);
        base = p - mis;
        phys_len = len + mis;
        pos = 8ull * mis;
        end = pos + 8 * (rle ? logical_len : len);
        widx = 0;
        w0 = next_word();
        w1 = next_word();
    }
    __device__ __forceinline__ uint64_t avail() const { return end - pos; }
    __device__ __forceinline__ uint64_t peek()
    {
        const uint64_t i = pos >> 6;
        while (widx < i) { w0 = w1; w1 = next_word(); widx++; }
        const uint32_t sh = (uint32_t)pos & 63;
        uint64_t v = sh ? (w0 >> sh) | (w1 << (64 - sh)) : w0;
        const uint64_t av = end - pos;
        if (av < 64) v &= lowmask64((uint32_t)av);
        return v;
    }
    __device__ __forceinline__ void advance(uint64_t d) { pos = (pos + d < end) ? pos + d : end; }
    __device__ __forceinline__ uint64_t get(uint32_t n)
    {
        const uint64_t v = peek() & lowmask64(n);
        advance(n);
        return v;
    }
};

/* expanded size of an RLE payload (reference: QB3decode.cpp:294-307) */
__device__ static uint64_t derle_size(const uint8_t *p, uint64_t len)
{
    uint64_t i = 0, n = 0;
    while (i + 2 < len) {
        if (p[i] != 0xff || p[i + 1] != 0xff) { n++; i++; continue; }
        n += p[i + 2] == 0xff ? 2 : 4 + (uint64_t)p[i + 2];
        i += 3;
    }
    return n + (len - i);
}


__device__ __forceinline__ uint32_t lds32(uint32_t addr)
{
    uint32_t v;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr));
    return v;
}

/* dequantize (reference: QB3decode.cpp:77-107): multiply by quanta, saturating at the type's range */
template <int BITS> __device__ __forceinline__ uint64_t dequantize_value(uint64_t v, uint64_t q, bool is_signed)
{
    const uint64_t UM = lowmask64(BITS);
    if (!is_signed) return v <= UM / q ? v * q : UM;
    const long long smax = (long long)(UM >> 1), smin = -smax - 1;
    const long long d = (long long)(v << (64 - BITS)) >> (64 - BITS);
    long long t = d <= smax / (long long)q ? (long long)((uint64_t)d * q) : smax;
    if (q > 2 && d < smin / (long long)q) t = smin;
    return (uint64_t)t & UM;
}

/* tile states that only live between the decode kernels of one batch */
constexpr uint32_t ST_DEFER = 0x80000000u;  /* walk_kernel left the tile to the general path */
constexpr uint32_t ST_FINISH = 0x40000000u; /* pixels are in place but still band differenced / quantised */

/* ------------------------------------------------------------------ group parse */

/* 16 values at a rung (reference: QB3decode.h:142-290); use_step undoes the step-down flip (:285-289) */
template <typename W, typename S> __device__ __forceinline__ void read_group(S &s, uint32_t rung, W (&g)[16], bool use_step)
{
    if (rung == 0) {
        uint32_t b = 0;
        if (s.get(1)) b = (uint32_t)s.get(16);
#pragma unroll
        for (int i = 0; i < 16; i++) g[i] = (b >> i) & 1;
        return;
    }
    uint32_t M = 0;
#pragma unroll
    for (int i = 0; i < 16; i++) {
        const uint64_t x = s.peek();
        uint32_t len, x64 = 0;
        if (sizeof(W) == 8 && rung == 63 && (x & 3) == 3) { /* 65 bit code, reference: QB3decode.h:272-282 */
            S t = s;
            t.advance(64);
            x64 = (uint32_t)t.peek() & 1;
        }
        W v = (W)decode_bits(x, x64, rung, len);
        if (group_swaps(rung)) v = mswap(v, rung);
        s.advance(len);
        g[i] = v;
        M |= ((uint32_t)(v >> rung) & 1u) << i;
    }
    if (use_step) {
        const int k = step_decode_index(M);
#pragma unroll
        for (int i = 0; i < 16; i++) if (i == k) g[i] ^= (W)1 << rung;
    }
}

/* stand-alone value (reference: qb3dsztbl, QB3decode.h:132-138) */
template <typename S> __device__ __forceinline__ uint64_t read_single(S &s, uint32_t rung)
{
    if (rung == 0) return s.get(1);
    uint32_t len;
    uint64_t v = decode_bits(s.peek(), 0, rung, len);
    if (single_swaps(rung)) v = mswap(v, rung);
    s.advance(len);
    return v;
}


/*
 * The group that follows a SIGNAL: a common factor group or an index group (reference: QB3decode.h:624-716).
 * rb is the band's running rung, pcf its last common factor; both are updated. Returns true on the reference's
 * failure conditions (QB3decode.h:642,665,683,703).
 */
template <typename W, int BITS, int U, typename S>
__device__ __noinline__ bool read_special_group(S &s, W (&g)[16], uint8_t &rb, W &pcf)
{
    constexpr uint32_t UMASK = (1u << U) - 1, LMASK = 2 * UMASK + 1;
    const W TM = (W)lowmask64(BITS);
    bool failed = false;
    uint32_t cs = ds_entry(U, (uint32_t)s.peek() & LMASK);
    uint32_t rung = (rb + cs) & UMASK;
    s.advance((cs >> 12) - 1);
    if (rung != UMASK) { /* common factor */
        uint32_t cfrung = rung;
        W cf = pcf;
        if (s.get(1)) {
            const uint32_t own = (uint32_t)s.get(1);
            if (own) {
                cs = ds_entry(U, (uint32_t)s.peek() & LMASK);
                cfrung = (rung + cs) & UMASK;
                failed |= cfrung == rung;
                s.advance((cs >> 12) - 1);
            }
            if (own && cfrung == 0) return true; /* the reference indexes a table at -1 here */
            cf = (W)(read_single(s, cfrung - own) + ((uint64_t)own << cfrung)) & TM;
            pcf = cf;
        }
        cf = (cf + 2) & TM;
        if (rung) {
            W used = 0;
            read_group<W>(s, rung, g, true);
#pragma unroll
            for (int i = 0; i < 16; i++) /* magsmul, reference: QB3decode.h:575 */
                used |= g[i] = (magsabs(g[i]) * (W)(cf << 1) - (g[i] & 1)) & TM;
            rb = (uint8_t)topbit((W)(used | 1));
            failed |= cf > used;
        }
        else {
            const W v = (((cf - 1) << 1) | 1) & TM;
            const uint32_t b = (uint32_t)s.get(16);
#pragma unroll
            for (int i = 0; i < 16; i++) g[i] = ((b >> i) & 1) ? v : (W)0;
            rb = (uint8_t)topbit((W)(v | 1));
        }
        return failed;
    }
    /* index group */
    W tbl[8];
    uint32_t maxidx = 0, used = 0;
    uint64_t idx = 0;
    cs = ds_entry(U, (uint32_t)s.peek() & LMASK);
    rung = (rb + cs) & UMASK;
    rb = (uint8_t)rung;
    failed |= rung == 63;
    s.advance((cs >> 12) - 1);
#pragma unroll
    for (int i = 0; i < 16; i++) {
        uint32_t l;
        const uint32_t j = (uint32_t)decode_bits(s.peek(), 0, 2, l); /* no swap, reference: QB3decode.h:697 */
        s.advance(l);
        used += l;
        idx |= (uint64_t)j << (3 * i);
        maxidx = max(maxidx, j);
    }
    failed |= used > 52;
#pragma unroll
    for (int i = 0; i < 8; i++) tbl[i] = 0;
    for (uint32_t i = 0; i <= maxidx; i++) {
        const W v = (W)read_single(s, rung);
#pragma unroll
        for (int k = 0; k < 8; k++) if (k == (int)i) tbl[k] = v;
    }
#pragma unroll
    for (int i = 0; i < 16; i++) {
        const uint32_t j = (uint32_t)(idx >> (3 * i)) & 7;
        W v = 0;
#pragma unroll
        for (int k = 0; k < 8; k++) if (k == (int)j) v = tbl[k];
        g[i] = v;
    }
    return failed;
}

constexpr uint32_t ST_SCANNING = 0x10000000u; /* decode_kernel has taken the tile (transient, never seen by a caller) */

/* 16 bytes global -> shared, the tail beyond nbytes zero filled */
__device__ __forceinline__ void cp_async16_zfill(uint32_t smem_addr, const void *gmem, uint32_t nbytes)
{
    asm volatile("cp.async.ca.shared.global [%0], [%1], 16, %2;" :: "r"(smem_addr), "l"(gmem), "r"(nbytes) : "memory");
}

/* Position based reader over a lane's ring, for the 32 and 64 bit scan: codes there can be 65 bits, so nothing is
   buffered in registers; a read takes the words it needs from shared memory. Position is in bits from the ring's origin. */
template <int RWORDS> struct RingBits {
    const uint32_t *ring;
    uint32_t pos;
    __device__ __forceinline__ uint32_t peek32() const
    {
        const uint32_t w = pos >> 5;
        return __funnelshift_r(ring[w & (RWORDS - 1)], ring[(w + 1) & (RWORDS - 1)], pos & 31);
    }
    __device__ __forceinline__ uint64_t peek() const
    {
        const uint32_t w = pos >> 5, sh = pos & 31;
        const uint32_t w0 = ring[w & (RWORDS - 1)], w1 = ring[(w + 1) & (RWORDS - 1)], w2 = ring[(w + 2) & (RWORDS - 1)];
        return (uint64_t)__funnelshift_r(w0, w1, sh) | ((uint64_t)__funnelshift_r(w1, w2, sh) << 32);
    }
    __device__ __forceinline__ void advance(uint64_t n) { pos += (uint32_t)n; }
    __device__ __forceinline__ uint64_t get(uint32_t n)
    {
        const uint64_t v = peek() & lowmask64(n);
        pos += n;
        return v;
    }
};

/* Bit reader of the rebuild warps' general path: a thread reads one group at a known bit position straight from global memory
   (neighbouring threads read neighbouring words). Same contract as the others: 33 valid bits after refill(), zeros
   past the end of the payload. */
struct GroupBits {
    const uint32_t *base; /* 4 byte aligned, at or before the payload */
    uint64_t buf;
    uint32_t nb, k, nwords, tailmask;

    __device__ __forceinline__ uint32_t load(uint32_t i) const
    {
        uint32_t w = i < nwords ? __ldg(base + i) : 0u;
        if (i + 1 == nwords) w &= tailmask;
        return w;
    }
    /* bit is counted from the payload's first byte, which sits mis bytes after base */
    __device__ __forceinline__ void open(const uint8_t *payload, uint64_t plen, uint64_t bit)
    {
        const uint32_t mis = (uint32_t)((uintptr_t)payload & 3);
        const uint64_t span = mis + plen;
        const uint32_t tail = (uint32_t)span & 3;
        base = reinterpret_cast<const uint32_t *>(payload - mis);
        nwords = (uint32_t)((span + 3) >> 2);
        tailmask = tail ? (1u << (8 * tail)) - 1 : 0xffffffffu;
        const uint64_t abs = bit + 8 * mis;
        k = (uint32_t)(abs >> 5);
        const uint32_t sh = (uint32_t)abs & 31;
        buf = (uint64_t)(load(k) >> sh);
        nb = 32 - sh;
        buf |= (uint64_t)load(k + 1) << nb;
        nb += 32;
        k += 2;
    }
    __device__ __forceinline__ void refill()
    {
        if (nb <= 32) {
            buf |= (uint64_t)load(k) << nb;
            nb += 32;
            k++;
        }
    }
    __device__ __forceinline__ uint64_t peek() { refill(); return buf; }
    __device__ __forceinline__ void advance(uint32_t n) { buf >>= n; nb -= n; }
    __device__ __forceinline__ uint64_t get(uint32_t n)
    {
        refill();
        const uint64_t v = buf & lowmask64(n);
        advance(n);
        return v;
    }
};

/* The same for 32 and 64 bit types, where one code can be 65 bits long: no register buffer, just a bit position;
   peek() assembles 64 bits from three words. Slower per value, which a throughput kernel can afford. */
struct WideBits {
    const uint32_t *base;
    uint64_t pos; /* bit position from base */
    uint32_t nwords, tailmask;

    __device__ __forceinline__ uint32_t load(uint64_t i) const
    {
        uint32_t w = i < nwords ? __ldg(base + i) : 0u;
        if (i + 1 == nwords) w &= tailmask;
        return w;
    }
    __device__ __forceinline__ void open(const uint8_t *payload, uint64_t plen, uint64_t bit)
    {
        const uint32_t mis = (uint32_t)((uintptr_t)payload & 3);
        const uint64_t span = mis + plen;
        const uint32_t tail = (uint32_t)span & 3;
        base = reinterpret_cast<const uint32_t *>(payload - mis);
        nwords = (uint32_t)((span + 3) >> 2);
        tailmask = tail ? (1u << (8 * tail)) - 1 : 0xffffffffu;
        pos = bit + 8 * mis;
    }
    __device__ __forceinline__ uint64_t peek() const
    {
        const uint64_t k = pos >> 5;
        const uint32_t sh = (uint32_t)pos & 31;
        const uint32_t w0 = load(k), w1 = load(k + 1), w2 = load(k + 2);
        return (uint64_t)__funnelshift_r(w0, w1, sh) | ((uint64_t)__funnelshift_r(w1, w2, sh) << 32);
    }
    __device__ __forceinline__ void advance(uint64_t n) { pos += n; }
    __device__ __forceinline__ uint64_t get(uint32_t n)
    {
        const uint64_t v = peek() & lowmask64(n);
        pos += n;
        return v;
    }
};

#include "qb3_decode_fused.cuh"

template <typename T>
__global__ void __launch_bounds__(32, 1) parse_kernel(const DecArgs a, const bool only_deferred)
{
    typedef typename traits<T>::W W;
    constexpr int BITS = traits<T>::BITS, U = traits<T>::U;
    constexpr uint32_t UMASK = (1u << U) - 1, LMASK = 2 * UMASK + 1;
    const W TM = (W)lowmask64(BITS);

    extern __shared__ __align__(16) uint8_t smem[];
    /* per band state, [band][lane] so that the lanes of a warp hit different banks */
    W *prev = reinterpret_cast<W *>(smem);
    W *pcf = prev + 32 * a.bands;
    uint8_t *runbits = reinterpret_cast<uint8_t *>(pcf + 32 * a.bands);
    uint8_t *cband = runbits + 32 * a.bands;

    const uint32_t lane = threadIdx.x, tile = blockIdx.x * 32 + lane;
    if (tile >= a.ntiles) return;
    if (only_deferred && a.status[tile] != ST_DEFER) return; /* walk_kernel has dealt with it */
    const uint8_t *stream = a.streams + a.offsets[tile];
    const uint64_t slen = a.lens[tile];
    T *out = reinterpret_cast<T *>(a.dst + (uint64_t)tile * a.dst_pitch);

    StreamInfo info;
    parse_header(stream, slen, a, info, cband + lane, 32);
    if (info.bad) { a.status[tile] = QB3CU_TILE_BAD_HEADER; return; }
    const uint8_t *payload = stream + info.data_off;
    const uint64_t plen = slen - info.data_off;
    const uint64_t raw = (uint64_t)a.w * a.h * a.bands * sizeof(T);

    if (info.mode == M_STORED) { /* reference: QB3decode.cpp:356-375 */
        a.status[tile] = plen == raw ? ST_FINISH : (uint32_t)QB3CU_TILE_CORRUPT; /* finish_kernel copies the pixels */
        return;
    }
    if ((uint64_t)a.w * a.h < 16) { a.status[tile] = QB3CU_TILE_CORRUPT; return; } /* reference: QB3decode.cpp:389 */

    const bool rle = info.mode == 2 || info.mode == 3 || info.mode == 6 || info.mode == 7;
    uint64_t logical = plen;
    if (rle) {
        logical = derle_size(payload, plen);
        if (logical > raw) { a.status[tile] = QB3CU_TILE_RLE_TOO_BIG; return; } /* reference: QB3decode.cpp:401 */
    }
    Reader s;
    s.open(payload, plen, rle, logical);

    /* coded geometry: the image or its small-image reorder (reference: QB3decode.cpp:321-329) */
    uint32_t vw = a.w, vh = a.h, small = 0;
    if (a.w < 4 || a.h < 4) {
        const uint32_t ng = (a.w * a.h + 15) / 16;
        if (a.w < 4) { small = 1; vw = 4; vh = ng * 4; } else { small = 2; vw = ng * 4; vh = 4; }
    }
    const uint64_t npixels = (uint64_t)a.w * a.h;
    const uint64_t order = info.order ? info.order : HILBERT;
    const bool ftl = info.mode == M_FTL;
    for (uint32_t c = 0; c < a.bands; c++) { prev[c * 32 + lane] = 0; pcf[c * 32 + lane] = 0; runbits[c * 32 + lane] = 0; }

    const uint32_t nbx = (vw + 3) / 4, nby = (vh + 3) / 4;
    bool failed = false;
    for (uint32_t by = 0; by < nby && !failed; by++) {
        const uint32_t y0 = min(4 * by, vh - 4);
        for (uint32_t bx = 0; bx < nbx && !failed; bx++) {
            const uint32_t x0 = min(4 * bx, vw - 4);
            for (uint32_t c = 0; c < a.bands; c++) {
                W g[16];
                uint32_t cs = 0;
                if (s.get(1)) {
                    cs = ds_entry(U, (uint32_t)s.peek() & LMASK);
                    s.advance((cs >> 12) - 1);
                }
                if (ftl || (cs & 0xfff) != 0 || cs == 0) {
                    const uint32_t rung = (runbits[c * 32 + lane] + cs) & UMASK;
                    runbits[c * 32 + lane] = (uint8_t)rung;
                    read_group<W>(s, rung, g, !ftl);
                }
                else { /* signal: common factor or index group */
                    W sg[16];
                    if (read_special_group<W, BITS, U>(s, sg, runbits[c * 32 + lane], pcf[c * 32 + lane])) { failed = true; break; }
#pragma unroll
                    for (int i = 0; i < 16; i++) g[i] = sg[i];
                }
                /* undo the running delta and scatter (reference: QB3decode.h:717-722) */
                W prv = prev[c * 32 + lane];
#pragma unroll
                for (int i = 0; i < 16; i++) {
                    const uint32_t n = (uint32_t)(order >> (4 * (15 - i))) & 15;
                    prv = (prv + smag<BITS, W>(g[i])) & TM;
                    uint64_t x = x0 + (n & 3), y = y0 + (n >> 2);
                    bool inside = true;
                    if (small == 1) { const uint64_t p = y * 4 + x; inside = p < npixels; y = p / a.w; x = p % a.w; }
                    else if (small == 2) { const uint64_t p = y * vw + x; inside = p < npixels; x = p / a.h; y = p % a.h; }
                    if (inside) out[y * a.stride + x * a.bands + c] = (T)prv;
                }
                prev[c * 32 + lane] = prv;
            }
        }
    }
    if (failed || s.avail() > 7) { a.status[tile] = QB3CU_TILE_CORRUPT; return; } /* reference: QB3decode.h:740 */
    a.status[tile] = ST_FINISH;
}

/*
 * Band difference and quantisation undo, element-wise. One CTA per (tile, row chunk).
 * Bands are processed in ascending order in place, like the reference's sweep (QB3decode.h:730-737),
 * then every value is multiplied by quanta with saturation (QB3decode.cpp:77-107).
 */
template <typename T>
__global__ void __launch_bounds__(256) finish_kernel(const DecArgs a, uint32_t rows_per_cta)
{
    typedef typename traits<T>::W W;
    constexpr int BITS = traits<T>::BITS;
    __shared__ StreamInfo info;
    __shared__ uint8_t cband[MAXBANDS];
    __shared__ uint32_t derived;
    const uint32_t tile = blockIdx.x;
    if (a.status[tile] != ST_FINISH) return;
    if (threadIdx.x == 0) {
        parse_header(a.streams + a.offsets[tile], a.lens[tile], a, info, cband, 1);
        uint32_t d = 0;
        for (uint32_t c = 0; c < a.bands; c++) d |= cband[c] != c;
        derived = d;
    }
    __syncthreads();
    T *out = reinterpret_cast<T *>(a.dst + (uint64_t)tile * a.dst_pitch);
    const uint32_t ybeg = blockIdx.y * rows_per_cta, yend = min(a.h, ybeg + rows_per_cta);
    if (info.mode == M_STORED) { /* raw pixels follow the headers, reference: QB3decode.cpp:356-375 */
        const uint8_t *payload = a.streams + a.offsets[tile] + info.data_off;
        const uint64_t line = (uint64_t)a.w * a.bands * sizeof(T);
        for (uint32_t y = ybeg; y < yend; y++) {
            uint8_t *row = reinterpret_cast<uint8_t *>(out + (uint64_t)y * a.stride);
            for (uint64_t i = threadIdx.x; i < line; i += blockDim.x) row[i] = payload[y * line + i];
        }
        return;
    }
    if (!derived && info.quanta < 2) return;
    const bool is_signed = a.dtype & 1;
    const uint64_t q = info.quanta;
    for (uint32_t y = ybeg; y < yend; y++) {
        T *row = out + (uint64_t)y * a.stride;
        for (uint32_t x = threadIdx.x; x < a.w; x += blockDim.x) {
            T *p = row + (uint64_t)x * a.bands;
            if (derived)
                for (uint32_t c = 0; c < a.bands; c++)
                    if (cband[c] != c) p[c] = (T)(p[c] + p[cband[c]]);
            if (q > 1)
                for (uint32_t c = 0; c < a.bands; c++) p[c] = (T)dequantize_value<BITS>((uint64_t)p[c], q, is_signed);
        }
    }
    (void)sizeof(W);
}

/* the transient bits leave the status words */
__global__ void __launch_bounds__(256) seal_kernel(uint32_t *status, uint32_t ntiles)
{
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t < ntiles && status[t] == ST_FINISH) status[t] = QB3CU_TILE_OK;
}

/* ------------------------------------------------------------------ launch */

/*
 * Scratch memory of the two pass decode comes from a stream ordered pool of our own, one per device:
 *  - it keeps what it has been given (the default is to hand freed memory back to the driver at every
 *    synchronisation, and mapping 800 MB afresh per batch costs more than the kernels)
 *  - it never reuses a block freed on another stream by making the new owner wait for the old one: batches decoded
 *    concurrently on different streams must stay concurrent (they would otherwise run one after the other)
 */
cudaMemPool_t scratch_pool()
{
    static cudaMemPool_t pools[64] = {};
    static std::mutex mu;
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return nullptr;
    std::lock_guard<std::mutex> lock(mu);
    if (!pools[dev]) {
        cudaMemPoolProps props = {};
        props.allocType = cudaMemAllocationTypePinned;
        props.handleTypes = cudaMemHandleTypeNone;
        props.location.type = cudaMemLocationTypeDevice;
        props.location.id = dev;
        if (cudaMemPoolCreate(&pools[dev], &props) != cudaSuccess) { pools[dev] = nullptr; return nullptr; }
        uint64_t keep = ~0ull;
        int off = 0;
        cudaMemPoolSetAttribute(pools[dev], cudaMemPoolAttrReleaseThreshold, &keep);
        cudaMemPoolSetAttribute(pools[dev], cudaMemPoolReuseAllowInternalDependencies, &off);
    }
    return pools[dev];
}

/*
 * RLE streams (modes 2, 3, 6, 7) ahead of the two pass decode: a warp per stream expands the payload (deRLE0,
 * QB3decode.cpp:267-291) into a slot of scratch memory behind a copy of the headers whose mode byte names the plain
 * mode, and the decode then runs on (offs2, lens2), which point there. The two pass kernels never see an RLE stream;
 * without this they leave it to parse_kernel, one lane per stream for the whole decode, twenty times slower. A stream
 * that does not fit its slot, or is not RLE, keeps its place and its path (and its error reporting).
 */
__global__ void __launch_bounds__(128) derle_kernel(const DecArgs a, uint8_t *xbuf, uint64_t xslot, unsigned long long *offs2,
                                                    unsigned long long *lens2)
{
    __shared__ uint8_t cbs[4 * MAXBANDS];
    const uint32_t FULL = 0xffffffffu, lane = threadIdx.x & 31, warp = threadIdx.x >> 5, tile = blockIdx.x * 4 + warp;
    if (tile >= a.ntiles) return;
    const uint64_t off = a.offsets[tile], len = a.lens[tile];
    const uint8_t *stream = a.streams + off;
    StreamInfo info;
    parse_header(stream, len, a, info, cbs + warp * MAXBANDS, 1); /* every lane the same: the warp stays together */
    unsigned long long noff = off, nlen = len;
    const bool rle = info.mode == 2 || info.mode == 3 || info.mode == 6 || info.mode == 7;
    if (!info.bad && rle && info.data_off < 2048) {
        const uint32_t hdr = info.data_off;
        uint8_t *slot = xbuf + (uint64_t)tile * xslot;
        uint8_t *out = slot + 2048, *head = out - hdr; /* payload 16 byte aligned, headers right in front of it */
        const uint64_t cap = xslot - 2048 - 8;
        const uint8_t *p = stream + hdr;
        const uint64_t plen = len - hdr;
        /* A marker is FF FF n with all three bytes inside the payload, found scanning left to right; everything else
           is literal. The warp looks at 32 bytes at a time, copies the literals in front of the first marker together
           and resolves that marker, so the serial rule holds and the rare markers cost one step each. */
        uint64_t i = 0, o = 0;
        bool fits = true;
        while (i < plen && fits) {
            constexpr int BPL = 8; /* bytes per lane and step: independent loads, one latency for 256 bytes */
            const uint64_t pos = i + BPL * lane;
            uint32_t b[BPL + 1];
#pragma unroll
            for (int j = 0; j < BPL; j++) b[j] = pos + j < plen ? p[pos + j] : 0u;
            b[BPL] = __shfl_down_sync(FULL, b[0], 1);
            if (lane == 31) b[BPL] = pos + BPL < plen ? p[pos + BPL] : 0u;
            uint32_t k = 32 * BPL;
#pragma unroll
            for (int j = 0; j < BPL; j++) {
                const bool cand = b[j] == 0xff && b[j + 1] == 0xff && pos + j + 2 < plen;
                const uint32_t mask = __ballot_sync(FULL, cand);
                if (mask) k = min(k, BPL * ((uint32_t)__ffs((int)mask) - 1) + j);
            }
            const uint32_t nlit = (uint32_t)min((uint64_t)k, plen - i);
            fits = o + 32 * BPL + 260 <= cap;
            if (!fits) break;
#pragma unroll
            for (int j = 0; j < BPL; j++)
                if (BPL * lane + j < nlit) out[o + BPL * lane + j] = (uint8_t)b[j];
            o += nlit;
            i += nlit;
            if (k == 32 * BPL) continue;
            const uint32_t c = p[i + 2];
            if (c == 0xff) {
                if (lane < 2) out[o + lane] = 0xff;
                o += 2;
            }
            else {
                for (uint32_t j = lane; j < 4 + c; j += 32) out[o + j] = 0;
                o += 4 + c;
            }
            i += 3;
        }
        /* an expanded payload larger than the raw image is refused by the reference (QB3decode.cpp:398-404): such a
           stream keeps its place, and the general path reports it */
        if (o > ((uint64_t)a.w * a.h * a.bands << (a.dtype >> 1))) fits = false;
        if (fits) {
            for (uint32_t j = lane; j < 8; j += 32) out[o + j] = 0; /* the readers look a word past the end */
            for (uint32_t j = lane; j < hdr; j += 32) head[j] = j == 10 ? (uint8_t)(info.mode - 2) : stream[j];
            noff = (unsigned long long)(head - a.streams); /* may wrap: added back to a.streams modulo 2^64 */
            nlen = hdr + o;
        }
    }
    if (lane == 0) {
        offs2[tile] = noff;
        lens2[tile] = nlen;
    }
}

/*
 * decode_kernel's launch geometry. Streams per CTA: the batch spread over the SMs (a stream costs its scanner lane the
 * same time whether the warp holds one stream or 32, so few streams are spread thin and many are packed); units of
 * about 128 groups in whole rebuild iterations; three unit slots between the scanner and the rebuild warps.
 */
template <typename T> static cudaError_t launch_fused(const DecArgs &a, cudaStream_t st, uint32_t &launches)
{
    constexpr uint32_t RWORDS = FUSE_RWORDS(8 * sizeof(T)), WB = FUSE_WBYTES(8 * sizeof(T));
    int dev = 0, nsm = 0, smem_max = 0;
    cudaError_t err = cudaGetDevice(&dev);
    if (err != cudaSuccess) return err;
    cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, dev);
    cudaDeviceGetAttribute(&smem_max, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
    int smem_sm = 0;
    cudaDeviceGetAttribute(&smem_sm, cudaDevAttrMaxSharedMemoryPerMultiprocessor, dev);
    const uint32_t bands = a.bands, nbx = (a.w + 3) / 4;
    FusePlan pl = {};
    pl.nwarps = 12; /* the scanner, two warps that leave its scheduler alone, the feeder, eight that rebuild */
    pl.rwarps = 8;
    pl.nu = 3;
    pl.sel_or = 0x4440;
    pl.bpi = bands <= 32 ? 32 / bands : 1;
    pl.gpi = bands <= 32 ? pl.bpi * bands : 32;
    /* a unit is whole rebuild iterations, and its rows whole 16 byte units so that they leave as 16 byte vectors */
    uint32_t step = 1;
    {
        uint32_t unit = 16;
        const uint32_t bb = 4 * bands * (uint32_t)sizeof(T);
        while (bb % unit) unit >>= 1;
        const uint32_t al = 16 / unit, it = bands <= 32 ? pl.bpi : 1;
        step = it;
        while (step % al) step += it; /* least common multiple */
    }
    auto whole = [&](uint32_t ub) { return (ub + step - 1) / step * step; };
    uint32_t ub = 128 / bands;
    ub = ub < step ? step : ub / step * step;
    if (ub > nbx) ub = nbx;
    pl.upr = (nbx + ub - 1) / ub;
    pl.ub = whole((nbx + pl.upr - 1) / pl.upr); /* the block row cut evenly */
    if (pl.ub > nbx) pl.ub = nbx;
    pl.upr = (nbx + pl.ub - 1) / pl.ub;
    pl.rec_stride = (2 + pl.ub * bands) | 1; /* odd: the scanner's lanes write one record each, a stride apart */
    pl.rowpitch = (pl.ub * 4 * bands * (uint32_t)sizeof(T) + 15) & ~15u;
    auto layout = [&](uint32_t spc) -> size_t {
        auto up = [](size_t v) { return (v + 15) & ~(size_t)15; };
        size_t off = up((size_t)32 * (RWORDS + 4) * 4);
        pl.off_band = (uint32_t)off; off = up(off + (size_t)32 * bands * (WB + 1));
        pl.off_rec = (uint32_t)off;  off = up(off + (size_t)pl.nu * (spc + 1) * pl.rec_stride * 4);
        pl.off_info = (uint32_t)off; off = up(off + (size_t)spc * sizeof(FuseStream));
        pl.off_cb = (uint32_t)off;   off = up(off + (size_t)spc * bands);
        pl.off_carry = (uint32_t)off; off = up(off + (size_t)2 * spc * bands * WB);
        pl.off_stage = (uint32_t)off; off = up(off + (size_t)pl.rwarps * 4 * pl.rowpitch);
        pl.off_tbl = (uint32_t)off;  off = up(off + 1024 + 2048 + 1024); /* alignment slack, value table, the two switch tables */
        pl.off_bar = (uint32_t)off;  off = up(off + (size_t)2 * pl.nu * 8);
        return off;
    };
    uint32_t spc = (a.ntiles + (uint32_t)nsm - 1) / (uint32_t)nsm;
    /* Several batches in flight at once (the host pipeline): a batch takes as long on few SMs as on all of them -- its
       time is one stream's serial parse -- so it is packed onto as few as possible and leaves the rest to the others. */
    if (a.shared_sm) spc = 32;
    if (spc > 32) spc = 32;
    if (spc < 1) spc = 1;
    size_t smem = layout(spc);
    while (smem > (size_t)smem_max && spc > 1) smem = layout(--spc);
    if (smem > (size_t)smem_max) return cudaErrorInvalidConfiguration;
    pl.nsm = (uint32_t)nsm;
    /* more streams than one CTA per SM takes: the build that fits two to an SM, when the shared memory does too, and
       the streams spread evenly over whole waves of two CTAs per SM */
    const bool dense = (a.ntiles + spc - 1) / spc > (uint32_t)nsm && 2 * (smem + 1024) <= (size_t)smem_sm;
    if (dense) {
        const uint32_t slots = 2 * (uint32_t)nsm, waves = (a.ntiles + 32 * slots - 1) / (32 * slots);
        spc = (a.ntiles + slots * waves - 1) / (slots * waves);
        smem = layout(spc);
    }
    pl.spc = spc;
    /* Streams are dealt to the rebuild warps for good (a stream's running values pass from unit to unit inside one warp),
       so a unit takes as long as the warp with the most streams: where eleven warps have a stream less than eight,
       and the rebuild is what the CTA would wait for, the build with sixteen warps runs */
    bool wide = !dense && spc > 16 && (spc + 10) / 11 < (spc + 7) / 8;
    if (wide) {
        pl.nwarps = 16;
        pl.rwarps = 11;
        smem = layout(spc);
        if (smem > (size_t)smem_max) { wide = false; pl.nwarps = 12; pl.rwarps = 8; smem = layout(spc); }
    }
    err = dense ? allow_max_smem<decode_kernel<T, true>>() : wide ? allow_max_smem<decode_kernel<T, false, true>>()
                                                                  : allow_max_smem<decode_kernel<T, false>>();
    if (err != cudaSuccess) return err;
    if (dense) decode_kernel<T, true><<<(a.ntiles + spc - 1) / spc, 32 * pl.nwarps, smem, st>>>(a, pl);
    else if (wide) decode_kernel<T, false, true><<<(a.ntiles + spc - 1) / spc, 32 * pl.nwarps, smem, st>>>(a, pl);
    else decode_kernel<T, false><<<(a.ntiles + spc - 1) / spc, 32 * pl.nwarps, smem, st>>>(a, pl);
    launches += 1;
    return cudaGetLastError();
}


/* kernels launched, for the bookkeeping of qb3cu_kernel_launches */
template <typename T> static cudaError_t launch_decode_t(const DecArgs &a0, cudaStream_t st, uint32_t &launches)
{
    typedef typename traits<T>::W W;
    cudaError_t err;
    bool walked = false;
    launches = 0;
    DecArgs a = a0;
    uint8_t *xscratch = nullptr;
    if (a.rle_hint && a.w >= 4 && a.h >= 4) {
        /* room for every stream expanded: what qb3_max_encoded_size allows, the headers, alignment */
        const uint64_t xslot = ((1024 + (uint64_t)16 * ((a.w + 3) / 4) * ((a.h + 3) / 4) * a.bands * (sizeof(T) * 8 + 2) / 8 + 2048 + 1024) + 15) & ~15ull;
        const size_t idx_bytes = ((size_t)a.ntiles * 16 + 15) & ~(size_t)15;
        cudaMemPool_t pool = scratch_pool();
        err = pool ? cudaMallocFromPoolAsync(reinterpret_cast<void **>(&xscratch), idx_bytes + xslot * a.ntiles, pool, st)
                   : cudaMallocAsync(reinterpret_cast<void **>(&xscratch), idx_bytes + xslot * a.ntiles, st);
        if (err != cudaSuccess) return err;
        unsigned long long *offs2 = reinterpret_cast<unsigned long long *>(xscratch), *lens2 = offs2 + a.ntiles;
        derle_kernel<<<(a.ntiles + 3) / 4, 128, 0, st>>>(a, xscratch + idx_bytes, xslot, offs2, lens2);
        err = cudaGetLastError();
        if (err != cudaSuccess) { cudaFreeAsync(xscratch, st); return err; }
        a.offsets = offs2;
        a.lens = lens2;
        launches += 1;
    }
    struct FreeLater { /* the expanded streams live until the last kernel of this call has run */
        uint8_t *p; cudaStream_t st;
        ~FreeLater() { if (p) cudaFreeAsync(p, st); }
    } free_later = {xscratch, st};
    if (a.w >= 4 && a.h >= 4) { /* one fused kernel: a scanner warp and rebuild warps per CTA */
        err = launch_fused<T>(a, st, launches);
        if (err != cudaSuccess) return err;
        walked = true;
    }
    const size_t smem = (size_t)32 * a.bands * (2 * sizeof(W) + 2);
    err = allow_max_smem<parse_kernel<T>>();
    if (err != cudaSuccess) return err;
    parse_kernel<T><<<(a.ntiles + 31) / 32, 32, smem, st>>>(a, walked);
    err = cudaGetLastError();
    if (err != cudaSuccess) return err;
    const uint32_t rows_per_cta = 16;
    dim3 grid(a.ntiles, (a.h + rows_per_cta - 1) / rows_per_cta);
    finish_kernel<T><<<grid, 256, 0, st>>>(a, rows_per_cta);
    err = cudaGetLastError();
    if (err != cudaSuccess) return err;
    seal_kernel<<<(a.ntiles + 255) / 256, 256, 0, st>>>(a.status, a.ntiles);
    launches += 3;
    return cudaGetLastError();
}

cudaError_t launch_decode(const DecArgs &a, uint32_t tsize, cudaStream_t st, uint32_t &launches)
{
    switch (tsize) {
    case 1: return launch_decode_t<uint8_t>(a, st, launches);
    case 2: return launch_decode_t<uint16_t>(a, st, launches);
    case 4: return launch_decode_t<uint32_t>(a, st, launches);
    default: return launch_decode_t<uint64_t>(a, st, launches);
    }
}

} // namespace qb3
