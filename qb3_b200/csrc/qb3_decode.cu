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


/*
 * Register bit buffer for the fast path (8 and 16 bit types, no RLE): 64 bits of look-ahead fed by aligned 32 bit
 * loads with one word of read-ahead, so the load latency stays off the parse chain. After refill() at least 33 bits
 * are valid, which covers every field of these types (a code is at most 17 bits). Zero fill past the end.
 */
struct FastBits {
    const uint32_t *base;
    uint64_t buf;
    uint32_t nwords, k, tailmask, mis8, nb, nxt;

    __device__ __forceinline__ uint32_t load(uint32_t i) const
    {
        uint32_t w = i < nwords ? __ldg(base + i) : 0u;
        if (i + 1 == nwords) w &= tailmask;
        return w;
    }
    __device__ __forceinline__ void open(const uint8_t *p, uint64_t len)
    {
        const uint32_t mis = (uint32_t)((uintptr_t)p & 3);
        const uint64_t span = mis + len;
        const uint32_t tail = (uint32_t)span & 3;
        base = reinterpret_cast<const uint32_t *>(p - mis);
        nwords = (uint32_t)((span + 3) >> 2);
        tailmask = tail ? (1u << (8 * tail)) - 1 : 0xffffffffu;
        mis8 = 8 * mis;
        buf = (uint64_t)(load(0) >> mis8);
        nb = 32 - mis8;
        buf |= (uint64_t)load(1) << nb;
        nb += 32;
        nxt = load(2);
        k = 3;
    }
    __device__ __forceinline__ void refill()
    {
        if (nb <= 32) {
            buf |= (uint64_t)nxt << nb;
            nb += 32;
            nxt = load(k);
            k++;
        }
    }
    __device__ __forceinline__ uint64_t peek() { refill(); return buf; }
    __device__ __forceinline__ void advance(uint32_t n) { buf >>= n; nb -= n; } /* n <= 33, after peek() / refill() */
    __device__ __forceinline__ uint64_t get(uint32_t n)
    {
        refill();
        const uint64_t v = buf & lowmask64(n);
        advance(n);
        return v;
    }
    __device__ __forceinline__ uint64_t consumed() const { return 32ull * (k - 1) - mis8 - nb; }
};

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

/*
 * Fast path: 8 and 16 bit types, no RLE, regular geometry. Same parse as the general path, but
 *  - bits come from the register buffer above, values are decoded with rung-uniform 32 bit arithmetic
 *    (no tables, no rung branches inside the 16 value loop)
 *  - pixels are not scattered to global memory one by one: each lane owns four staged rows of a few blocks
 *    in shared memory (odd word stride between lanes, so no bank conflicts) and writes them out as whole
 *    words when the staging group is full. Groups wider than the staging budget are written directly.
 */
template <typename T, bool STAGED>
__device__ bool decode_fast(const DecArgs &a, const StreamInfo &info, const uint8_t *payload, uint64_t plen, T *out,
                            uint32_t *prev, uint32_t *pcf, uint8_t *runbits, uint8_t *stage, uint32_t lane_stride,
                            uint32_t stage_blocks)
{
    typedef uint32_t W;
    constexpr int BITS = traits<T>::BITS, U = traits<T>::U;
    constexpr uint32_t UMASK = (1u << U) - 1, LMASK = 2 * UMASK + 1;
    constexpr W TM = (W)((1ull << BITS) - 1);
    const uint32_t lane = threadIdx.x;
    const uint64_t order = info.order ? info.order : HILBERT;
    const bool ftl = info.mode == M_FTL;
    const uint32_t nbx = (a.w + 3) / 4, nby = (a.h + 3) / 4, bands = a.bands;

    FastBits s;
    s.open(payload, plen);

    /* element offsets of the 16 curve positions inside the destination of a block */
    const uint32_t rowelems = STAGED ? stage_blocks * 4 * bands : (uint32_t)a.stride;
    uint32_t off[16];
#pragma unroll
    for (int i = 0; i < 16; i++) {
        const uint32_t n = (uint32_t)(order >> (4 * (15 - i))) & 15;
        off[i] = (n >> 2) * rowelems + (n & 3) * bands;
    }
    T *const lane_stage = reinterpret_cast<T *>(stage + (size_t)lane * lane_stride);
    for (uint32_t c = 0; c < bands; c++) { prev[c * 32 + lane] = 0; pcf[c * 32 + lane] = 0; runbits[c * 32 + lane] = 0; }

    bool failed = false;
    for (uint32_t by = 0; by < nby && !failed; by++) {
        const uint32_t y0 = min(4 * by, a.h - 4);
        for (uint32_t gb = 0; gb < nbx && !failed; gb += stage_blocks) {
            const uint32_t gend = min(nbx, gb + stage_blocks);
            const uint32_t xs = min(4 * gb, a.w - 4), xe = min(4 * gend, a.w);
            for (uint32_t bx = gb; bx < gend && !failed; bx++) {
                const uint32_t x0 = min(4 * bx, a.w - 4);
                for (uint32_t c = 0; c < bands; c++) {
                    W g[16];
                    uint32_t cs = 0;
                    s.refill();
                    if (s.buf & 1) {
                        cs = ds_entry(U, (uint32_t)(s.buf >> 1) & LMASK);
                        s.advance(cs >> 12);
                    }
                    else s.advance(1);
                    if (ftl || (cs & 0xfff) != 0 || cs == 0) {
                        const uint32_t r = (runbits[c * 32 + lane] + cs) & UMASK;
                        runbits[c * 32 + lane] = (uint8_t)r;
                        if (r == 0) { /* flag, then 16 raw bits (reference: QB3decode.h:148-160) */
                            s.refill();
                            const uint32_t x = (uint32_t)s.buf;
                            const uint32_t b = (x & 1) ? (x >> 1) & 0xffffu : 0u;
                            s.advance((x & 1) ? 17 : 1);
#pragma unroll
                            for (int i = 0; i < 16; i++) g[i] = (b >> i) & 1;
                        }
                        else {
                            const uint32_t half = 1u << (r - 1), fm1 = 2 * half - 1, sm = r < 8 ? 4 * half - 1 : 0;
                            uint32_t M = 0;
#pragma unroll
                            for (int i = 0; i < 16; i++) {
                                s.refill();
                                const uint32_t x = (uint32_t)s.buf;
                                const uint32_t b0 = x & 1, t = b0 & (x >> 1);
                                const uint32_t ht = half << t;
                                uint32_t v = ((x >> (1 + b0)) & (ht - 1)) | ((half & (0u - b0)) << t);
                                s.advance(r + b0 + t);
                                if (v - fm1 <= 1u) v ^= sm; /* middle swap at rungs 1..7 */
                                g[i] = v;
                                M |= ((v >> r) & 1u) << i;
                            }
                            if (!ftl) {
                                const int k = step_decode_index(M);
#pragma unroll
                                for (int i = 0; i < 16; i++) if (i == k) g[i] ^= 1u << r;
                            }
                        }
                    }
                    else { /* rare: keep the by-reference array of the out-of-line call away from the hot registers */
                        W sg[16];
                        if (read_special_group<W, BITS, U>(s, sg, runbits[c * 32 + lane], pcf[c * 32 + lane])) { failed = true; break; }
#pragma unroll
                        for (int i = 0; i < 16; i++) g[i] = sg[i];
                    }

                    W prv = prev[c * 32 + lane];
                    T *dstp = STAGED ? lane_stage + (size_t)(x0 - xs) * bands + c
                                     : out + (uint64_t)y0 * a.stride + (uint64_t)x0 * bands + c;
#pragma unroll
                    for (int i = 0; i < 16; i++) {
                        prv = (prv + smag<BITS, W>(g[i])) & TM;
                        dstp[off[i]] = (T)prv;
                    }
                    prev[c * 32 + lane] = prv;
                }
            }
            if (STAGED && !failed) { /* staged rows leave as whole words when the destination allows */
                const uint32_t rowbytes = (xe - xs) * bands * (uint32_t)sizeof(T), srow = rowelems * (uint32_t)sizeof(T);
                for (uint32_t r = 0; r < 4; r++) {
                    uint8_t *gp = reinterpret_cast<uint8_t *>(out + (uint64_t)(y0 + r) * a.stride + (uint64_t)xs * bands);
                    const uint8_t *sp = reinterpret_cast<const uint8_t *>(lane_stage) + r * srow;
                    if ((((uintptr_t)gp | rowbytes) & 15) == 0) {
                        for (uint32_t j = 0; j < rowbytes; j += 16) {
                            const uint32_t *w = reinterpret_cast<const uint32_t *>(sp + j);
                            *reinterpret_cast<uint4 *>(gp + j) = make_uint4(w[0], w[1], w[2], w[3]);
                        }
                    }
                    else if ((((uintptr_t)gp | rowbytes) & 3) == 0) {
                        for (uint32_t j = 0; j < rowbytes; j += 4)
                            *reinterpret_cast<uint32_t *>(gp + j) = *reinterpret_cast<const uint32_t *>(sp + j);
                    }
                    else {
                        for (uint32_t j = 0; j < rowbytes; j += sizeof(T))
                            *reinterpret_cast<T *>(gp + j) = *reinterpret_cast<const T *>(sp + j);
                    }
                }
            }
        }
    }
    if (failed) return true;
    const uint64_t total = 8 * plen, used = s.consumed();
    return total > used && total - used > 7; /* reference: QB3decode.h:411,740 */
}

template <typename T>
__global__ void __launch_bounds__(32, 1) parse_kernel(const DecArgs a, uint32_t stage_off, uint32_t lane_stride, uint32_t stage_blocks)
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
        a.status[tile] = plen == raw ? QB3CU_TILE_OK : QB3CU_TILE_CORRUPT; /* finish_kernel copies the pixels */
        return;
    }
    if ((uint64_t)a.w * a.h < 16) { a.status[tile] = QB3CU_TILE_CORRUPT; return; } /* reference: QB3decode.cpp:389 */

    const bool rle = info.mode == 2 || info.mode == 3 || info.mode == 6 || info.mode == 7;
    uint64_t logical = plen;
    if (rle) {
        logical = derle_size(payload, plen);
        if (logical > raw) { a.status[tile] = QB3CU_TILE_RLE_TOO_BIG; return; } /* reference: QB3decode.cpp:401 */
    }
    if (sizeof(T) <= 2 && !rle && a.w >= 4 && a.h >= 4) {
        uint32_t *p32 = reinterpret_cast<uint32_t *>(prev), *c32 = reinterpret_cast<uint32_t *>(pcf);
        const bool bad = stage_blocks
            ? decode_fast<T, true>(a, info, payload, plen, out, p32, c32, runbits, smem + stage_off, lane_stride, stage_blocks)
            : decode_fast<T, false>(a, info, payload, plen, out, p32, c32, runbits, nullptr, 0, 1);
        a.status[tile] = bad ? QB3CU_TILE_CORRUPT : QB3CU_TILE_OK;
        return;
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
    a.status[tile] = QB3CU_TILE_OK;
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
    if (a.status[tile] != QB3CU_TILE_OK) return;
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
    const uint64_t q = info.quanta, UM = lowmask64(BITS);
    const uint64_t umax_q = UM / q;
    const long long smax = (long long)(UM >> 1), smin = -smax - 1;
    const long long smax_q = smax / (long long)q, smin_q = smin / (long long)q;
    for (uint32_t y = ybeg; y < yend; y++) {
        T *row = out + (uint64_t)y * a.stride;
        for (uint32_t x = threadIdx.x; x < a.w; x += blockDim.x) {
            T *p = row + (uint64_t)x * a.bands;
            if (derived)
                for (uint32_t c = 0; c < a.bands; c++)
                    if (cband[c] != c) p[c] = (T)(p[c] + p[cband[c]]);
            if (q > 1)
                for (uint32_t c = 0; c < a.bands; c++) {
                    const uint64_t v = (uint64_t)p[c];
                    uint64_t r;
                    if (is_signed) {
                        const long long d = (long long)(v << (64 - BITS)) >> (64 - BITS);
                        long long t = d <= smax_q ? (long long)((uint64_t)d * q) : smax;
                        if (q > 2 && d < smin_q) t = smin;
                        r = (uint64_t)t & UM;
                    }
                    else r = v <= umax_q ? v * q : UM;
                    p[c] = (T)r;
                }
        }
    }
    (void)sizeof(W);
}

/* ------------------------------------------------------------------ launch */

template <typename T> static cudaError_t launch_decode_t(const DecArgs &a, cudaStream_t st)
{
    typedef typename traits<T>::W W;
    size_t smem = (size_t)32 * a.bands * (2 * sizeof(W) + 2);
    /* fast path staging: as many blocks per lane as fit 192 bytes per staged row, none when one block is wider */
    uint32_t stage_blocks = 0, lane_stride = 0, stage_off = 0;
    const uint32_t block_row_bytes = 4 * a.bands * (uint32_t)sizeof(T);
    if (sizeof(T) <= 2 && block_row_bytes <= 192) {
        stage_blocks = 192 / block_row_bytes;
        lane_stride = 4 * stage_blocks * block_row_bytes;
        lane_stride = ((lane_stride + 3) & ~3u) | 4; /* an odd number of words: lanes fall on different banks */
        stage_off = (uint32_t)((smem + 15) & ~(size_t)15);
        smem = stage_off + (size_t)32 * lane_stride;
    }
    cudaError_t err = cudaFuncSetAttribute(parse_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (err != cudaSuccess) return err;
    parse_kernel<T><<<(a.ntiles + 31) / 32, 32, smem, st>>>(a, stage_off, lane_stride, stage_blocks);
    err = cudaGetLastError();
    if (err != cudaSuccess) return err;
    const uint32_t rows_per_cta = 16;
    dim3 grid(a.ntiles, (a.h + rows_per_cta - 1) / rows_per_cta);
    finish_kernel<T><<<grid, 256, 0, st>>>(a, rows_per_cta);
    return cudaGetLastError();
}

cudaError_t launch_decode(const DecArgs &a, uint32_t tsize, cudaStream_t st)
{
    switch (tsize) {
    case 1: return launch_decode_t<uint8_t>(a, st);
    case 2: return launch_decode_t<uint16_t>(a, st);
    case 4: return launch_decode_t<uint32_t>(a, st);
    default: return launch_decode_t<uint64_t>(a, st);
    }
}

} // namespace qb3
