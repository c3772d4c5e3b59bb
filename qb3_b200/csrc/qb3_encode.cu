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
template <typename T>
__device__ __forceinline__ void stage_rows(const EncArgs &a, const uint8_t *src, uint8_t *stage,
                                           uint32_t y0, uint32_t xs, uint32_t npix)
{
    constexpr int BITS = traits<T>::BITS;
    const uint32_t rowvals = npix * a.bands;
    if (a.vec_stage) {
        const uint32_t rowbytes = rowvals * (uint32_t)sizeof(T), upr = a.rowpitch >> 4;
        for (uint32_t idx = threadIdx.x; idx < 4 * upr; idx += blockDim.x) {
            const uint32_t r = idx / upr, u = idx - r * upr;
            const uint8_t *g = src + ((uint64_t)(y0 + r) * a.stride + (uint64_t)xs * a.bands) * sizeof(T);
            const uint32_t mis = (uint32_t)((uintptr_t)g & 15);
            if (u >= ((mis + rowbytes + 15) >> 4)) continue;
            const uint8_t *ga = g - mis + 16 * (size_t)u;
            uint8_t *sa = stage + r * a.rowpitch + 16 * u;
            if (ga >= g && ga + 16 <= g + rowbytes)
                *reinterpret_cast<uint4 *>(sa) = ld_stream16(ga);
            else
                for (int b = 0; b < 16; b++)
                    if (ga + b >= g && ga + b < g + rowbytes) sa[b] = ga[b];
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

template <typename T, bool BEST>
__global__ void __launch_bounds__(512) encode_kernel(const __grid_constant__ EncArgs a)
{
    typedef typename traits<T>::W W;
    constexpr int BITS = traits<T>::BITS, U = traits<T>::U;
    constexpr uint32_t UMASK = (1u << U) - 1;

    extern __shared__ __align__(16) uint8_t smem[];
    uint32_t *win = reinterpret_cast<uint32_t *>(smem);
    uint8_t *stage = smem + (size_t)a.win_words * 4;
    unsigned long long *carry_prev = reinterpret_cast<unsigned long long *>(stage + 4 * (size_t)a.rowpitch); /* [2][bands] */
    uint32_t *scan_scratch = reinterpret_cast<uint32_t *>(carry_prev + 2 * a.bands);                   /* [33] */
    uint8_t *carry_rung = reinterpret_cast<uint8_t *>(scan_scratch + 36);                              /* [2][bands] */
    uint8_t *rung_s = carry_rung + 2 * a.bands;                                                        /* [blockDim] */

    const uint32_t tid = threadIdx.x, NT = blockDim.x, tile = blockIdx.x;
    const uint8_t *src = a.src + (uint64_t)tile * a.src_pitch;
    uint8_t *dst = a.dst + (uint64_t)tile * a.slot;
    const bool use_step = a.mode != M_FTL;

    /* running state in, zero unless the caller keeps it across calls (reference: QB3encode.h:391-394) */
    for (uint32_t c = tid; c < a.bands; c += NT) {
        const unsigned long long *st = a.state ? a.state + (uint64_t)tile * 3 * a.bands : nullptr;
        carry_prev[c] = st ? st[c] : 0ull;
        carry_rung[c] = st ? (uint8_t)st[a.bands + c] : (uint8_t)0;
    }
    for (uint32_t i = tid; i < a.win_words; i += NT) win[i] = 0;
    __syncthreads();
    for (uint32_t i = tid; i < a.hdr_len; i += NT) reinterpret_cast<uint8_t *>(win)[i] = a.hdr[i];

    uint32_t wbits = a.hdr_len * 8;    /* bits waiting in the window */
    uint64_t flushed = 0;              /* 16 byte units already in global memory */
    bool overflow = false;             /* output would not fit the slot: the tile ends up stored */
    uint32_t it = 0;

    /* a.small == 3: sixteen pixels or fewer are stored outright (reference: QB3encode.cpp:490-491) */
    for (uint32_t by = 0; by < a.nby && a.small != 3; by++) {
        const uint32_t y0 = min(4 * by, a.vh - 4);
        for (uint32_t sg = 0; sg < a.segs; sg++, it++) {
            const uint32_t bx0 = sg * a.seg_blocks, nblk = min(a.seg_blocks, a.nbx - bx0), ng = nblk * a.bands;
            const uint32_t xs = min(4 * bx0, a.vw - 4), xe = min(4 * (bx0 + nblk), a.vw);
            stage_rows<T>(a, src, stage, y0, xs, xe - xs);
            __syncthreads(); /* also orders the header / carry writes before their first use */

            const bool active = tid < ng;
            const uint32_t blk = tid / a.bands, c = tid - blk * a.bands;
            const uint32_t par = it & 1;
            W m[16];
            W bitsused = 0;
            uint32_t rung = 0;
            if (active) {
                const uint32_t bx = bx0 + blk, x0 = min(4 * bx, a.vw - 4), cb = a.cband[c];
                uint32_t rowoff[4];
#pragma unroll
                for (int r = 0; r < 4; r++) {
                    uint32_t mis = 0;
                    if (a.vec_stage)
                        mis = (uint32_t)((uintptr_t)(src + ((uint64_t)(y0 + r) * a.stride + (uint64_t)xs * a.bands) * sizeof(T)) & 15);
                    rowoff[r] = r * a.rowpitch + mis;
                }
                const W TM = (W)lowmask64(BITS);
                W prv;
                if (blk > 0) { /* last value of the previous block: curve position 15 */
                    const uint32_t n15 = (uint32_t)a.order & 15;
                    const T *p = reinterpret_cast<const T *>(stage + rowoff[n15 >> 2]) + (size_t)(4 * (bx - 1) - xs + (n15 & 3)) * a.bands;
                    prv = (W)p[c];
                    if (cb != c) prv = (prv - (W)p[cb]) & TM;
                }
                else prv = (W)carry_prev[par * a.bands + c] & TM;
#pragma unroll
                for (int i = 0; i < 16; i++) {
                    const uint32_t n = (uint32_t)(a.order >> (4 * (15 - i))) & 15;
                    const T *p = reinterpret_cast<const T *>(stage + rowoff[n >> 2]) + (size_t)(x0 - xs + (n & 3)) * a.bands;
                    W v = (W)p[c];
                    if (cb != c) v = (v - (W)p[cb]) & TM;
                    m[i] = mags<BITS, W>(v - prv);
                    prv = v;
                    bitsused |= m[i];
                }
                rung = topbit((W)(bitsused | 1));
                rung_s[tid] = (uint8_t)rung;
                if (blk == nblk - 1) { /* becomes the neighbour of the next segment's first block */
                    carry_prev[(par ^ 1) * a.bands + c] = (unsigned long long)prv;
                    carry_rung[(par ^ 1) * a.bands + c] = (uint8_t)rung;
                }
            }
            __syncthreads();

            uint32_t len = 0, cs = 0;
            if (active) {
                const uint32_t oldrung = blk > 0 ? rung_s[tid - a.bands] : carry_rung[par * a.bands + c];
                cs = switch_entry<U>(rung, oldrung);
                len = cs >> 12;
                if (bitsused <= 1) len += 1 + (bitsused ? 16 : 0); /* reference: QB3encode.h:159-166 */
                else {
                    prepare_group<W>(m, rung, use_step);
#pragma unroll
                    for (int i = 0; i < 16; i++) len += code_len<W>(m[i], rung);
                }
            }
            uint32_t total;
            const uint32_t off = block_exclusive_scan(len, scan_scratch, total);

            const uint32_t s = wbits + off, e = s + len;
            Packer pk;
            pk.start(win, s);
            if (active) {
                pk.put32(cs & 0xfff, cs >> 12);
                if (bitsused <= 1) {
                    uint32_t b = (uint32_t)bitsused;
                    if (bitsused) {
#pragma unroll
                        for (int i = 0; i < 16; i++) b |= (uint32_t)m[i] << (i + 1);
                    }
                    pk.put32(b, bitsused ? 17 : 1);
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
                if ((flushed + k + 1) * 16 <= a.slot)
                    st_stream16(dst + (flushed + k) * 16, reinterpret_cast<const uint4 *>(win)[k]);
            }
            if ((flushed + nu) * 16 > a.slot) overflow = true;
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

    /* running state out (reference: QB3encode.h:446-449) */
    if (a.state) {
        unsigned long long *st = a.state + (uint64_t)tile * 3 * a.bands;
        for (uint32_t c = tid; c < a.bands; c += NT) {
            st[c] = carry_prev[(it & 1) * a.bands + c];
            st[a.bands + c] = carry_rung[(it & 1) * a.bands + c];
        }
    }

    uint64_t len_bytes = flushed * 16 + ((wbits + 7) >> 3);
    if ((flushed + 1) * 16 > a.slot) overflow = true;
    if (wbits && !overflow && tid == 0)
        st_stream16(dst + flushed * 16, reinterpret_cast<const uint4 *>(win)[0]);

    /* stored fallback when coding did not shrink the tile (reference: QB3encode.cpp:570-573, 461-485) */
    if (a.small == 3 || overflow || a.raw_size <= len_bytes) {
        __syncthreads();
        for (uint32_t i = tid; i < a.hdr_stored_len; i += NT) dst[i] = a.hdr_stored[i];
        const uint64_t line = (uint64_t)a.w * a.bands * sizeof(T), pitch = a.stride * sizeof(T);
        for (uint64_t i = tid; i < a.raw_size; i += NT) {
            const uint64_t y = i / line, x = i - y * line;
            dst[a.hdr_stored_len + i] = src[y * pitch + x];
        }
        len_bytes = a.hdr_stored_len + a.raw_size;
    }
    if (tid == 0) {
        a.sizes[tile] = len_bytes;
        if (a.status) a.status[tile] = 0;
    }
}

/* ------------------------------------------------------------------ launch */

template <typename T> static cudaError_t launch_encode_t(const EncArgs &a, size_t ntiles, uint32_t threads, size_t smem, cudaStream_t st)
{
    const bool best = a.mode == M_CF_Z || a.mode == M_CF_H;
    auto kern = best ? encode_kernel<T, true> : encode_kernel<T, false>;
    cudaError_t err = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (err != cudaSuccess) return err;
    kern<<<(unsigned)ntiles, threads, smem, st>>>(a);
    return cudaGetLastError();
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
