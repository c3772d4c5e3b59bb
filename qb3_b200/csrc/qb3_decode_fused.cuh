/*
 * qb3_decode_fused.cuh -- decode_kernel: the whole decode of a batch of 8 or 16 bit streams in one launch.
 * Included by qb3_decode.cu after the readers and group parsers it shares with the other decode kernels.
 *
 * A CTA takes up to 32 streams.
 *
 *  - The last warp is the SCANNER: one stream per lane, the serial part of the format and nothing else (where every group
 *    starts and which rung its band was at before it, QB3decode.h:119-129, 334). What bounds a stream is this warp's
 *    instruction count per group (integer instructions issue every other cycle), so the reader is as lean as it
 *    gets: three raw stream words in registers, a 32 bit window cut from them with one funnel shift per block of
 *    values (three 9 bit codes or two 16 bit ones always fit), and inside a block one AND, one byte permute (the
 *    length from the code's two low bits) and one shift per value. Moving to the next block is an add and a few
 *    selects -- no bit buffer is kept shifted, nothing is refilled conditionally. Compressed bytes arrive by cp.async
 *    in a per lane ring many chunks ahead. Per group it stores one record, (start bit << 4) | old rung, into a ring
 *    of "units" (a run of blocks of one block row) in shared memory.
 *  - The other warps REBUILD: a warp takes one stream's unit, a lane per group: the group's bits are fetched at the
 *    recorded position (through L1 / L2, the scanner has just pulled them in), values decoded (8 bit data: one table
 *    look-up per value gives the sign-unfolded delta with the middle swap already undone; 16 bit: arithmetic),
 *    running sums inside the group, a strided shuffle scan over the band's groups, pixels scattered into four staged
 *    rows, core band added, quanta multiplied, rows stored as 16 byte vectors (QB3decode.h:603-737).
 *  Scanner and rebuild warps are coupled by two mbarriers per unit slot (full / empty); nothing goes through global
 *  memory between them, so the pass moves its algorithmic bytes once: streams in, pixels out.
 */
#ifndef QB3_B200_DECODE_FUSED_CUH
#define QB3_B200_DECODE_FUSED_CUH

#define FUSE_RWORDS(bits) ((bits) == 8 ? 256 : 512)
#define FUSE_WBYTES(bits) ((bits) == 64 ? 8 : 4) /* sizeof(traits<T>::W) */

struct FusePlan {
    uint32_t spc, rwarps;     /* streams per CTA (1..32), rebuild warps */
    uint32_t nwarps;          /* warps in the CTA: the scanner is the last, those on its scheduler idle */
    uint32_t nsm;             /* SMs of the device */
    uint32_t ub, upr, nu;     /* blocks per unit, units per block row, unit slots */
    uint32_t rec_stride;      /* words per (slot, stream): two anchor words, then a record per group */
    uint32_t rowpitch;        /* bytes between the staged rows of a rebuild warp */
    uint32_t gpi, bpi;        /* groups / blocks per rebuild iteration */
    uint32_t sel_or;          /* 0x4440, from the host so that the compiler keeps it in a register: (z & 3) | sel_or then
                                 is one LOP3 instead of two */
    uint32_t off_band, off_rec, off_info, off_cb, off_carry, off_stage, off_tbl, off_bar; /* shared memory, bytes */
};

struct FuseStream {
    const uint32_t *base;     /* 16 byte aligned, at or before the payload */
    uint8_t *out;
    uint64_t order, quanta, plen;
    uint32_t nwords, tailmask, mis, flags;
};
constexpr uint32_t FS_GO = 1, FS_FTL = 2, FS_DERIVED = 4, FS_SWEEP = 8;

__device__ __forceinline__ void mbar_init(uint32_t addr, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(addr), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t addr)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"(addr) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t addr, uint32_t parity)
{
    uint32_t ok;
    for (;;) {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok) : "r"(addr), "r"(parity) : "memory");
        if (ok) break;
        __nanosleep(1000); /* a waiting warp must not take issue slots from the scanner */
    }
}
__device__ __forceinline__ int32_t lds_s8(uint32_t addr)
{
    int32_t v;
    asm volatile("ld.shared.s8 %0, [%1];" : "=r"(v) : "r"(addr));
    return v;
}
__device__ __forceinline__ uint32_t lds_u16(uint32_t addr)
{
    uint32_t v;
    asm volatile("ld.shared.u16 %0, [%1];" : "=r"(v) : "r"(addr));
    return v;
}
__device__ __forceinline__ uint32_t lds_u8(uint32_t addr)
{
    uint32_t v;
    asm volatile("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(addr));
    return v;
}

/* 8 bit data, rebuild: the sixteen values of a block (curve order, low bytes) as its four rows, a register each */
__host__ __device__ constexpr int curve_index(uint64_t order, int n)
{
    for (int i = 0; i < 16; i++)
        if ((int)((order >> (4 * (15 - i))) & 15) == n) return i;
    return 0;
}
__device__ __forceinline__ uint32_t pack4(uint32_t a, uint32_t b, uint32_t c, uint32_t d)
{
    return __byte_perm(__byte_perm(a, b, 0x0040), __byte_perm(c, d, 0x0040), 0x5410);
}
template <uint64_t ORDER> __device__ __forceinline__ void pack_rows(const uint32_t (&v)[16], uint32_t (&R)[4])
{
#define QB3_AT(n) v[std::integral_constant<int, curve_index(ORDER, n)>::value]
    R[0] = pack4(QB3_AT(0), QB3_AT(1), QB3_AT(2), QB3_AT(3));
    R[1] = pack4(QB3_AT(4), QB3_AT(5), QB3_AT(6), QB3_AT(7));
    R[2] = pack4(QB3_AT(8), QB3_AT(9), QB3_AT(10), QB3_AT(11));
    R[3] = pack4(QB3_AT(12), QB3_AT(13), QB3_AT(14), QB3_AT(15));
#undef QB3_AT
}
__device__ __forceinline__ uint32_t add4(uint32_t a, uint32_t b) /* four byte sums, no carry between them */
{
    return ((a & 0x7f7f7f7fu) + (b & 0x7f7f7f7fu)) ^ ((a ^ b) & 0x80808080u);
}

/* A group the fast path of the rebuild does not take: a common factor or index group, or one at the very end of the
   stream. Parsed with the general reader; v gets the sign folded values. kind: 0 plain, 1 common factor group that
   reuses the band's factor (pc_in), 2 one that wrote a new factor (returned in pc_out). */
template <typename T>
__device__ __noinline__ void fuse_group_slow(const FuseStream &fs, uint64_t P, uint32_t oldrung, typename traits<T>::W pc_in,
                                             typename traits<T>::W (&v)[16], uint32_t &kind, typename traits<T>::W &pc_out)
{
    typedef typename traits<T>::W W;
    typedef typename std::conditional<(traits<T>::BITS <= 16), GroupBits, WideBits>::type Bits;
    constexpr int BITS = traits<T>::BITS, U = traits<T>::U;
    constexpr uint32_t UMASK = (1u << U) - 1, LMASK = 2 * UMASK + 1;
    const bool ftl = fs.flags & FS_FTL;
    const uint8_t *payload = reinterpret_cast<const uint8_t *>(fs.base) + fs.mis;
    Bits s;
    s.open(payload, fs.plen, P - 8 * fs.mis);
    kind = 0; pc_out = 0;
    uint32_t cs = 0;
    {
        const uint32_t x = (uint32_t)s.peek();
        if (x & 1) cs = ds_entry(U, (x >> 1) & LMASK);
        s.advance((x & 1) ? cs >> 12 : 1);
    }
    if (ftl || (cs & 0xfff) != 0 || cs == 0) {
        W g[16];
        read_group<W>(s, (oldrung + cs) & UMASK, g, !ftl);
#pragma unroll
        for (int i = 0; i < 16; i++) v[i] = g[i];
        return;
    }
    Bits p = s; /* what kind it is: the flag after the signal and the switch (QB3decode.h:624-640) */
    const uint32_t e = ds_entry(U, (uint32_t)p.peek() & LMASK);
    p.advance((e >> 12) - 1);
    if (((oldrung + e) & UMASK) != UMASK) kind = p.get(1) ? 2 : 1;
    uint8_t rbv = (uint8_t)oldrung;
    W pc = pc_in, g[16];
    read_special_group<W, BITS, U>(s, g, rbv, pc);
#pragma unroll
    for (int i = 0; i < 16; i++) v[i] = g[i];
    pc_out = pc;
}

/* DENSE: built for two CTAs to an SM (85 registers), for batches of more streams than one CTA per SM holds: the
   prologue and the tail of one CTA then pass behind the other's work, and a second scanner warp runs per SM. */
/* WIDE: sixteen warps instead of twelve (128 registers), eleven of them rebuilding: with 25 to 32 streams in the CTA every
   rebuild warp then has three streams per unit instead of four, and the scanner stops waiting for free unit slots
   (C2 decode, 28 streams per CTA: 10.6 -> 9.3 ms). */
template <typename T, bool DENSE = false, bool WIDE = false>
__global__ void __launch_bounds__(WIDE ? 512 : 384, DENSE ? 2 : 1) decode_kernel(const __grid_constant__ DecArgs a, const __grid_constant__ FusePlan pl)
{
    typedef typename traits<T>::W W;                     /* register type of a value */
    typedef typename std::make_signed<W>::type SW;
    constexpr int BITS = traits<T>::BITS, U = traits<T>::U;
    constexpr bool NARROW = BITS <= 16;          /* 8 and 16 bit data: windowed scanner, table / arithmetic rebuild */
    constexpr uint32_t UMASK = (1u << U) - 1, LMASK = 2 * UMASK + 1;
    constexpr uint32_t RB = NARROW ? 4 : 6;      /* a record is (start bit << RB) | old rung */
    const W TM = (W)lowmask64(BITS);
    constexpr int RWORDS = FUSE_RWORDS(BITS);    /* ring words per lane */
    constexpr int LSTRIDE = RWORDS + 4;          /* words between the rings of two lanes: banks shifted; the four spare
                                                    words mirror the ring's first four, so that a read of up to four
                                                    words on from any ring position never has to wrap */
    constexpr int EVERY = NARROW ? 8 : 2;        /* groups between ring upkeeps */
    constexpr int GWORDS = BITS == 8 ? 7 : BITS == 16 ? 12 : BITS == 32 ? 22 : 40; /* ring words one group of any kind can consume */
    constexpr int AHEAD = RWORDS / 4 - 2;        /* chunks kept requested beyond the one being read */
    constexpr int DRAIN = EVERY * GWORDS / 4 + 1; /* chunks an upkeep interval can consume */
    constexpr int NV0 = BITS == 8 ? 3 : 1;       /* values that share a 32 bit window with the rung switch */
    constexpr int VPB = BITS == 8 ? 3 : 2;       /* values per window after that */
    constexpr int NWL = BITS == 8 ? 5 : 9;       /* words of a group's bits, aligned; one more is loaded */
    static_assert(AHEAD >= 2 * DRAIN + 4, "ring too small for an upkeep interval");

    extern __shared__ __align__(16) uint8_t smem[];
    const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, bands = a.bands, spc = pl.spc;
    const uint32_t smem_sa = (uint32_t)__cvta_generic_to_shared(smem);
    W *pcfs = reinterpret_cast<W *>(smem + pl.off_band);                           /* scanner: [band][32] last factor */
    uint8_t *rbs = reinterpret_cast<uint8_t *>(pcfs + 32 * bands);                 /* scanner: [band][32] running rung */
    uint32_t *recs = reinterpret_cast<uint32_t *>(smem + pl.off_rec);              /* [nu][spc + 1][rec_stride] */
    FuseStream *infos = reinterpret_cast<FuseStream *>(smem + pl.off_info);        /* [spc] */
    uint8_t *cbs = smem + pl.off_cb;                                               /* [spc][bands] band maps */
    W *prevS = reinterpret_cast<W *>(smem + pl.off_carry);                         /* rebuild: [spc][bands] running value */
    W *pcfS = prevS + spc * bands;                                                 /* rebuild: [spc][bands] last factor */
    const uint32_t tbl_sa = (smem_sa + pl.off_tbl + 1023) & ~1023u;                /* value table, 1 KB aligned: deltas, then flags */
    uint8_t *tbl = smem + (tbl_sa - smem_sa);
    uint16_t *dsw = reinterpret_cast<uint16_t *>(tbl + 2048);                      /* rung switch decode table */
    uint16_t *csb = reinterpret_cast<uint16_t *>(tbl + 2048 + 2 * (2u << U));      /* the same for the scanner, see below */
    const uint32_t bar_sa = smem_sa + pl.off_bar;                                  /* full[nu], empty[nu] */

    const uint32_t nbx = (a.w + 3) / 4, nby = (a.h + 3) / 4, nunits = nby * pl.upr;

    /* ---- set-up, all threads */
    for (uint32_t i = tid; i < (2u << U); i += blockDim.x) dsw[i] = (uint16_t)ds_entry(U, i);
    /* scanner: (length << 8) | delta by the U + 2 low bits of a group, change flag included: an even index is "no
       change", one bit long (QB3decode.h:98-116, 334) */
    for (uint32_t i = tid; i < (4u << U); i += blockDim.x) {
        const uint32_t d = ds_entry(U, i >> 1);
        csb[i] = (i & 1) ? (uint16_t)(((d >> 12) << 8) | (d & 0xff)) : (uint16_t)0x100;
    }
    if (BITS == 8) {
        /* entry (4 << r) + x for the r + 2 low bits x of a code at rung r = 1..7: the sign-unfolded delta of the value
           it decodes to, middle swap undone (QB3decode.h:24-95 with QB3common.h:133), and the value's rung bit */
        for (uint32_t i = 8 + tid; i < 1024; i += blockDim.x) {
            const uint32_t r = topbit32(i) - 2, x = i - (4u << r);
            uint32_t len;
            const uint32_t v = mswap<uint32_t>((uint32_t)decode_bits(x, 0, r, len), r);
            tbl[i] = (uint8_t)smag<8, uint32_t>(v);
            tbl[1024 + i] = (uint8_t)((v >> r) & 1);
        }
    }
    for (uint32_t i = tid; i < 2 * spc * bands; i += blockDim.x) prevS[i] = 0;
    if (tid == 0) {
        for (uint32_t i = 0; i < pl.nu; i++) {
            mbar_init(bar_sa + 8 * i, 1);
            mbar_init(bar_sa + 8 * (pl.nu + i), pl.rwarps);
        }
    }

    /* scanner state that the header gives; parsed by the scanner's lanes, shared with the rebuild through infos */
    const uint32_t tile = blockIdx.x * spc + lane;
    /* The scanner is the last warp (the scheduler's arbiter prefers the highest warp id); with two CTAs to an SM the
       CTAs of odd waves take the warp before it, so that the two scanners of an SM sit on different schedulers. */
    const uint32_t scan_warp = pl.nwarps - 1 - (DENSE ? (blockIdx.x / pl.nsm) & 1 : 0);
    const bool scanner = warp == scan_warp;
    /* The FEEDER, the warp before the scanner (another scheduler): keeps the scanner's rings filled, a lane per stream.
       The scanner tells it where it reads (feed_cons, in 16 byte chunks), the feeder tells how far the ring is valid
       (feed_fill). Filling its own ring cost the scanner a seventh of its instructions. */
    const uint32_t feed_warp = scan_warp - 1;
    __shared__ uint32_t feed_cons[32], feed_fill[32], feed_done;
    if (tid < 32) { feed_cons[tid] = 0; feed_fill[tid] = 0; }
    if (tid == 0) feed_done = 0;
    const bool live = scanner && lane < spc && tile < a.ntiles;
    bool go = false, ftl_l = false;
    uint32_t mis = 0, span = 0;
    uint64_t plen = 0;
    const uint8_t *abase = a.streams;
    if (scanner && lane < spc) {
        FuseStream fs;
        fs.base = nullptr; fs.out = nullptr; fs.order = HILBERT; fs.quanta = 1; fs.plen = 0;
        fs.nwords = 0; fs.tailmask = 0; fs.mis = 0; fs.flags = 0;
        if (live) {
            const uint8_t *stream = a.streams + a.offsets[tile];
            const uint64_t slen = a.lens[tile];
            StreamInfo info;
            parse_header(stream, slen, a, info, cbs + lane * bands, 1);
            const bool rle = info.mode == 2 || info.mode == 3 || info.mode == 6 || info.mode == 7;
            /* positions are kept in 32 bits: streams of 256 MB and more go the general way */
            go = !info.bad && info.mode != M_STORED && !rle && slen < (1ull << 28);
            a.status[tile] = go ? ST_SCANNING : info.bad ? (uint32_t)QB3CU_TILE_BAD_HEADER : ST_DEFER;
            if (go) {
                const uint8_t *payload = stream + info.data_off;
                plen = slen - info.data_off;
                mis = (uint32_t)((uintptr_t)payload & 15);
                abase = payload - mis;
                span = (uint32_t)(mis + plen);
                ftl_l = info.mode == M_FTL;
                uint32_t bf = 0;
                for (uint32_t c = 0; c < bands; c++) {
                    const uint32_t k = cbs[lane * bands + c];
                    if (k != c) bf |= 1 | (cbs[lane * bands + k] != k ? 2 : 0);
                }
                /* the plain case adds the core band from the core band's lane's pixels; chained band maps, quantised
                   derived bands and blocks wider than a warp go through the reference's per pixel sweep instead */
                const bool sweep = (bf & 2) || (bf && (info.quanta > 1 || bands > 32));
                fs.base = reinterpret_cast<const uint32_t *>(abase);
                fs.out = a.dst + (uint64_t)tile * a.dst_pitch;
                fs.order = info.order ? info.order : HILBERT;
                fs.quanta = info.quanta;
                fs.plen = plen;
                fs.nwords = (span + 3) >> 2;
                fs.tailmask = (span & 3) ? (1u << (8 * (span & 3))) - 1 : 0xffffffffu;
                fs.mis = mis;
                fs.flags = FS_GO | (ftl_l ? FS_FTL : 0) | (bf ? FS_DERIVED : 0) | (sweep ? FS_SWEEP : 0);
            }
        }
        infos[lane] = fs;
    }
    __syncthreads();

    if (scanner) {
        /* ================================================================ scanner */
        const uint32_t *ring = reinterpret_cast<const uint32_t *>(smem) + lane * LSTRIDE;
        const uint32_t ring_addr = smem_sa + lane * LSTRIDE * 4;
        /* the band count in a register of its own: read from the constant bank in the loop, its latency is on every
           group's path (the compiler prefers to reload it) */
        uint32_t nbands;
        asm volatile("mov.u32 %0, %1;" : "=r"(nbands) : "r"(bands));
        /* how far the feeder has filled this lane's ring, as last seen; waited for whenever the reader could get there
           before the next look */
        volatile uint32_t *vfill = feed_fill, *vcons = feed_cons;
        uint32_t fill_seen = 0;
        auto upkeep = [&](uint32_t at_bit) {
            vcons[lane] = at_bit >> 7;
            const uint32_t need = (at_bit >> 7) + DRAIN + 2;
            while (__any_sync(0xffffffffu, (int32_t)(fill_seen - need) < 0)) {
                fill_seen = vfill[lane];
                __threadfence_block();
            }
        };
        upkeep(8 * mis);
        for (uint32_t cc = 0; cc < bands; cc++) { rbs[cc * 32 + lane] = 0; pcfs[cc * 32 + lane] = 0; }
        __syncwarp();

        /*
         * The reader. z holds the 32 stream bits from bit pos on, zh the 32 after them; wa, wb are the ring words that
         * follow the one holding bit pos, wc and wd the two after those (fetched ahead). Within a window of values only z moves
         * (plain right shifts: the parse chain is AND, permute, shift per value and nothing else); the window's last
         * shift is a funnel shift from a copy of (z, zh) moved along beside the chain, which lands on the next window's
         * z directly. The next zh is then cut afresh from the raw words -- off the chain, it is first needed a window
         * later.
         */
        uint32_t pos = 8 * mis;
        uint32_t wa = ring[(mis >> 2) + 1], wb = ring[(mis >> 2) + 2], wc = ring[(mis >> 2) + 3], wd = ring[(mis >> 2) + 4];
        uint32_t z = __funnelshift_r(ring[mis >> 2], wa, pos), zh = __funnelshift_r(wa, wb, pos);
        const uint32_t c4440 = pl.sel_or;
        /* the length of a code from its two low bits: a byte table in a register read with one permute */
        auto code_len = [&](uint32_t lens, uint32_t zz) { return __byte_perm(lens, 0u, (zz & 3) | c4440); };
        /* ends a window: (zs, zhs) are z and zh as the window found them, moved on by before (what all fields of the
           window but the last took), last is the last field's length. before + last <= 32. */
        auto next_window = [&](uint32_t zs, uint32_t zhs, uint32_t before, uint32_t last) {
            const uint32_t z1 = __funnelshift_r(zs, zhs, before), zh1 = __funnelshift_r(zhs, 0u, before);
            z = __funnelshift_rc(z1, zh1, last);           /* the chain's only instruction here */
            const uint32_t npos = pos + before + last;
            const bool cross = ((npos ^ pos) & 32u) != 0;   /* at most one word is left behind */
            wa = cross ? wb : wa;
            wb = cross ? wc : wb;
            wc = cross ? wd : wc;
            /* the second word after wb, crossed or not: asked for two words ahead, its latency has two windows to pass */
            wd = lds32(ring_addr + ((npos >> 3) & (4 * RWORDS - 4)) + 16);
            zh = __funnelshift_r(wa, wb, npos);
            pos = npos;
        };

        bool failed = false;
        uint64_t abs_bits = 0; /* position of the unit's anchor, from abase */
        uint32_t apos = 0;
        uint32_t c = 0;
        const uint32_t csb_sa = tbl_sa + 2048 + 2 * (2u << U);
        const uint32_t rlane = lane < spc ? lane : spc; /* lanes without a stream write their records to a spare row */
        /* RARE: with the groups that are parsed again after the walk -- common factor and index groups, and the 16 bit
           groups at rung 15. A batch of FTL streams of 8 bit data has neither, and the loop without the check saves
           a branch with its reconvergence point per group. */
        uint32_t e_next = lds_u16(csb_sa + 2 * (z & ((4u << U) - 1))), rung_next = 0;
        auto walk = [&](auto rare_tag) {
            constexpr bool RARE = decltype(rare_tag)::value;
            for (uint32_t u = 0; u < nunits; u++) {
                const uint32_t slot = u % pl.nu, j = u % pl.upr;
                const uint32_t b0 = j * pl.ub, ng = min(pl.ub, nbx - b0) * bands;
                if (u >= pl.nu) mbar_wait(bar_sa + 8 * (pl.nu + slot), ((u / pl.nu) - 1) & 1);
                uint32_t *rp = recs + (size_t)(slot * (spc + 1) + rlane) * pl.rec_stride;
                abs_bits += (uint32_t)(pos - apos);
                apos = pos;
                rp[0] = (uint32_t)abs_bits;
                rp[1] = (uint32_t)(abs_bits >> 32);
                rp += 2;
                for (uint32_t g0 = 0; g0 < ng; g0 += EVERY) {
                    /* ring upkeep every EVERY groups: the feeder hears where the reader is, and the reader makes sure the
                       ring is valid as far as it can get before the next upkeep */
                    upkeep(pos);
                    const uint32_t gn = min((uint32_t)EVERY, ng - g0);
                    /* two groups to a turn of the loop where it was measured to pay (8 bit FTL, the twelve warp build:
                       8.8 -> 8.5 ms; every other build lost by it) */
                    constexpr int TURN = (BITS == 8 && !RARE && !WIDE && !DENSE) ? 2 : 1;
    #pragma unroll(TURN)
                    for (uint32_t gi = 0; gi < gn; gi++) {
                        const uint32_t oldrung = rung_next, e = e_next;
                        const uint32_t gpos = pos;
                        /* the old rung of the next group's band, asked for a whole group ahead; with one band it is this
                           group's own rung */
                        const uint32_t cnext = c + 1 == nbands ? 0 : c + 1;
                        const uint32_t rung_ahead = rbs[cnext * 32 + lane];
                        const uint32_t swl = e >> 8, delta = e & 0xff;
                        /* a signal (a change flag with delta 0, QB3decode.h:619) opens a common factor or index group: the
                           walk below then runs on meaningless lengths, harmlessly, and the group is parsed again after it */
                        const bool special = RARE && !ftl_l && delta == 0 && swl != 1;
                        uint32_t r = (oldrung + delta) & UMASK;
                        /* Lengths by the code's two low bits, x0 -> r, 01 -> r + 1, 11 -> r + 2 (QB3decode.h:119-129). At rung 0
                           all sixteen are zero, so that every lane walks the same code, and the group's flag with its sixteen
                           raw bits (QB3decode.h:148-160) goes as the first value's length: 1 or 17 by the flag. Products, not
                           selects: a predicate takes three times as long to arrive as a register. */
                        const uint32_t nz = min(r, 1u);
                        const uint32_t lens = nz * (0x02000100u + (NARROW ? r : 0u) * 0x01010101u);
                        const uint32_t lens_first = (nz ^ 1) * 0x11011101u + lens;
                        if constexpr (!NARROW) {
                            /* 32 and 64 bit data: a code can be 65 bits, so no window is moved along. Value i starts at
                               s + i * r + o_i with o_i <= 2 i, the bits it is longer than r by summed over the values before
                               it: the 32 bits from s + i * r on hold its two low bits wherever it starts, and they are
                               fetched for all sixteen values at once, ahead of the chain, which is then a shift, an AND, a
                               permute and an add per value. lens is the table of the extra bits here (0, 1, 0, 2). */
                            uint32_t b = gpos + swl, o = 0;
    #pragma unroll
                            for (int i = 0; i < 16; i++) {
                                const uint32_t adr = ring_addr + ((b >> 3) & (4 * RWORDS - 4));
                                const uint32_t lo = __funnelshift_r(lds32(adr), lds32(adr + 4), b);
                                o += code_len(i ? lens : lens_first, __funnelshift_r(lo, 0u, o));
                                b += r;
                            }
                            pos = b + o;
                            const uint32_t adr = ring_addr + ((pos >> 3) & (4 * RWORDS - 4));
                            z = __funnelshift_r(lds32(adr), lds32(adr + 4), pos);
                        }
                        else {
                            const uint32_t zs = z, zhs = zh;
                            uint32_t before = swl, len;
                            z >>= swl;
    #pragma unroll
                            for (int i = 0; i < NV0; i++) {
                                len = code_len(i ? lens : lens_first, z);
                                if (i + 1 < NV0) { z >>= len; before += len; }
                            }
                            next_window(zs, zhs, before, len);
    #pragma unroll
                            for (int i0 = NV0; i0 < 16; i0 += VPB) {
                                const uint32_t zs2 = z, zhs2 = zh;
                                before = 0;
    #pragma unroll
                                for (int i = i0; i < i0 + VPB && i < 16; i++) {
                                    len = code_len(lens, z);
                                    if (i + 1 < i0 + VPB && i + 1 < 16) { z >>= len; before += len; }
                                }
                                next_window(zs2, zhs2, before, len);
                            }
                        }
                        auto reopen = [&](uint32_t at) { /* the reader afresh at a ring position */
                            pos = at;
                            const uint32_t wi = at >> 5;
                            if (!NARROW) { z = __funnelshift_r(ring[wi & (RWORDS - 1)], ring[(wi + 1) & (RWORDS - 1)], at); return; }
                            wa = ring[(wi + 1) & (RWORDS - 1)]; wb = ring[(wi + 2) & (RWORDS - 1)]; wc = ring[(wi + 3) & (RWORDS - 1)];
                        wd = ring[(wi + 4) & (RWORDS - 1)];
                            z = __funnelshift_r(ring[wi & (RWORDS - 1)], wa, at);
                            zh = __funnelshift_r(wa, wb, at);
                        };
                        /* the two rare cases behind one branch: a rarely taken branch costs its reconvergence point on every
                           group (~25 cycles), and 16 bit data had two of them */
                        if (RARE && (special || (BITS == 16 && r == 15))) {
                            if (!special) {
                                /* two 17 bit codes do not fit a window, so the walk above may have gone wrong: again, a value
                                   at a time */
                                reopen(gpos);
                                next_window(z, zh, 0, swl);
    #pragma unroll 1
                                for (int i = 0; i < 16; i++) next_window(z, zh, 0, code_len(lens, z));
                            }
                            else { /* common factor or index group: parsed in full from the ring */
                                RingBits<RWORDS> t;
                                t.ring = ring;
                                t.pos = gpos + swl;
                                W sg[16];
                                uint8_t rbv = (uint8_t)oldrung;
                                W pc = pcfs[c * 32 + lane];
                                failed |= read_special_group<W, BITS, U>(t, sg, rbv, pc);
                                pcfs[c * 32 + lane] = pc;
                                r = rbv;
                                reopen(t.pos);
                            }
                        }
                        rbs[c * 32 + lane] = (uint8_t)r;
                        c = cnext;
                        /* the next group's switch entry is asked for now: part of its latency passes behind the stores and the
                           loop's end */
                        e_next = lds_u16(csb_sa + 2 * (z & ((4u << U) - 1)));
                        rung_next = nbands == 1 ? r : rung_ahead;
                        rp[g0 + gi] = ((gpos - apos) << RB) | oldrung;
                    }
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(bar_sa + 8 * slot);
            }
        };
        if (BITS >= 16 || __any_sync(0xffffffffu, go && !ftl_l)) walk(std::true_type());
        else walk(std::false_type());
        if (lane == 0) *(volatile uint32_t *)&feed_done = 1;
        if (go) {
            const uint64_t total = 8 * plen, used = (uint32_t)(pos - 8 * mis);
            const bool bad = failed || (total > used && total - used > 7); /* reference: QB3decode.h:411,740 */
            a.status[tile] = bad ? (uint32_t)QB3CU_TILE_CORRUPT : (uint32_t)QB3CU_TILE_OK;
        }
        return;
    }

    if (warp == feed_warp) {
        /* ================================================================ feeder */
        const bool has = lane < spc && infos[lane].base != nullptr;
        const uint8_t *rsrc = has ? reinterpret_cast<const uint8_t *>(infos[lane].base) : a.streams;
        int32_t rleft = has ? (int32_t)(infos[lane].mis + infos[lane].plen) : 0; /* bytes of the stream left from rsrc */
        const uint32_t ring_addr = smem_sa + lane * LSTRIDE * 4;
        uint32_t rdst = 0, issued = 0, before = 0;
        volatile uint32_t *vfill = feed_fill, *vcons = feed_cons;
        auto request = [&](bool on) { /* the next 16 bytes of the stream, zeros beyond its end, into the ring */
            const uint32_t nbytes = (uint32_t)min(max(rleft, 0), 16);
            if (on) cp_async16_zfill(ring_addr + rdst, rsrc, nbytes);
            if (on && rdst == 0) cp_async16_zfill(ring_addr + 4 * RWORDS, rsrc, nbytes); /* the mirror of the first four words */
            if (on && rleft > 16) rsrc += 16; /* never points past the stream's last chunk */
            rleft -= on ? 16 : 0;
            rdst = (rdst + (on ? 16u : 0u)) & (4 * RWORDS - 1);
            issued += on ? 1u : 0u;
        };
        for (;;) {
            /* chunks up to AHEAD beyond the one being read; as many rounds as the lane that is furthest behind needs */
            const uint32_t want = vcons[lane] + 1 + AHEAD;
            /* not for less than eight chunks of the lane that is furthest behind (the ring is 62 ahead, the scanner asks
               for 17): the feeder shares its scheduler with rebuild warps */
            const uint32_t lag = __reduce_max_sync(0xffffffffu, want - issued), rounds = lag >= 8 ? lag : 0;
#pragma unroll 1
            for (uint32_t q = 0; q < rounds; q++) request(issued < want);
            cp_async_commit();
            cp_async_wait<1>(); /* everything but the copies just asked for has landed */
            __threadfence_block();
            vfill[lane] = before;
            before = issued;
            if (*(volatile uint32_t *)&feed_done) break;
            if (rounds == 0) __nanosleep(512);
        }
        cp_async_wait<0>();
        return;
    }

    /* ==================================================================== rebuild warps */
    /* The scanner keeps its warp scheduler to itself: the warps that would share it (same warp id modulo 4) leave, the
       others are numbered 0 .. rwarps - 1. Whatever a rebuild warp issues there is taken from the one warp the whole
       CTA waits for. (Letting those two rebuild as well when a CTA has 28 streams and the rebuild is what it waits
       for was measured: 10.7 -> 11.3 ms for 4096 tiles, the scanner loses more than the rebuild gains.) */
    if ((warp & 3) == (scan_warp & 3)) return;
    const uint32_t rw = warp - (warp + 3 - (scan_warp & 3)) / 4 - (warp > feed_warp ? 1 : 0), FULL = 0xffffffffu;
    const uint32_t rowpitch = pl.rowpitch, rowelems = rowpitch / (uint32_t)sizeof(T);
    uint8_t *stage = smem + pl.off_stage + (size_t)rw * 4 * rowpitch;
    const bool small_bands = bands <= 32;
    const uint32_t lb = small_bands ? lane / bands : 0, lc = small_bands ? lane - lb * bands : 0;
    const bool is_signed = a.dtype & 1;
    uint32_t poff[16]; /* where the 16 values of a block go in the staged rows, by the stream's scan curve */
    uint64_t poff_order = HILBERT;
#pragma unroll
    for (int i = 0; i < 16; i++) {
        const uint32_t n = (uint32_t)(HILBERT >> (4 * (15 - i))) & 15;
        poff[i] = (n >> 2) * rowelems + (n & 3) * bands;
    }
    const uint32_t c4440r = pl.sel_or;
    /* 8 bit data, up to four bands, width a multiple of four: a block's row in the staged rows is `bands` whole words. The
       lanes of a block exchange their rows (four pixels of one band each, a register) by shuffles and every lane stores
       one word per row: selectors that pick word lc of the interleaved row out of the bands' registers */
    const bool simd_geo = BITS == 8 && bands <= 4 && (a.w & 3) == 0;
    uint32_t selA = 0, selB = 0, selC = 0;
    if (simd_geo) {
        for (uint32_t t = 0; t < 4; t++) {
            const uint32_t j = 4 * lc + t, x = j / bands, k = j - x * bands;
            selA |= (k == 0 ? x : k == 1 ? 4 + x : 0u) << (4 * t);
            selB |= (k == 2 ? x : k == 3 ? 4 + x : 0u) << (4 * t);
            selC |= (k < 2 ? t : 4 + t) << (4 * t);
        }
    }

    for (uint32_t u = 0; u < nunits; u++) {
        const uint32_t slot = u % pl.nu, by = u / pl.upr, j = u - by * pl.upr;
        const uint32_t b0 = j * pl.ub, nblk = min(pl.ub, nbx - b0), ng = nblk * bands;
        const uint32_t y0 = min(4 * by, a.h - 4), xs = min(4 * b0, a.w - 4), xe = min(4 * (b0 + nblk), a.w), npx = xe - xs;
        mbar_wait(bar_sa + 8 * slot, (u / pl.nu) & 1);
        for (uint32_t s = rw; s < spc; s += pl.rwarps) {
            const FuseStream &fs = infos[s];
            if (!(fs.flags & FS_GO)) continue;
            const uint32_t *rec = recs + (size_t)(slot * (spc + 1) + s) * pl.rec_stride;
            const uint64_t anchor = (uint64_t)rec[0] | ((uint64_t)rec[1] << 32);
            const uint8_t *cb = cbs + s * bands;
            const bool ftl = fs.flags & FS_FTL, derived = fs.flags & FS_DERIVED, sweep = fs.flags & FS_SWEEP;
            const uint64_t quanta = fs.quanta;
            const bool simd_out = simd_geo && !sweep && quanta == 1 && (fs.order == HILBERT || fs.order == ZCURVE);
            if (fs.order != poff_order) { /* hardly ever: the streams of a batch come from one encoder */
                poff_order = fs.order;
#pragma unroll
                for (int i = 0; i < 16; i++) {
                    const uint32_t n = (uint32_t)(fs.order >> (4 * (15 - i))) & 15;
                    poff[i] = (n >> 2) * rowelems + (n & 3) * bands;
                }
            }
            W carry = small_bands && lane < pl.gpi ? prevS[s * bands + lc] : (W)0;
            W *pcfc = pcfS + s * bands;

            for (uint32_t it0 = 0, blk0 = 0; it0 < ng; it0 += pl.gpi, blk0 += pl.bpi) {
                const uint32_t g = it0 + lane;
                const bool active = lane < pl.gpi && g < ng;
                uint32_t blk, c;
                if (small_bands) { blk = blk0 + lb; c = lc; }
                else { blk = g / bands; c = g - blk * bands; }
                const uint32_t core = active ? cb[c] : c;

                SW d[16]; /* the group's values, sign unfolded; then their running sums */
                uint32_t kind = 0, oldrung = 0;
                W pcw = 0;
                uint64_t P = 0;
                bool slow = false;
#pragma unroll
                for (int i = 0; i < 16; i++) d[i] = 0;
                if (active) {
                    const uint32_t rc = rec[2 + g];
                    oldrung = rc & ((1u << RB) - 1);
                    P = anchor + (rc >> RB);
                    const uint32_t widx = (uint32_t)(P >> 5), sh = (uint32_t)P & 31;
                    /* the fast path is for 8 and 16 bit data and never touches the stream's last word */
                    slow = !NARROW || widx + NWL + 1 >= fs.nwords;
                    if constexpr (NARROW) if (!slow) {
                        uint32_t wv[NWL + 1];
#pragma unroll
                        for (int q = 0; q <= NWL; q++) wv[q] = __ldg(fs.base + widx + q);
#pragma unroll
                        for (int q = 0; q < NWL; q++) wv[q] = __funnelshift_r(wv[q], wv[q + 1], sh);
                        uint32_t x = wv[0], cs = 0;
                        if (x & 1) cs = dsw[(x >> 1) & LMASK];
                        const uint32_t swl = (x & 1) ? cs >> 12 : 1;
                        if (!(ftl || (cs & 0xfff) != 0 || cs == 0)) slow = true;
                        else {
                            const uint32_t r = (oldrung + cs) & UMASK;
                            if (r == 0) { /* flag, then 16 raw bits (reference: QB3decode.h:148-160) */
                                const uint32_t y = x >> swl;
                                const uint32_t b = (y & 1) ? (y >> 1) & 0xffffu : 0u;
#pragma unroll
                                for (int i = 0; i < 16; i++) d[i] = -(SW)((b >> i) & 1);
                            }
                            else {
                                const uint32_t lens = 0x02000100u + r * 0x01010101u;
                                const uint32_t half = 1u << (r - 1);
                                uint32_t M = 0;
                                /* 8 bit data: table address of a code = table | 4 << r | its r + 2 low bits */
                                const uint32_t cmask = (4u << r) - 1, gb = tbl_sa | (4u << r);
                                /* 16 bit data: arithmetic (QB3decode.h:119-129), then the middle swap of rungs 1..7 */
                                const uint32_t fm1 = 2 * half - 1, sm = r < 8 ? 4 * half - 1 : 0;
                                uint32_t z = x >> swl, used = swl;
                                auto take = [&](const int i, auto step_tag) {
                                    constexpr bool STEP = decltype(step_tag)::value;
                                    const uint32_t len = __byte_perm(lens, 0u, (z & 3) | c4440r);
                                    if (BITS == 8) {
                                        const uint32_t ta = (z & cmask) | gb;
                                        d[i] = lds_s8(ta);
                                        if (STEP) M += lds_u8(ta + 1024) << i;
                                    }
                                    else {
                                        const uint32_t shv = __byte_perm(0x02010201u, 0u, (z & 3) | 0x4440);
                                        uint32_t val = ((z & ((1u << len) - 1)) >> shv) + half * (len - r);
                                        if (val - fm1 <= 1u) val ^= sm;
                                        if (STEP) M += ((val >> r) & 1) << i;
                                        d[i] = (SW)((val >> 1) ^ (0u - (val & 1)));
                                    }
                                    z >>= len;
                                    used += len;
                                };
                                /* after a window the group's words move down by what it used (at most 32 bits); only as
                                   many words as the values still to come can need */
                                auto shift_window = [&](const int left) {
                                    constexpr int MAXC = BITS + 1;
                                    const int need = (left * MAXC + 31) / 32;
#pragma unroll
                                    for (int q = 0; q < NWL; q++)
                                        if (q < need) wv[q] = __funnelshift_rc(wv[q], wv[q + 1], used);
                                    z = wv[0];
                                    used = 0;
                                };
                                /* FTL streams have no step coding: their build of the sixteen values leaves the rung bits alone
                                   (a warp works on one stream, so the branch is uniform) */
                                auto values = [&](auto step_tag) {
                                    if (BITS == 16 && r == 15) { /* two 17 bit codes do not fit a window; rare */
                                        shift_window(16);
    #pragma unroll
                                        for (int i = 0; i < 16; i++) { take(i, step_tag); shift_window(15 - i); }
                                    }
                                    else {
    #pragma unroll
                                        for (int i = 0; i < NV0; i++) take(i, step_tag);
                                        shift_window(16 - NV0);
    #pragma unroll
                                        for (int i0 = NV0; i0 < 16; i0 += VPB) {
    #pragma unroll
                                            for (int i = i0; i < i0 + VPB && i < 16; i++) take(i, step_tag);
                                            if (i0 + VPB < 16) shift_window(16 - i0 - VPB);
                                        }
                                    }
                                };
                                if (ftl) values(std::false_type());
                                else values(std::true_type());
                                if (!ftl) { /* step undo (QB3decode.h:285-289) on the unfolded value: bit r set in a value
                                               moves an even one up by 2^(r-1) and an odd one down */
                                    const int kk = step_decode_index(M);
                                    if (kk >= 0) {
#pragma unroll
                                        for (int i = 0; i < 16; i++)
                                            if (i == kk) d[i] += d[i] < 0 ? -(SW)half : (SW)half;
                                    }
                                }
                            }
                        }
                    }
                }
                if (__any_sync(FULL, slow)) { /* the rare groups, and the last few of a stream */
                    W v[16];
                    if (slow) fuse_group_slow<T>(fs, P, oldrung, (W)0, v, kind, pcw);
                    if (__any_sync(FULL, kind != 0)) {
                        /* the band's last written factor, in group order (QB3decode.h:629-640): rare enough for a walk */
                        W mypc = 0;
                        for (uint32_t t = 0; t < pl.gpi; t++) {
                            if (lane == t && kind == 2) pcfc[c] = pcw;
                            if (lane == t && kind == 1) mypc = pcfc[c];
                            __syncwarp();
                        }
                        if (kind == 1) fuse_group_slow<T>(fs, P, oldrung, mypc, v, kind, pcw);
                    }
                    if (slow) {
#pragma unroll
                        for (int i = 0; i < 16; i++) d[i] = (SW)((v[i] >> 1) ^ ((W)0 - (v[i] & 1)));
                    }
                }
                /* running sum inside the group, then across the band's groups */
                W tot = 0;
#pragma unroll
                for (int i = 0; i < 16; i++) { tot += (W)d[i]; d[i] = (SW)tot; }
                W base;
                if (small_bands) {
                    W inc = tot;
                    for (uint32_t dd = bands; dd < pl.gpi; dd <<= 1) {
                        const W o = __shfl_up_sync(FULL, inc, dd);
                        if (lane >= dd) inc += o;
                    }
                    base = carry + inc - tot;
                    carry += __shfl_sync(FULL, inc, (pl.bpi - 1) * bands + lc);
                }
                else {
                    base = active ? prevS[s * bands + c] : (W)0;
                    __syncwarp();
                    if (active) prevS[s * bands + c] = base + tot;
                }
                T *p = reinterpret_cast<T *>(stage) + (size_t)(min(4 * (b0 + blk), a.w - 4) - xs) * bands + c;
                bool stored = false;
                if constexpr (BITS == 8) {
                    if (simd_out) {
                        uint32_t v[16], R[4];
#pragma unroll
                        for (int i = 0; i < 16; i++) v[i] = (uint32_t)base + (uint32_t)d[i];
                        if (fs.order == HILBERT) pack_rows<HILBERT>(v, R);
                        else pack_rows<ZCURVE>(v, R);
                        if (derived) { /* the core band's pixels of the same block, four at a time (QB3decode.h:730-737) */
                            const uint32_t dm = core != c ? 0xffffffffu : 0u, from = lane + core - c;
#pragma unroll
                            for (int y = 0; y < 4; y++) R[y] = add4(R[y], __shfl_sync(FULL, R[y], from) & dm);
                        }
                        const uint32_t l0 = lane - c;
                        uint8_t *wp = reinterpret_cast<uint8_t *>(p) + 3 * c; /* word c of the block's row */
#pragma unroll
                        for (int y = 0; y < 4; y++) {
                            const uint32_t r0 = __shfl_sync(FULL, R[y], l0), r1 = bands > 1 ? __shfl_sync(FULL, R[y], l0 + 1) : 0u;
                            const uint32_t r2 = bands > 2 ? __shfl_sync(FULL, R[y], l0 + 2) : 0u, r3 = bands > 3 ? __shfl_sync(FULL, R[y], l0 + 3) : 0u;
                            const uint32_t wv = __byte_perm(__byte_perm(r0, r1, selA), __byte_perm(r2, r3, selB), selC);
                            if (active) *reinterpret_cast<uint32_t *>(wp + y * rowpitch) = wv;
                        }
                        stored = true;
                    }
                }
                /* the scattered stores of the lanes named by `mine` */
                auto put_block = [&](bool mine) {
                    if (sweep) {
                        if (mine) {
#pragma unroll
                            for (int i = 0; i < 16; i++) p[poff[i]] = (T)(base + (W)d[i]);
                        }
                        return;
                    }
                    /* core bands first, then the derived bands add their core band's pixels: the same 16 places, one
                       band over (QB3decode.h:730-737) */
                    if (mine && core == c) {
                        if (quanta > 1) {
#pragma unroll
                            for (int i = 0; i < 16; i++)
                                p[poff[i]] = (T)dequantize_value<BITS>((uint64_t)((base + (W)d[i]) & TM), quanta, is_signed);
                        }
                        else {
#pragma unroll
                            for (int i = 0; i < 16; i++) p[poff[i]] = (T)(base + (W)d[i]);
                        }
                    }
                    if (derived) {
                        __syncwarp();
                        if (mine && core != c) {
                            const T *q = p + (int)core - (int)c;
#pragma unroll
                            for (int i = 0; i < 16; i++) p[poff[i]] = (T)(base + (W)d[i] + q[poff[i]]);
                        }
                    }
                };
                if (!stored) {
                    /* A width that is not a multiple of four moves the row's last block back over the one before it, and
                       the reference writes block after block: the last block's pixels are the ones that stay
                       (QB3decode.h:603-737). The two agree on every stream an encoder wrote; on a damaged stream they need
                       not, and the device decodes those like the reference, pixel for pixel. */
                    const bool moved = (a.w & 3) != 0 && b0 + blk == nbx - 1;
                    if (__any_sync(FULL, active && moved)) {
                        put_block(active && !moved);
                        __syncwarp();
                        put_block(active && moved);
                    }
                    else put_block(active);
                }
                __syncwarp();
            }
            if (small_bands && lane < bands) prevS[s * bands + lc] = carry;
            if (sweep) { /* reference: QB3decode.h:730-737 (ascending bands, in place), QB3decode.cpp:434-450 */
                for (uint32_t i = lane; i < 4 * npx; i += 32) {
                    const uint32_t r = i / npx, px = i - r * npx;
                    T *q = reinterpret_cast<T *>(stage) + (size_t)r * rowelems + (size_t)px * bands;
                    for (uint32_t kb = 0; kb < bands; kb++) {
                        const uint32_t kc = cb[kb];
                        if (kc != kb) q[kb] = (T)(q[kb] + q[kc]);
                    }
                    if (quanta > 1)
                        for (uint32_t kb = 0; kb < bands; kb++)
                            q[kb] = (T)dequantize_value<BITS>((uint64_t)q[kb], quanta, is_signed);
                }
                __syncwarp();
            }
            /* staged rows leave as the widest vectors the destination allows */
            T *out = reinterpret_cast<T *>(fs.out);
            const uint32_t rowbytes = npx * bands * (uint32_t)sizeof(T);
            for (uint32_t r = 0; r < 4; r++) {
                uint8_t *gp = reinterpret_cast<uint8_t *>(out + (uint64_t)(y0 + r) * a.stride + (uint64_t)xs * bands);
                const uint8_t *sp = stage + r * rowpitch;
                if ((((uintptr_t)gp | rowbytes) & 15) == 0)
                    for (uint32_t q = 16 * lane; q < rowbytes; q += 16 * 32)
                        st_stream16(gp + q, *reinterpret_cast<const uint4 *>(sp + q));
                else if ((((uintptr_t)gp | rowbytes) & 3) == 0)
                    for (uint32_t q = 4 * lane; q < rowbytes; q += 4 * 32)
                        *reinterpret_cast<uint32_t *>(gp + q) = *reinterpret_cast<const uint32_t *>(sp + q);
                else
                    for (uint32_t q = sizeof(T) * lane; q < rowbytes; q += sizeof(T) * 32)
                        *reinterpret_cast<T *>(gp + q) = *reinterpret_cast<const T *>(sp + q);
            }
            __syncwarp();
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_sa + 8 * (pl.nu + slot));
    }
}

#endif
