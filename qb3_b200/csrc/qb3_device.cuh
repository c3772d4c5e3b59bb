/*
 * qb3_device.cuh -- launch argument blocks and small device helpers shared by the kernels.
 */
#ifndef QB3_B200_DEVICE_CUH
#define QB3_B200_DEVICE_CUH

#include <cuda_runtime.h>
#include <stdint.h>

#include <mutex>
#include <set>
#include <utility>

#include "qb3_codes.h"
#include "../../include/qb3cu.h"

namespace qb3 {

constexpr int MAXBANDS = 256;
constexpr int MAXHDR = 320; /* 11 + CB(4+256) + QV(4+8) + SC(12) + DT(2) = 297 */

/* Everything the encode kernel needs, passed by value as a __grid_constant__ parameter. */
struct EncArgs {
    const uint8_t *src;
    uint64_t src_pitch;   /* bytes between tiles */
    uint8_t *dst;
    uint64_t slot;        /* bytes per destination slot, multiple of 16 */
    unsigned long long *sizes;
    uint32_t *status;
    unsigned long long *state; /* optional [tile][3][bands]: prev, runbits, cf */
    uint64_t stride;      /* source line stride in values */
    uint64_t quanta;
    uint64_t order;       /* scan curve actually used */
    uint64_t raw_size;    /* w*h*bands*sizeof(T) */
    uint64_t max_size;    /* qb3_max_encoded_size */
    uint32_t w, h, bands; /* image geometry */
    uint32_t vw, vh;      /* geometry being coded: the image, or its small-image reorder */
    uint32_t small;       /* 0 none, 1 narrow (w < 4): rows concatenated, 2 short (h < 4): column major pixels */
    uint32_t nbx, nby;    /* 4x4 blocks per row / column of the coded geometry */
    uint32_t seg_blocks;  /* blocks handled per CTA iteration */
    uint32_t segs;        /* iterations per block row */
    uint32_t mode;        /* stream mode with RLE stripped: 0, 1, 4, 5 or 8 */
    uint32_t rle_mode;    /* 0, or the RLE mode byte the caller asked for (2, 3, 6, 7) */
    uint32_t is_signed, away;
    uint32_t vec_stage;   /* rows can be staged with 16 byte copies (no quantisation, no reorder) */
    uint32_t bulk_stage;  /* ... and every staged row is whole 16 byte units at a 16 byte address: bulk copies (TMA) */
    uint32_t simd8;       /* ... and 8 bit data, at most four bands, width a multiple of four: byte SIMD front end */
    uint32_t rowpitch;    /* bytes between staged rows in shared memory, multiple of 16 */
    uint32_t win_words;   /* bit window size in 32 bit words */
    uint32_t best_off;    /* byte offset of the BEST mode scratch in shared memory, 8 byte aligned */
    uint32_t lut_off;     /* byte offset of the code tables in shared memory, 8 byte aligned */
    uint32_t hdr_len, hdr_stored_len;
    /* a tile coded by several CTAs (few, large tiles): each takes part_rows block rows and writes its bits, from bit 0,
       into its own region of tmp; stitch_kernel then joins the parts in the tile's slot */
    uint32_t parts, part_rows;
    uint64_t tmp_slot;          /* bytes per part region, multiple of 16 */
    uint8_t *tmp;
    unsigned long long *part_bits; /* [tile][part] bits written, all ones when the part did not fit */
    /* BEST in parts. The band's last written factor reaches back without bound, so a part other than the first starts
       without knowing it (best_pass 1) and notes per band which factors it met while it did not know (part_pcf:
       [tile][part][band] x {factor going out, one if the part wrote one, mask of the factors that depended on the one
       coming in (by their six low bits), factor coming in}); best_resolve_kernel then hands the factors down the parts
       and marks in part_redo the parts that met the very factor that came in -- those are coded again knowing it
       (best_pass 2). */
    uint32_t best_pass;
    uint32_t size_only;         /* nothing is packed or stored, only sizes[] is written (the band map search of cqb3 -m x) */
    unsigned long long *part_pcf;
    uint32_t *part_redo;        /* [tile][part] */
    uint8_t hdr[MAXHDR];        /* headers up to and including "DT", mode byte = mode */
    uint8_t hdr_stored[MAXHDR]; /* same for the stored fallback (mode 255, no CB / SC) */
    uint8_t cband[MAXBANDS];
};

struct DecArgs {
    const uint8_t *streams;
    const unsigned long long *offsets, *lens;
    uint8_t *dst;
    uint64_t dst_pitch;   /* bytes between tiles */
    uint32_t *status;
    uint64_t stride;      /* destination line stride in values */
    uint32_t w, h, bands, dtype;
    uint32_t ref_compat;
    uint32_t ntiles;
    uint32_t rle_hint;    /* host side only: the caller expects RLE streams (cfg->mode is an RLE mode): expand them first */
    uint32_t shared_sm;   /* host side only: the batch shares the device with other kernels (many batches in flight at once) */
};

/* header fields of one stream, parsed on the device */
struct StreamInfo {
    uint64_t order, quanta;
    uint32_t mode;
    uint32_t data_off;    /* payload offset from the stream start */
    uint32_t has_cb;
    uint32_t bad;
};

__device__ __forceinline__ uint32_t lane_id() { return threadIdx.x & 31; }

/* streaming 16 byte global load, bypassing L1 allocation: every input byte is read once */
__device__ __forceinline__ uint4 ld_stream16(const void *p)
{
    uint4 v;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
    return v;
}
/* streaming 16 byte global store */
__device__ __forceinline__ void st_stream16(void *p, uint4 v)
{
    asm volatile("st.global.cs.v4.u32 [%0], {%1,%2,%3,%4};" :: "l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}

/* cp.async: 16 bytes global -> shared without passing through registers (LDGSTS) */
__device__ __forceinline__ void cp_async16(void *smem_dst, const void *gmem_src)
{
    const uint32_t d = (uint32_t)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" :: "r"(d), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" :: "n"(N) : "memory"); }

/* CTA wide exclusive scan of one uint32 per thread; returns the prefix, total gets the sum. blockDim.x <= 1024.
   scratch: 33 words of shared memory. Contains two __syncthreads(). */
__device__ __forceinline__ uint32_t block_exclusive_scan(uint32_t v, uint32_t *scratch, uint32_t &total)
{
    const uint32_t lane = lane_id(), warp = threadIdx.x >> 5, nwarps = (blockDim.x + 31) >> 5;
    uint32_t inc = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        uint32_t o = __shfl_up_sync(0xffffffffu, inc, d);
        if (lane >= d) inc += o;
    }
    if (lane == 31) scratch[warp] = inc;
    __syncthreads();
    if (warp == 0) {
        uint32_t s = lane < nwarps ? scratch[lane] : 0, t = s;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            uint32_t o = __shfl_up_sync(0xffffffffu, t, d);
            if (lane >= d) t += o;
        }
        scratch[lane] = t - s; /* exclusive prefix of the warp sums */
        if (lane == 31) scratch[32] = t;
    }
    __syncthreads();
    total = scratch[32];
    return scratch[warp] + inc - v;
}

/*
 * cudaFuncAttributeMaxDynamicSharedMemorySize belongs to the kernel, process wide: it is raised once per kernel and
 * device to everything the device allows, never per call (two host threads launching the same kernel with different
 * sizes would otherwise race between setting it and launching).
 */
inline cudaError_t allow_max_smem_of(const void *kernel)
{
    static std::mutex mu;
    static std::set<std::pair<int, const void *>> done;
    int dev = 0, optin = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    std::lock_guard<std::mutex> lock(mu);
    if (done.count(std::make_pair(dev, kernel))) return cudaSuccess;
    cudaFuncAttributes fa;
    e = cudaFuncGetAttributes(&fa, kernel);
    if (e == cudaSuccess) e = cudaDeviceGetAttribute(&optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, optin - (int)fa.sharedSizeBytes);
    if (e == cudaSuccess) done.insert(std::make_pair(dev, kernel));
    return e;
}
template <auto Kernel> static cudaError_t allow_max_smem() { return allow_max_smem_of(reinterpret_cast<const void *>(Kernel)); }

} // namespace qb3
#endif
