/*
 * gj_markers.cu -- K0: finds every marker inside the entropy-coded part of a JPEG file on the GPU.
 *
 * The reference splits scans into restart segments on the host with a memchr(0xFF) walk over the whole
 * stream and copies every segment into a second buffer (src/gpujpeg_reader.c:1038-1155); its FAQ quotes
 * 543 ms for that step on one sample.  On a B200 host the same walk costs 2.1 ms for an 8K frame --
 * six times the GPU time of the entire decode.  Here the file is uploaded once, untouched, and three
 * small launches build the ordered marker list (position, code) directly in device memory:
 *
 *     k_marker_count   per-CTA marker counts (16 bytes per thread, one 16-byte load)
 *     k_marker_scan    exclusive scan of the CTA counts (one CTA)
 *     k_marker_write   ordered compaction: list[rank] = {position, code}; markers other than RSTn are
 *                      also appended (with their rank) to a short list the host reads back
 *
 * Inside entropy-coded data a 0xFF byte is always followed by 0x00 (stuffing), 0xFF (fill) or a marker
 * code, so "FF xx, xx not in {00, FF}" is exact there.  The host only trusts the list inside scans; the
 * marker segments between scans are walked on the host by their length fields (gj_decoder.c).
 */
#include <cuda_runtime.h>
#include <stdint.h>

#include "gj_internal.h"

namespace {

constexpr int MK_THREADS = 256;
constexpr int MK_BYTES = 16;                       // bytes per thread
constexpr int MK_TILE = MK_THREADS * MK_BYTES;     // bytes per CTA
constexpr unsigned FULL = 0xFFFFFFFFu;

/* bit i of the result is set when byte i of the thread's 16-byte chunk starts a marker */
__device__ __forceinline__ uint32_t marker_bits(const uint8_t* __restrict__ file, size_t begin, size_t end, size_t pos,
                                                uint32_t (&w)[5])
{
    // chunk [pos, pos+16) plus one look-ahead byte; positions are relative to the 16-byte aligned `begin`
    if ( pos + MK_BYTES + 4 <= end ) {
        const uint4 v = __ldg(reinterpret_cast<const uint4*>(file + pos));
        w[0] = v.x; w[1] = v.y; w[2] = v.z; w[3] = v.w;
        w[4] = __ldg(reinterpret_cast<const uint32_t*>(file + pos + 16));
    }
    else {
#pragma unroll
        for ( int i = 0; i < 5; i++ ) {
            uint32_t x = 0;
#pragma unroll
            for ( int j = 0; j < 4; j++ ) {
                const size_t p = pos + 4 * i + j;
                x |= (p < end ? (uint32_t)file[p] : 0u) << (8 * j);
            }
            w[i] = x;
        }
    }
    uint32_t bits = 0;
#pragma unroll
    for ( int i = 0; i < MK_BYTES; i++ ) {
        const uint32_t b0 = (w[i >> 2] >> (8 * (i & 3))) & 0xFFu;
        const uint32_t b1 = (w[(i + 1) >> 2] >> (8 * ((i + 1) & 3))) & 0xFFu;
        if ( b0 == 0xFFu && b1 != 0u && b1 != 0xFFu && pos + i >= begin && pos + i + 1 < end ) bits |= 1u << i;
    }
    return bits;
}

__global__ void __launch_bounds__(MK_THREADS)
k_marker_count(const uint8_t* __restrict__ file, size_t begin, size_t end, size_t base, uint32_t* __restrict__ cta_count)
{
    __shared__ uint32_t s_warp[MK_THREADS / 32];
    const size_t pos = base + (size_t)blockIdx.x * MK_TILE + (size_t)threadIdx.x * MK_BYTES;
    uint32_t w[5];
    uint32_t n = pos < end ? __popc(marker_bits(file, begin, end, pos, w)) : 0u;
#pragma unroll
    for ( int d = 16; d > 0; d >>= 1 )
        n += __shfl_down_sync(FULL, n, d);
    if ( (threadIdx.x & 31) == 0 ) s_warp[threadIdx.x >> 5] = n;
    __syncthreads();
    if ( threadIdx.x == 0 ) {
        uint32_t t = 0;
        for ( int i = 0; i < MK_THREADS / 32; i++ )
            t += s_warp[i];
        cta_count[blockIdx.x] = t;
    }
}

__global__ void __launch_bounds__(1024)
k_marker_scan(uint32_t* __restrict__ cta_count, int n_cta, uint32_t* __restrict__ result /*[0]=total*/)
{
    __shared__ uint32_t s_warp[32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t carry = 0;
    for ( int base = 0; base < n_cta; base += 1024 ) {
        const int i = base + threadIdx.x;
        const uint32_t v = i < n_cta ? cta_count[i] : 0u;
        uint32_t incl = v;
#pragma unroll
        for ( int d = 1; d < 32; d <<= 1 ) {
            const uint32_t t = __shfl_up_sync(FULL, incl, d);
            if ( lane >= d ) incl += t;
        }
        if ( lane == 31 ) s_warp[warp] = incl;
        __syncthreads();
        if ( warp == 0 ) {
            uint32_t x = s_warp[lane];
#pragma unroll
            for ( int d = 1; d < 32; d <<= 1 ) {
                const uint32_t t = __shfl_up_sync(FULL, x, d);
                if ( lane >= d ) x += t;
            }
            s_warp[lane] = x;
        }
        __syncthreads();
        if ( i < n_cta ) cta_count[i] = carry + (warp ? s_warp[warp - 1] : 0u) + incl - v;   // exclusive
        carry += s_warp[31];
        __syncthreads();
    }
    if ( threadIdx.x == 0 ) result[0] = carry;
}

__global__ void __launch_bounds__(MK_THREADS)
k_marker_write(const uint8_t* __restrict__ file, size_t begin, size_t end, size_t base, const uint32_t* __restrict__ cta_base,
               uint32_t* __restrict__ list_pos, uint8_t* __restrict__ list_code, uint32_t list_cap,
               uint32_t* __restrict__ result /*[1]=other count, [2]=overflow*/, uint32_t* __restrict__ other /*{rank,pos,code}*/,
               uint32_t other_cap)
{
    __shared__ uint32_t s_warp[MK_THREADS / 32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const size_t pos = base + (size_t)blockIdx.x * MK_TILE + (size_t)threadIdx.x * MK_BYTES;
    uint32_t w[5];
    uint32_t bits = pos < end ? marker_bits(file, begin, end, pos, w) : 0u;
    const uint32_t n = __popc(bits);
    uint32_t incl = n;
#pragma unroll
    for ( int d = 1; d < 32; d <<= 1 ) {
        const uint32_t t = __shfl_up_sync(FULL, incl, d);
        if ( lane >= d ) incl += t;
    }
    if ( lane == 31 ) s_warp[warp] = incl;
    __syncthreads();
    uint32_t rank = cta_base[blockIdx.x] + incl - n;
    for ( int i = 0; i < warp; i++ )
        rank += s_warp[i];
    while ( bits ) {
        const int i = __ffs(bits) - 1;
        bits &= bits - 1;
        const uint32_t code = (w[(i + 1) >> 2] >> (8 * ((i + 1) & 3))) & 0xFFu;
        if ( rank < list_cap ) {
            list_pos[rank] = (uint32_t)(pos + i);
            list_code[rank] = (uint8_t)code;
        }
        else {
            result[2] = 1;
        }
        if ( code < 0xD0u || code > 0xD7u ) {
            const uint32_t k = atomicAdd(&result[1], 1u);
            if ( k < other_cap ) {
                other[3 * k] = rank;
                other[3 * k + 1] = (uint32_t)(pos + i);
                other[3 * k + 2] = code;
            }
        }
        rank++;
    }
}

struct ScanBounds {
    uint32_t begin[4], end[4];
    int segments[4];   // restart segments the geometry expects in each scan
};
/* one thread per scan: rank of its first marker in the list and a check of the restart count */
__global__ void k_scan_ranks(const uint32_t* __restrict__ list_pos, uint32_t* result, int scan_count, ScanBounds sb,
                             uint32_t* __restrict__ first_rank)
{
    const int s = threadIdx.x;
    if ( s >= scan_count ) return;
    const uint32_t total = result[0];
    uint32_t r[2];
    const uint32_t key[2] = {sb.begin[s], sb.end[s]};
    for ( int q = 0; q < 2; q++ ) {   // lower_bound
        uint32_t lo = 0, hi = total;
        while ( lo < hi ) {
            const uint32_t mid = (lo + hi) >> 1;
            if ( list_pos[mid] < key[q] ) lo = mid + 1;
            else hi = mid;
        }
        r[q] = lo;
    }
    first_rank[s] = r[0];
    if ( r[1] - r[0] != (uint32_t)(sb.segments[s] - 1) ) atomicExch(&result[3], 1u + (uint32_t)s);
}

}  // namespace

extern "C" int gj_launch_scan_ranks(const uint32_t* d_list_pos, const uint32_t* d_result, int scan_count,
                                    const uint32_t scan_begin[4], const uint32_t scan_end[4], const int scan_segments[4],
                                    uint32_t* d_first_rank, gj_stream_t stream)
{
    ScanBounds sb;
    for ( int i = 0; i < 4; i++ ) {
        sb.begin[i] = scan_begin[i];
        sb.end[i] = scan_end[i];
        sb.segments[i] = scan_segments[i];
    }
    k_scan_ranks<<<1, 32, 0, stream>>>(d_list_pos, const_cast<uint32_t*>(d_result), scan_count, sb, d_first_rank);
    return cudaGetLastError() == cudaSuccess ? 0 : -1;
}

/* Scans file[begin, end) (device memory; begin need not be aligned).  Outputs, all in device memory:
 *   d_list_pos/d_list_code[0..total) : every marker in file order
 *   d_result[0] = total, [1] = number of non-RST markers, [2] = overflow flag
 *   d_other[3*k..] = {rank, position, code} of the non-RST markers (unordered, at most other_cap)
 * d_cta must hold ceil((end - begin + 16) / 4096) + 1 words. */
extern "C" int gj_launch_marker_scan(const uint8_t* d_file, size_t begin, size_t end, uint32_t* d_cta, uint32_t* d_list_pos,
                                     uint8_t* d_list_code, uint32_t list_cap, uint32_t* d_result, uint32_t* d_other,
                                     uint32_t other_cap, gj_stream_t stream)
{
    if ( end <= begin ) return -1;
    /* d_file comes from cudaMalloc (256-byte aligned): tile from a 16-byte aligned base below `begin` */
    const size_t base = begin & ~static_cast<size_t>(15);
    const int n_cta = (int)((end - base + MK_TILE - 1) / MK_TILE);
    if ( cudaMemsetAsync(d_result, 0, 4 * sizeof(uint32_t), stream) != cudaSuccess ) return -1;
    k_marker_count<<<n_cta, MK_THREADS, 0, stream>>>(d_file, begin, end, base, d_cta);
    k_marker_scan<<<1, 1024, 0, stream>>>(d_cta, n_cta, d_result);
    k_marker_write<<<n_cta, MK_THREADS, 0, stream>>>(d_file, begin, end, base, d_cta, d_list_pos, d_list_code, list_cap,
                                                     d_result, d_other, other_cap);
    return cudaGetLastError() == cudaSuccess ? 0 : -1;
}
