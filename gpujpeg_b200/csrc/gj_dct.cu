/*
 * gj_dct.cu -- the two transform kernels of the hot path (sm_100a).
 *
 *   K1  k_fdct_rgb444 : RGB u8 interleaved -> int16 zig-zag coefficients
 *       = colour transform + component split + 8x8 forward DCT + quantisation in ONE pass over HBM
 *       (the reference runs a preprocessor kernel and three DCT launches with a planar u8 round trip
 *        in between: src/gpujpeg_preprocessor.cu:163-201, src/gpujpeg_dct_gpu.cu:180-294).
 *       Algorithmic traffic: 3 B read + 6 B written per pixel.
 *
 *   K4  k_idct_rgb444 : int16 zig-zag coefficients -> RGB u8 interleaved
 *       = dequantisation + inverse DCT + level shift + colour transform + interleave in one pass
 *       (reference: three IDCT launches + a postprocessor kernel, src/gpujpeg_dct_gpu.cu:472-618,
 *        src/gpujpeg_postprocessor.cu:183-216).  6 B read + 3 B written per pixel.
 *
 * Work decomposition (both kernels): one CTA owns a strip of TB = 64 horizontally adjacent 8x8
 * blocks (512 x 8 pixels).  192 threads; during the transform phase thread t owns block (t % 64) of
 * component (t / 64) entirely in registers -- both 1-D passes run without any transpose or shuffle,
 * which matters because the arithmetic must follow the reference's rounding sequence exactly and
 * the kernels are issue-bound, not bandwidth-bound, near the roofline.  Pixels move through shared
 * memory twice: raw interleaved bytes (coalesced 16 B global accesses) and a planar staging area.
 */
#include <cuda_runtime.h>
#include <stdint.h>

#include "gj_device.cuh"
#include "gj_internal.h"

namespace {

constexpr int TB = 64;                 // blocks per strip
constexpr int NT = 192;                // threads per CTA = 3 components x TB
constexpr int STRIP_PX = TB * 8;       // 512 pixels
constexpr int ROW_BYTES = STRIP_PX * 3;  // 1536 raw bytes per strip row
constexpr int BLK_F = 68;              // floats per block in the planar staging area (64 + 4 pad:
                                       // block stride 272 B makes the 16 B row reads of consecutive
                                       // blocks hit distinct bank groups)
constexpr int K1_SMEM = 8 * ROW_BYTES + 3 * TB * BLK_F * 4;   // 12288 + 52224 = 64512 B
constexpr int K4_PLANE = 8 * STRIP_PX;                        // bytes per component plane of a strip
constexpr int K4_SMEM = 8 * ROW_BYTES + 3 * K4_PLANE;         // 12288 + 12288 = 24576 B

struct FdctParams {
    float fwd_zz[2][64];
};
struct IdctParams {
    uint16_t q_zz[3][64];
};

/* ---- raw strip <-> global memory, with whatever alignment the caller's image has ---- */

// copy `nbytes` of each of `rows` image rows into smem rows of ROW_BYTES; VEC = 16, 4 or 1
template <int VEC>
__device__ __forceinline__ void load_strip(uint8_t* s_raw, const uint8_t* g, size_t pitch, int rows, int nbytes)
{
    if ( VEC == 16 ) {
        const int nvec = nbytes >> 4;
        for ( int i = threadIdx.x; i < rows * (ROW_BYTES / 16); i += NT ) {
            const int r = i / (ROW_BYTES / 16), c = i % (ROW_BYTES / 16);
            if ( c < nvec ) {
                const int4 v = __ldg(reinterpret_cast<const int4*>(g + (size_t)r * pitch) + c);
                reinterpret_cast<int4*>(s_raw + r * ROW_BYTES)[c] = v;
            }
        }
        const int tail0 = nvec << 4;
        for ( int i = threadIdx.x; i < rows * 16; i += NT ) {
            const int r = i >> 4, c = tail0 + (i & 15);
            if ( c < nbytes ) s_raw[r * ROW_BYTES + c] = __ldg(g + (size_t)r * pitch + c);
        }
    }
    else if ( VEC == 4 ) {
        const int nvec = nbytes >> 2;
        for ( int i = threadIdx.x; i < rows * (ROW_BYTES / 4); i += NT ) {
            const int r = i / (ROW_BYTES / 4), c = i % (ROW_BYTES / 4);
            if ( c < nvec )
                reinterpret_cast<uint32_t*>(s_raw + r * ROW_BYTES)[c] =
                    __ldg(reinterpret_cast<const uint32_t*>(g + (size_t)r * pitch) + c);
        }
        const int tail0 = nvec << 2;
        for ( int i = threadIdx.x; i < rows * 4; i += NT ) {
            const int r = i >> 2, c = tail0 + (i & 3);
            if ( c < nbytes ) s_raw[r * ROW_BYTES + c] = __ldg(g + (size_t)r * pitch + c);
        }
    }
    else {
        for ( int i = threadIdx.x; i < rows * ROW_BYTES; i += NT ) {
            const int r = i / ROW_BYTES, c = i % ROW_BYTES;
            if ( c < nbytes ) s_raw[r * ROW_BYTES + c] = __ldg(g + (size_t)r * pitch + c);
        }
    }
}

template <int VEC>
__device__ __forceinline__ void store_strip(const uint8_t* s_raw, uint8_t* g, size_t pitch, int rows, int nbytes)
{
    if ( VEC == 16 ) {
        const int nvec = nbytes >> 4;
        for ( int i = threadIdx.x; i < rows * (ROW_BYTES / 16); i += NT ) {
            const int r = i / (ROW_BYTES / 16), c = i % (ROW_BYTES / 16);
            if ( c < nvec )
                reinterpret_cast<int4*>(g + (size_t)r * pitch)[c] = reinterpret_cast<const int4*>(s_raw + r * ROW_BYTES)[c];
        }
        const int tail0 = nvec << 4;
        for ( int i = threadIdx.x; i < rows * 16; i += NT ) {
            const int r = i >> 4, c = tail0 + (i & 15);
            if ( c < nbytes ) g[(size_t)r * pitch + c] = s_raw[r * ROW_BYTES + c];
        }
    }
    else if ( VEC == 4 ) {
        const int nvec = nbytes >> 2;
        for ( int i = threadIdx.x; i < rows * (ROW_BYTES / 4); i += NT ) {
            const int r = i / (ROW_BYTES / 4), c = i % (ROW_BYTES / 4);
            if ( c < nvec )
                reinterpret_cast<uint32_t*>(g + (size_t)r * pitch)[c] =
                    reinterpret_cast<const uint32_t*>(s_raw + r * ROW_BYTES)[c];
        }
        const int tail0 = nvec << 2;
        for ( int i = threadIdx.x; i < rows * 4; i += NT ) {
            const int r = i >> 2, c = tail0 + (i & 3);
            if ( c < nbytes ) g[(size_t)r * pitch + c] = s_raw[r * ROW_BYTES + c];
        }
    }
    else {
        for ( int i = threadIdx.x; i < rows * ROW_BYTES; i += NT ) {
            const int r = i / ROW_BYTES, c = i % ROW_BYTES;
            if ( c < nbytes ) g[(size_t)r * pitch + c] = s_raw[r * ROW_BYTES + c];
        }
    }
}

__device__ __forceinline__ float byte_f(uint32_t w, int i) { return (float)((w >> (8 * i)) & 0xFFu); }

/* =========================================================================================== */
/* K1                                                                                            */

template <int VEC>
__global__ void __launch_bounds__(NT)
k_fdct_rgb444(const uint8_t* __restrict__ raw, int width, int height, size_t pitch, int16_t* __restrict__ coef,
              int bcx, int nblk, const __grid_constant__ FdctParams prm)
{
    extern __shared__ __align__(16) uint8_t smem[];
    uint8_t* s_raw = smem;
    float* s_pl = reinterpret_cast<float*>(smem + 8 * ROW_BYTES);

    const int bx0 = blockIdx.x * TB;
    const int by = blockIdx.y;
    const int x0 = bx0 * 8;
    const int vw = min(STRIP_PX, width - x0);        // valid pixels in this strip (>= 1)
    const int vh = min(8, height - by * 8);          // valid rows (>= 1)

    /* phase A: raw rows -> smem */
    load_strip<VEC>(s_raw, raw + (size_t)by * 8 * pitch + (size_t)x0 * 3, pitch, vh, vw * 3);
    __syncthreads();

    /* phase B: colour transform, 4 pixels (12 bytes = 3 words) per step, planar float staging.
     * Pixels outside the image are 0 in every component, as in the reference whose planes are
     * zero-initialised and only written inside the image [ref: src/gpujpeg_common.c:941-944]. */
    for ( int g = threadIdx.x; g < 8 * (STRIP_PX / 4); g += NT ) {
        const int row = g >> 7, gx = g & 127;
        const int px0 = gx * 4;
        float4 y4 = make_float4(0.f, 0.f, 0.f, 0.f), cb4 = y4, cr4 = y4;
        if ( row < vh && px0 < vw ) {
            const uint32_t* w = reinterpret_cast<const uint32_t*>(s_raw + row * ROW_BYTES + gx * 12);
            const uint32_t w0 = w[0], w1 = w[1], w2 = w[2];
            gj_rgb_to_ycbcr(byte_f(w0, 0), byte_f(w0, 1), byte_f(w0, 2), y4.x, cb4.x, cr4.x);
            gj_rgb_to_ycbcr(byte_f(w0, 3), byte_f(w1, 0), byte_f(w1, 1), y4.y, cb4.y, cr4.y);
            gj_rgb_to_ycbcr(byte_f(w1, 2), byte_f(w1, 3), byte_f(w2, 0), y4.z, cb4.z, cr4.z);
            gj_rgb_to_ycbcr(byte_f(w2, 1), byte_f(w2, 2), byte_f(w2, 3), y4.w, cb4.w, cr4.w);
            if ( px0 + 4 > vw ) {  // strip ends inside this group (width not a multiple of 4)
                if ( px0 + 1 >= vw ) { y4.y = cb4.y = cr4.y = 0.f; }
                if ( px0 + 2 >= vw ) { y4.z = cb4.z = cr4.z = 0.f; }
                if ( px0 + 3 >= vw ) { y4.w = cb4.w = cr4.w = 0.f; }
            }
        }
        const int off = (gx >> 1) * BLK_F + row * 8 + (gx & 1) * 4;
        *reinterpret_cast<float4*>(s_pl + off) = y4;
        *reinterpret_cast<float4*>(s_pl + TB * BLK_F + off) = cb4;
        *reinterpret_cast<float4*>(s_pl + 2 * TB * BLK_F + off) = cr4;
    }
    __syncthreads();

    /* phase C: one thread = one 8x8 block of one component, everything in registers */
    const int comp = threadIdx.x >> 6;
    const int b = threadIdx.x & 63;
    if ( bx0 + b >= bcx ) return;
    float v[64];
    {
        const float4* src = reinterpret_cast<const float4*>(s_pl + (comp * TB + b) * BLK_F);
#pragma unroll
        for ( int i = 0; i < 16; i++ ) {
            const float4 t = src[i];
            v[4 * i] = t.x; v[4 * i + 1] = t.y; v[4 * i + 2] = t.z; v[4 * i + 3] = t.w;
        }
    }
    gj_fdct_block(v);
    /* quantise: q = rint(c * table) [ref: src/gpujpeg_dct_gpu.cu:276-283], emit in zig-zag order */
    const float* tab = prm.fwd_zz[comp == 0 ? 0 : 1];
    uint32_t packed[32];
#pragma unroll
    for ( int k = 0; k < 64; k += 2 ) {
        const int q0 = GJ_RINT(GJ_FMUL(v[gj_zz2nat(k)], tab[k]));
        const int q1 = GJ_RINT(GJ_FMUL(v[gj_zz2nat(k + 1)], tab[k + 1]));
        packed[k >> 1] = ((uint32_t)q0 & 0xFFFFu) | ((uint32_t)q1 << 16);
    }
    uint4* dst = reinterpret_cast<uint4*>(coef + ((size_t)comp * nblk + (size_t)by * bcx + bx0 + b) * 64);
#pragma unroll
    for ( int i = 0; i < 8; i++ )
        dst[i] = make_uint4(packed[4 * i], packed[4 * i + 1], packed[4 * i + 2], packed[4 * i + 3]);
}

/* =========================================================================================== */
/* K4                                                                                            */

template <int VEC, int FLAVOUR>
__global__ void __launch_bounds__(NT)
k_idct_rgb444(const int16_t* __restrict__ coef, int bcx, int nblk, uint8_t* __restrict__ raw, int width, int height,
              size_t pitch, const __grid_constant__ IdctParams prm)
{
    extern __shared__ __align__(16) uint8_t smem[];
    uint8_t* s_raw = smem;
    uint8_t* s_pl = smem + 8 * ROW_BYTES;

    const int bx0 = blockIdx.x * TB;
    const int by = blockIdx.y;
    const int x0 = bx0 * 8;
    const int vw = min(STRIP_PX, width - x0);
    const int vh = min(8, height - by * 8);

    /* phase A: one thread = one block of one component */
    {
        const int comp = threadIdx.x >> 6;
        const int b = threadIdx.x & 63;
        if ( bx0 + b < bcx ) {
            const uint4* src = reinterpret_cast<const uint4*>(coef + ((size_t)comp * nblk + (size_t)by * bcx + bx0 + b) * 64);
            uint32_t packed[32];
#pragma unroll
            for ( int i = 0; i < 8; i++ ) {
                const uint4 t = __ldg(src + i);
                packed[4 * i] = t.x; packed[4 * i + 1] = t.y; packed[4 * i + 2] = t.z; packed[4 * i + 3] = t.w;
            }
            const uint16_t* q = prm.q_zz[comp];
            uint32_t px[16];  // 64 output bytes, row-major
            if ( FLAVOUR == 0 ) {
                /* integer path == gpujpeg_idct_cpu: dequantise with int16 wrap, rows, columns,
                 * +128, clamp [ref: src/gpujpeg_dct_cpu.c:178-189, 239-251] */
                int v[64];
#pragma unroll
                for ( int k = 0; k < 64; k++ ) {
                    const int c = (int)(short)(k & 1 ? packed[k >> 1] >> 16 : packed[k >> 1] & 0xFFFFu);
                    v[gj_zz2nat(k)] = gj_s16(c * (int)(short)q[k]);
                }
                gj_idct_int_block(v);
#pragma unroll
                for ( int i = 0; i < 16; i++ ) {
                    uint32_t w = 0;
#pragma unroll
                    for ( int j = 0; j < 4; j++ ) {
                        const int s = gj_s16(v[4 * i + j] + 128);
                        w |= (uint32_t)gj_clamp8(s) << (8 * j);
                    }
                    px[i] = w;
                }
            }
            else {
                /* float path == the reference CUDA kernel [ref: src/gpujpeg_dct_gpu.cu:497-501, 597-617] */
                float f[64];
#pragma unroll
                for ( int k = 0; k < 64; k++ ) {
                    const int c = (int)(short)(k & 1 ? packed[k >> 1] >> 16 : packed[k >> 1] & 0xFFFFu);
                    f[gj_zz2nat(k)] = (float)(c * (int)q[k]);
                }
                gj_idct_float_block(f);
#pragma unroll
                for ( int i = 0; i < 16; i++ ) {
                    uint32_t w = 0;
#pragma unroll
                    for ( int j = 0; j < 4; j++ )
                        w |= (uint32_t)gj_clamp8(GJ_RINT(GJ_FADD(f[4 * i + j], 128.0f))) << (8 * j);
                    px[i] = w;
                }
            }
            uint8_t* dst = s_pl + comp * K4_PLANE + b * 8;
#pragma unroll
            for ( int r = 0; r < 8; r++ )
                *reinterpret_cast<uint2*>(dst + r * STRIP_PX) = make_uint2(px[2 * r], px[2 * r + 1]);
        }
    }
    __syncthreads();

    /* phase B: 4 pixels per step: 3 plane words -> 3 interleaved words */
    for ( int g = threadIdx.x; g < 8 * (STRIP_PX / 4); g += NT ) {
        const int row = g >> 7, gx = g & 127;
        if ( row >= vh || gx * 4 >= vw ) continue;
        const uint32_t yw = *reinterpret_cast<const uint32_t*>(s_pl + row * STRIP_PX + gx * 4);
        const uint32_t bw = *reinterpret_cast<const uint32_t*>(s_pl + K4_PLANE + row * STRIP_PX + gx * 4);
        const uint32_t rw = *reinterpret_cast<const uint32_t*>(s_pl + 2 * K4_PLANE + row * STRIP_PX + gx * 4);
        int r[4], gg[4], bb[4];
#pragma unroll
        for ( int j = 0; j < 4; j++ )
            gj_ycbcr_to_rgb((yw >> (8 * j)) & 0xFF, (bw >> (8 * j)) & 0xFF, (rw >> (8 * j)) & 0xFF, r[j], gg[j], bb[j]);
        uint32_t* o = reinterpret_cast<uint32_t*>(s_raw + row * ROW_BYTES + gx * 12);
        o[0] = (uint32_t)r[0] | ((uint32_t)gg[0] << 8) | ((uint32_t)bb[0] << 16) | ((uint32_t)r[1] << 24);
        o[1] = (uint32_t)gg[1] | ((uint32_t)bb[1] << 8) | ((uint32_t)r[2] << 16) | ((uint32_t)gg[2] << 24);
        o[2] = (uint32_t)bb[2] | ((uint32_t)r[3] << 8) | ((uint32_t)gg[3] << 16) | ((uint32_t)bb[3] << 24);
    }
    __syncthreads();

    /* phase C: interleaved rows -> global */
    store_strip<VEC>(s_raw, raw + (size_t)by * 8 * pitch + (size_t)x0 * 3, pitch, vh, vw * 3);
}

int pick_vec(const void* p, size_t pitch)
{
    const uintptr_t a = reinterpret_cast<uintptr_t>(p) | pitch;
    return (a & 15) == 0 ? 16 : (a & 3) == 0 ? 4 : 1;
}

template <typename K>
int set_smem(K kernel, int bytes)
{
    return cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes) == cudaSuccess ? 0 : -1;
}

}  // namespace

extern "C" int gj_launch_fdct_rgb444(const uint8_t* d_raw, int width, int height, int pitch, int16_t* d_coef, int bcx,
                                     int bcy, const struct gj_dev_enc_tables* h_tables, gj_stream_t stream)
{
    FdctParams prm;
    memcpy(prm.fwd_zz, h_tables->fwd_zz, sizeof prm.fwd_zz);
    const dim3 grid((bcx + TB - 1) / TB, bcy);
    const int nblk = bcx * bcy;
    const int vec = pick_vec(d_raw, (size_t)pitch);
    static bool attr_done = false;
    if ( !attr_done ) {
        if ( set_smem(k_fdct_rgb444<16>, K1_SMEM) || set_smem(k_fdct_rgb444<4>, K1_SMEM) ||
             set_smem(k_fdct_rgb444<1>, K1_SMEM) )
            return -1;
        attr_done = true;
    }
    if ( vec == 16 )
        k_fdct_rgb444<16><<<grid, NT, K1_SMEM, stream>>>(d_raw, width, height, (size_t)pitch, d_coef, bcx, nblk, prm);
    else if ( vec == 4 )
        k_fdct_rgb444<4><<<grid, NT, K1_SMEM, stream>>>(d_raw, width, height, (size_t)pitch, d_coef, bcx, nblk, prm);
    else
        k_fdct_rgb444<1><<<grid, NT, K1_SMEM, stream>>>(d_raw, width, height, (size_t)pitch, d_coef, bcx, nblk, prm);
    return cudaGetLastError() == cudaSuccess ? 0 : -1;
}

extern "C" int gj_launch_idct_rgb444(const int16_t* d_coef, int bcx, int bcy, const int comp_tq[3], uint8_t* d_raw,
                                     int width, int height, int pitch, int idct_flavour,
                                     const struct gj_dev_dec_tables* h_tables, gj_stream_t stream)
{
    IdctParams prm;
    for ( int c = 0; c < 3; c++ )
        memcpy(prm.q_zz[c], h_tables->qinv_zz[comp_tq[c]], sizeof prm.q_zz[c]);
    const dim3 grid((bcx + TB - 1) / TB, bcy);
    const int nblk = bcx * bcy;
    const int vec = pick_vec(d_raw, (size_t)pitch);
#define GJ_K4(V, F) k_idct_rgb444<V, F><<<grid, NT, K4_SMEM, stream>>>(d_coef, bcx, nblk, d_raw, width, height, (size_t)pitch, prm)
    if ( idct_flavour == 0 ) {
        if ( vec == 16 ) GJ_K4(16, 0); else if ( vec == 4 ) GJ_K4(4, 0); else GJ_K4(1, 0);
    }
    else {
        if ( vec == 16 ) GJ_K4(16, 1); else if ( vec == 4 ) GJ_K4(4, 1); else GJ_K4(1, 1);
    }
#undef GJ_K4
    return cudaGetLastError() == cudaSuccess ? 0 : -1;
}
