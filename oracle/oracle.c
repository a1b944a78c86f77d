/*
 * oracle.c -- CPU restatement of the reference's JPEG hot path.  TEST INFRASTRUCTURE ONLY
 * (see oracle.h for who may use it and how it is pinned).
 *
 * Every function cites the reference file:line whose behaviour it restates (paths relative to the
 * reference tree root).  Nothing here is copied: the arithmetic is re-derived from the survey's
 * appendix A and re-checked against the reference sources compiled in place (oracle/_ref).
 *
 * Build: gcc -O2 -ffp-contract=off -fopenmp -shared -fPIC   (contraction OFF is mandatory: the
 * float FDCT must round exactly where the reference CUDA kernel rounds; fused ops are spelled
 * out with fmaf()).
 */
#include "oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

/* ------------------------------------------------------------------------------------------- */
/* scratch buffers are kept between calls (grow-only, not thread-safe across concurrent callers): a
 * fresh malloc of ~0.5 GB per 8K frame spends more time in page faults than in the codec once the
 * loops run on 16+ threads */
static void* scratch(int slot, size_t size)
{
    static void* buf[8];
    static size_t cap[8];
    if ( cap[slot] < size ) {
        free(buf[slot]);
        buf[slot] = malloc(size);
        cap[slot] = buf[slot] ? size : 0;
    }
    return buf[slot];
}

/* ------------------------------------------------------------------------------------------- */
/* constant tables                                                                              */

/* zig-zag index -> natural (row-major) index.  [ref: src/gpujpeg_table.h:73-84] */
const uint8_t orc_zigzag_to_natural[64] = {
    0,  1,  8,  16, 9,  2,  3,  10, 17, 24, 32, 25, 18, 11, 4,  5,  12, 19, 26, 33, 40, 48,
    41, 34, 27, 20, 13, 6,  7,  14, 21, 28, 35, 42, 49, 56, 57, 50, 43, 36, 29, 22, 15, 23,
    30, 37, 44, 51, 58, 59, 52, 45, 38, 31, 39, 46, 53, 60, 61, 54, 47, 55, 62, 63};

/* JPEG Annex K.1 base quantisation tables, stored in zig-zag order as the reference keeps them.
 * [ref: src/gpujpeg_table.c:36-56] */
static const uint8_t k_base_quant[2][64] = {
    {16, 11, 12, 14, 12, 10, 16, 14, 13, 14, 18, 17, 16, 19, 24, 40, 26, 24, 22, 22, 24, 49,
     35, 37, 29, 40, 58, 51, 61, 60, 57, 51, 56, 55, 64, 72, 92, 78, 64, 68, 87, 69, 55, 56,
     80, 109, 81, 87, 95, 98, 103, 104, 103, 62, 77, 113, 121, 112, 100, 120, 92, 101, 103, 99},
    {17, 18, 18, 24, 21, 24, 47, 26, 26, 47, 99, 66, 56, 66, 99, 99, 99, 99, 99, 99, 99, 99,
     99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99,
     99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99}};

/* JPEG Annex K.3 Huffman specifications: BITS (index 1..16 used) and HUFFVAL.
 * [ref: src/gpujpeg_table.c:190-256] */
static const uint8_t k_bits_dc_y[17] = {0, 0, 1, 5, 1, 1, 1, 1, 1, 1, 0, 0, 0, 0, 0, 0, 0};
static const uint8_t k_bits_dc_c[17] = {0, 0, 3, 1, 1, 1, 1, 1, 1, 1, 1, 1, 0, 0, 0, 0, 0};
static const uint8_t k_vals_dc[12] = {0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11};
static const uint8_t k_bits_ac_y[17] = {0, 0, 2, 1, 3, 3, 2, 4, 3, 5, 5, 4, 4, 0, 0, 1, 0x7d};
static const uint8_t k_bits_ac_c[17] = {0, 0, 2, 1, 2, 4, 4, 3, 4, 7, 5, 4, 4, 0, 1, 2, 0x77};
static const uint8_t k_vals_ac_y[162] = {
    0x01, 0x02, 0x03, 0x00, 0x04, 0x11, 0x05, 0x12, 0x21, 0x31, 0x41, 0x06, 0x13, 0x51, 0x61, 0x07, 0x22, 0x71,
    0x14, 0x32, 0x81, 0x91, 0xa1, 0x08, 0x23, 0x42, 0xb1, 0xc1, 0x15, 0x52, 0xd1, 0xf0, 0x24, 0x33, 0x62, 0x72,
    0x82, 0x09, 0x0a, 0x16, 0x17, 0x18, 0x19, 0x1a, 0x25, 0x26, 0x27, 0x28, 0x29, 0x2a, 0x34, 0x35, 0x36, 0x37,
    0x38, 0x39, 0x3a, 0x43, 0x44, 0x45, 0x46, 0x47, 0x48, 0x49, 0x4a, 0x53, 0x54, 0x55, 0x56, 0x57, 0x58, 0x59,
    0x5a, 0x63, 0x64, 0x65, 0x66, 0x67, 0x68, 0x69, 0x6a, 0x73, 0x74, 0x75, 0x76, 0x77, 0x78, 0x79, 0x7a, 0x83,
    0x84, 0x85, 0x86, 0x87, 0x88, 0x89, 0x8a, 0x92, 0x93, 0x94, 0x95, 0x96, 0x97, 0x98, 0x99, 0x9a, 0xa2, 0xa3,
    0xa4, 0xa5, 0xa6, 0xa7, 0xa8, 0xa9, 0xaa, 0xb2, 0xb3, 0xb4, 0xb5, 0xb6, 0xb7, 0xb8, 0xb9, 0xba, 0xc2, 0xc3,
    0xc4, 0xc5, 0xc6, 0xc7, 0xc8, 0xc9, 0xca, 0xd2, 0xd3, 0xd4, 0xd5, 0xd6, 0xd7, 0xd8, 0xd9, 0xda, 0xe1, 0xe2,
    0xe3, 0xe4, 0xe5, 0xe6, 0xe7, 0xe8, 0xe9, 0xea, 0xf1, 0xf2, 0xf3, 0xf4, 0xf5, 0xf6, 0xf7, 0xf8, 0xf9, 0xfa};
static const uint8_t k_vals_ac_c[162] = {
    0x00, 0x01, 0x02, 0x03, 0x11, 0x04, 0x05, 0x21, 0x31, 0x06, 0x12, 0x41, 0x51, 0x07, 0x61, 0x71, 0x13, 0x22,
    0x32, 0x81, 0x08, 0x14, 0x42, 0x91, 0xa1, 0xb1, 0xc1, 0x09, 0x23, 0x33, 0x52, 0xf0, 0x15, 0x62, 0x72, 0xd1,
    0x0a, 0x16, 0x24, 0x34, 0xe1, 0x25, 0xf1, 0x17, 0x18, 0x19, 0x1a, 0x26, 0x27, 0x28, 0x29, 0x2a, 0x35, 0x36,
    0x37, 0x38, 0x39, 0x3a, 0x43, 0x44, 0x45, 0x46, 0x47, 0x48, 0x49, 0x4a, 0x53, 0x54, 0x55, 0x56, 0x57, 0x58,
    0x59, 0x5a, 0x63, 0x64, 0x65, 0x66, 0x67, 0x68, 0x69, 0x6a, 0x73, 0x74, 0x75, 0x76, 0x77, 0x78, 0x79, 0x7a,
    0x82, 0x83, 0x84, 0x85, 0x86, 0x87, 0x88, 0x89, 0x8a, 0x92, 0x93, 0x94, 0x95, 0x96, 0x97, 0x98, 0x99, 0x9a,
    0xa2, 0xa3, 0xa4, 0xa5, 0xa6, 0xa7, 0xa8, 0xa9, 0xaa, 0xb2, 0xb3, 0xb4, 0xb5, 0xb6, 0xb7, 0xb8, 0xb9, 0xba,
    0xc2, 0xc3, 0xc4, 0xc5, 0xc6, 0xc7, 0xc8, 0xc9, 0xca, 0xd2, 0xd3, 0xd4, 0xd5, 0xd6, 0xd7, 0xd8, 0xd9, 0xda,
    0xe2, 0xe3, 0xe4, 0xe5, 0xe6, 0xe7, 0xe8, 0xe9, 0xea, 0xf2, 0xf3, 0xf4, 0xf5, 0xf6, 0xf7, 0xf8, 0xf9, 0xfa};

/* test hook: Huffman tables other than Annex K (what an optimising foreign encoder writes).  The override is used
 * by the header writer and by the encoder until it is cleared again with bits17 == NULL. */
static uint8_t g_custom_bits[2][2][17], g_custom_vals[2][2][256];
static int g_custom_n[2][2], g_custom_on = 0;
static int g_enc_ready;
void orc_set_huffman_override(int cls, int kind, const uint8_t* bits17, const uint8_t* vals, int nvals)
{
    if ( !bits17 ) {
        g_custom_on = 0;
        memset(g_custom_n, 0, sizeof g_custom_n);
    }
    else {
        memcpy(g_custom_bits[cls][kind], bits17, 17);
        memcpy(g_custom_vals[cls][kind], vals, (size_t)nvals);
        g_custom_n[cls][kind] = nvals;
        g_custom_on = 1;
    }
    g_enc_ready = 0;
}

void orc_huff_spec(int cls, int kind, const uint8_t** bits17, const uint8_t** vals, int* nvals)
{
    if ( g_custom_on && g_custom_n[cls][kind] ) {
        *bits17 = g_custom_bits[cls][kind];
        *vals = g_custom_vals[cls][kind];
        *nvals = g_custom_n[cls][kind];
        return;
    }
    if ( kind == 0 ) {
        *bits17 = cls == 0 ? k_bits_dc_y : k_bits_dc_c;
        *vals = k_vals_dc;
        *nvals = 12;
    }
    else {
        *bits17 = cls == 0 ? k_bits_ac_y : k_bits_ac_c;
        *vals = cls == 0 ? k_vals_ac_y : k_vals_ac_c;
        *nvals = 162;
    }
}

/* ------------------------------------------------------------------------------------------- */
/* synthetic inputs                                                                             */

/* [ref: src/utils/image_delegate.c:561-581] uint32 LCG, byte = state % 256 */
void orc_gen_random(uint8_t* data, size_t len, int seed)
{
    uint32_t state = (uint32_t)seed;
    for ( size_t i = 0; i < len; i++ ) {
        state = (1664525u * state + 1013904223u) % 2147483647u;
        data[i] = (uint8_t)(state % 256u);
    }
}

/* [ref: src/utils/image_delegate.c:596-603] every byte of row y = y*255/height */
void orc_gen_gradient(uint8_t* data, int width, int height, int bpp)
{
    size_t line = (size_t)width * bpp;
    for ( int y = 0; y < height; y++ )
        memset(data + (size_t)y * line, y * 255 / height, line);
}

/* S-photo: smooth triangle waves plus 5 bits of LCG noise, integer only. */
static inline int tri(int t)
{
    int m = t % 512;
    int a = m - 256;
    if ( a < 0 ) a = -a;
    return a > 255 ? 255 : a;
}
void orc_gen_photo(uint8_t* data, int width, int height, int seed)
{
    uint32_t state = (uint32_t)seed;
    size_t i = 0;
    for ( int y = 0; y < height; y++ ) {
        for ( int x = 0; x < width; x++ ) {
            for ( int ch = 0; ch < 3; ch++ ) {
                state = (1664525u * state + 1013904223u) % 2147483647u;
                int noise = (int)((state % 256u) & 31u) - 16;
                int v = (tri((int)((long long)x * 1024 / width) + 85 * ch) +
                         tri((int)((long long)y * 768 / height) + 40 * ch)) / 2 + noise;
                data[i++] = (uint8_t)(v < 0 ? 0 : v > 255 ? 255 : v);
            }
        }
    }
}

/* ------------------------------------------------------------------------------------------- */
/* tables                                                                                        */

/* quality scaling  [ref: src/gpujpeg_table.c:83-99]; forward float table [ref: :102-129];
 * inverse natural-order table [ref: :133-166] */
void orc_quant_tables(int quality, uint8_t raw[2][64], float fwd[2][64], uint16_t inv[2][64])
{
    static const double aan[8] = {1.0, 1.387039845, 1.306562965, 1.175875602,
                                  1.0, 0.785694958, 0.541196100, 0.275899379};
    if ( quality <= 0 ) quality = 1;
    if ( quality > 100 ) quality = 100;
    int s = quality < 50 ? 5000 / quality : 200 - 2 * quality;
    for ( int t = 0; t < 2; t++ ) {
        for ( int i = 0; i < 64; i++ ) {
            int v = (s * (int)k_base_quant[t][i] + 50) / 100;
            if ( v == 0 ) v = 1;
            if ( v > 255 ) v = 255;
            raw[t][i] = (uint8_t)v;
        }
        for ( int i = 0; i < 64; i++ ) {
            int n = orc_zigzag_to_natural[i];
            int x = n % 8, y = n / 8;
            if ( fwd ) fwd[t][x * 8 + y] = (float)(1.0 / (raw[t][i] * aan[x] * aan[y] * 8));
            if ( inv ) inv[t][n] = raw[t][i];
        }
    }
}

/* canonical code assignment, JPEG Annex C figures C.1-C.3  [ref: src/gpujpeg_table.c:264-306] */
static int huff_codes(const uint8_t* bits17, uint16_t* codes, uint8_t* sizes)
{
    int p = 0;
    unsigned code = 0;
    for ( int l = 1; l <= 16; l++ ) {
        for ( int i = 0; i < bits17[l]; i++ ) {
            codes[p] = (uint16_t)code++;
            sizes[p] = (uint8_t)l;
            p++;
        }
        code <<= 1;
    }
    return p;
}

void orc_huff_encoder_table(int cls, int kind, uint16_t code[256], uint8_t size[256])
{
    const uint8_t *bits, *vals;
    int n;
    uint16_t c[256];
    uint8_t s[256];
    orc_huff_spec(cls, kind, &bits, &vals, &n);
    int cnt = huff_codes(bits, c, s);
    memset(code, 0, 256 * sizeof(uint16_t));
    memset(size, 0, 256);
    for ( int p = 0; p < cnt; p++ ) {
        code[vals[p]] = c[p];
        size[vals[p]] = s[p];
    }
}

/* ------------------------------------------------------------------------------------------- */
/* colour transforms (integer)  [ref: src/gpujpeg_colorspace.h:52-57, 64-101, 251-283]            */

static inline uint8_t clamp8(int v) { return (uint8_t)(v < 0 ? 0 : v > 255 ? 255 : v); }

/* [ref: src/gpujpeg_preprocessor.cu:163-201 + gpujpeg_colorspace.h:251-266]
 * padding outside the image stays 0 [ref: src/gpujpeg_common.c:941-944] */
void orc_preprocess_rgb444(const uint8_t* rgb, int w, int h, int pad, uint8_t* planes, int dw, int dh)
{
    size_t psz = (size_t)dw * dh;
    if ( dw != w || dh != h ) memset(planes, 0, 3 * psz); /* padding is 0; skip the pass when there is none */
    size_t pitch = (size_t)3 * w + pad;
#pragma omp parallel for schedule(static)
    for ( int y = 0; y < h; y++ ) {
        const uint8_t* s = rgb + (size_t)y * pitch;
        uint8_t* d = planes + (size_t)y * dw;
        for ( int x = 0; x < w; x++ ) {
            int r = (int)s[3 * x + 0] * 256 / 255;
            int g = (int)s[3 * x + 1] * 256 / 255;
            int b = (int)s[3 * x + 2] * 256 / 255;
            d[x] = clamp8(((77 * r + 150 * g + 29 * b + 128) >> 8) + 0);
            d[psz + x] = clamp8(((-43 * r - 85 * g + 128 * b + 128) >> 8) + 128);
            d[2 * psz + x] = clamp8(((128 * r - 107 * g - 21 * b + 128) >> 8) + 128);
        }
    }
}

/* [ref: src/gpujpeg_postprocessor.cu:183-216 + gpujpeg_colorspace.h:86-101, 268-283] */
void orc_postprocess_rgb444(const uint8_t* planes, int dw, int dh, uint8_t* rgb, int w, int h, int pad)
{
    size_t psz = (size_t)dw * dh;
    size_t pitch = (size_t)3 * w + pad;
#pragma omp parallel for schedule(static)
    for ( int y = 0; y < h; y++ ) {
        const uint8_t* s = planes + (size_t)y * dw;
        uint8_t* d = rgb + (size_t)y * pitch;
        for ( int x = 0; x < w; x++ ) {
            int yy = ((int)s[x] - 0) * 256 / 255;
            int cb = ((int)s[psz + x] - 128) * 256 / 255; /* C division: truncates toward zero */
            int cr = ((int)s[2 * psz + x] - 128) * 256 / 255;
            d[3 * x + 0] = clamp8((256 * yy + 0 * cb + 359 * cr + 128) >> 8);
            d[3 * x + 1] = clamp8((256 * yy - 88 * cb - 183 * cr + 128) >> 8);
            d[3 * x + 2] = clamp8((256 * yy + 454 * cb + 0 * cr + 128) >> 8);
        }
    }
}

/* ------------------------------------------------------------------------------------------- */
/* forward DCT + quantisation (float32)                                                          */

/* One 8-point AAN pass, op-for-op as the reference CUDA kernel executes it after nvcc's FMA
 * contraction.  [ref: src/gpujpeg_dct_gpu.cu:121-161] */
static inline void fdct1(const float in[8], float out[8], float shift)
{
    float d0 = in[0] + in[7], d1 = in[1] + in[6], d2 = in[2] + in[5], d3 = in[3] + in[4];
    float d4 = in[3] - in[4], d5 = in[2] - in[5], d6 = in[1] - in[6], d7 = in[0] - in[7];
    float e0 = d0 + d3, e1 = d1 + d2, e2 = d1 - d2, e3 = d0 - d3;
    float ed = e2 + e3;
    float o0 = d4 + d5, o1 = d5 + d6, o2 = d6 + d7;
    float od5 = (o0 - o2) * 0.382683433f;
    float od4 = fmaf(1.306562965f, o2, od5);
    float od3 = fmaf(-0.707106781f, o1, d7);
    float od2 = fmaf(0.541196100f, o0, od5);
    float od1 = fmaf(0.707106781f, o1, d7);
    out[0] = (e0 + e1) + shift;
    out[4] = e0 - e1;
    out[2] = fmaf(ed, 0.707106781f, e3);
    out[6] = fmaf(ed, -0.707106781f, e3);
    out[1] = od1 + od4;
    out[7] = od1 - od4;
    out[3] = od3 - od2;
    out[5] = od3 + od2;
}

/* [ref: src/gpujpeg_dct_gpu.cu:231-293]: columns first (level shift -1024 folded into the column
 * DC), rows second, then q = rint(c * fwd[x*8+y]), natural row-major store. */
void orc_fdct_quant_block(const uint8_t* px, int stride, const float fwd[64], int16_t out[64])
{
    float col[8][8]; /* col[v][x]: vertical frequency v at column x */
    for ( int x = 0; x < 8; x++ ) {
        float in[8], o[8];
        for ( int y = 0; y < 8; y++ )
            in[y] = (float)px[y * stride + x];
        fdct1(in, o, -1024.0f);
        for ( int v = 0; v < 8; v++ )
            col[v][x] = o[v];
    }
    for ( int v = 0; v < 8; v++ ) {
        float o[8];
        fdct1(col[v], o, 0.0f);
        for ( int u = 0; u < 8; u++ ) {
            int q = (int)rintf(o[u] * fwd[u * 8 + v]);
            out[v * 8 + u] = (int16_t)q;
        }
    }
}

void orc_fdct_quant_plane(const uint8_t* plane, int dw, int dh, const float fwd[64], int16_t* coef)
{
    int bcx = dw / 8, bcy = dh / 8;
#pragma omp parallel for schedule(static)
    for ( int by = 0; by < bcy; by++ )
        for ( int bx = 0; bx < bcx; bx++ )
            orc_fdct_quant_block(plane + (size_t)by * 8 * dw + bx * 8, dw, fwd,
                                 coef + ((size_t)by * bcx + bx) * 64);
}

/* ------------------------------------------------------------------------------------------- */
/* inverse DCT, integer flavour = the reference's gpujpeg_idct_cpu                                */

#define W1 2841
#define W2 2676
#define W3 2408
#define W5 1609
#define W6 1108
#define W7 565

static inline int iclip(int v) { return v < -256 ? -256 : v > 255 ? 255 : v; }

/* [ref: src/gpujpeg_dct_cpu.c:55-107] (the all-zero shortcut there is arithmetically identical to
 * the general path, so it is not restated) */
static inline void idct_row(int16_t* b)
{
    int x0 = ((int)b[0] << 11) + 128, x1 = (int)b[4] << 11, x2 = b[6], x3 = b[2], x4 = b[1], x5 = b[7],
        x6 = b[5], x7 = b[3], x8;
    x8 = W7 * (x4 + x5);
    x4 = x8 + (W1 - W7) * x4;
    x5 = x8 - (W1 + W7) * x5;
    x8 = W3 * (x6 + x7);
    x6 = x8 - (W3 - W5) * x6;
    x7 = x8 - (W3 + W5) * x7;
    x8 = x0 + x1;
    x0 -= x1;
    x1 = W6 * (x3 + x2);
    x2 = x1 - (W2 + W6) * x2;
    x3 = x1 + (W2 - W6) * x3;
    x1 = x4 + x6;
    x4 -= x6;
    x6 = x5 + x7;
    x5 -= x7;
    x7 = x8 + x3;
    x8 -= x3;
    x3 = x0 + x2;
    x0 -= x2;
    x2 = (181 * (x4 + x5) + 128) >> 8;
    x4 = (181 * (x4 - x5) + 128) >> 8;
    b[0] = (int16_t)((x7 + x1) >> 8);
    b[1] = (int16_t)((x3 + x2) >> 8);
    b[2] = (int16_t)((x0 + x4) >> 8);
    b[3] = (int16_t)((x8 + x6) >> 8);
    b[4] = (int16_t)((x8 - x6) >> 8);
    b[5] = (int16_t)((x0 - x4) >> 8);
    b[6] = (int16_t)((x3 - x2) >> 8);
    b[7] = (int16_t)((x7 - x1) >> 8);
}

/* [ref: src/gpujpeg_dct_cpu.c:119-171]; the reference indexes a 1024-entry clip table with
 * (x>>14): inside [-512,511] that equals iclip(); outside the reference reads out of bounds. */
static inline void idct_col(int16_t* b)
{
    int x0 = ((int)b[0] << 8) + 8192, x1 = (int)b[32] << 8, x2 = b[48], x3 = b[16], x4 = b[8], x5 = b[56],
        x6 = b[40], x7 = b[24], x8;
    x8 = W7 * (x4 + x5) + 4;
    x4 = (x8 + (W1 - W7) * x4) >> 3;
    x5 = (x8 - (W1 + W7) * x5) >> 3;
    x8 = W3 * (x6 + x7) + 4;
    x6 = (x8 - (W3 - W5) * x6) >> 3;
    x7 = (x8 - (W3 + W5) * x7) >> 3;
    x8 = x0 + x1;
    x0 -= x1;
    x1 = W6 * (x3 + x2) + 4;
    x2 = (x1 - (W2 + W6) * x2) >> 3;
    x3 = (x1 + (W2 - W6) * x3) >> 3;
    x1 = x4 + x6;
    x4 -= x6;
    x6 = x5 + x7;
    x5 -= x7;
    x7 = x8 + x3;
    x8 -= x3;
    x3 = x0 + x2;
    x0 -= x2;
    x2 = (181 * (x4 + x5) + 128) >> 8;
    x4 = (181 * (x4 - x5) + 128) >> 8;
    b[0] = (int16_t)iclip((x7 + x1) >> 14);
    b[8] = (int16_t)iclip((x3 + x2) >> 14);
    b[16] = (int16_t)iclip((x0 + x4) >> 14);
    b[24] = (int16_t)iclip((x8 + x6) >> 14);
    b[32] = (int16_t)iclip((x8 - x6) >> 14);
    b[40] = (int16_t)iclip((x0 - x4) >> 14);
    b[48] = (int16_t)iclip((x3 - x2) >> 14);
    b[56] = (int16_t)iclip((x7 - x1) >> 14);
}

/* [ref: src/gpujpeg_dct_cpu.c:178-189] dequantise (wraps to int16 like the reference), 8 rows, 8 cols */
void orc_idct_int_block(int16_t blk[64], const uint16_t inv[64])
{
    for ( int i = 0; i < 64; i++ )
        blk[i] = (int16_t)((int)blk[i] * (int)(int16_t)inv[i]);
    for ( int i = 0; i < 8; i++ )
        idct_row(blk + 8 * i);
    for ( int i = 0; i < 8; i++ )
        idct_col(blk + i);
}

/* ------------------------------------------------------------------------------------------- */
/* inverse DCT, float flavour = what the reference CUDA kernel computes     */

/* [ref: src/gpujpeg_dct_gpu.cu:312-363] executed strictly in source order; a*b+c patterns are the
 * FFMAs nvcc forms, products inside a sum of two products are rounded first (FMUL) */
static inline void idct1_float(float* V)
{
    const float k0 = 0.4142135623f, k1 = 0.3535533905f, k2 = 0.4619397662f, k3 = 0.1989123673f,
                k4 = 0.7071067811f;
    V[2] = V[2] * 0.5411961f;
    V[4] = V[4] * 0.509795579f;
    V[5] = V[5] * 0.601344887f;
    V[1] = (V[0] - V[1]) * k1;
    V[0] = fmaf(V[0], k4, -V[1]);
    V[3] = fmaf(V[2], k1, V[3] * k2);
    V[2] = fmaf(V[3], k0, -V[2]);
    V[6] = fmaf(V[5], k2, V[6] * k0);
    V[5] = fmaf(-0.6681786379f, V[6], V[5]);
    V[7] = fmaf(V[4], k3, V[7] * 0.49039264f);
    V[4] = fmaf(V[7], k3, -V[4]);
    V[1] = V[2] + V[1];
    V[2] = fmaf(-2.0f, V[2], V[1]);
    V[4] = V[5] + V[4];
    V[5] = fmaf(2.0f, V[5], -V[4]);
    V[7] = V[6] + V[7];
    V[6] = fmaf(-2.0f, V[6], V[7]);
    V[0] = V[3] + V[0];
    V[3] = fmaf(-2.0f, V[3], V[0]);
    V[5] = fmaf(V[6], k0, V[5]);
    V[6] = fmaf(V[5], -k4, V[6]);
    V[5] = fmaf(V[6], k0, V[5]);
    V[3] = V[3] + V[4];
    V[4] = fmaf(-2.0f, V[4], V[3]);
    V[2] = V[2] + V[5];
    V[5] = fmaf(-2.0f, V[5], V[2]);
    V[1] = V[6] + V[1];
    V[6] = fmaf(-2.0f, V[6], V[1]);
    V[0] = V[0] + V[7];
    V[7] = fmaf(-2.0f, V[7], V[0]);
}

/* [ref: src/gpujpeg_dct_gpu.cu:497-501, 532-550, 581-617]: dequantise in int, permute inputs
 * {0,4,6,2,7,5,3,1}, columns then rows, pixel = clamp(rint(v+128)) */
void orc_idct_float_block(const int16_t in[64], const uint16_t inv[64], uint8_t out[64])
{
    static const int perm[8] = {0, 4, 6, 2, 7, 5, 3, 1};
    float f[64];
    for ( int i = 0; i < 64; i++ )
        f[i] = (float)((int)in[i] * (int)inv[i]);
    for ( int x = 0; x < 8; x++ ) {
        float V[8];
        for ( int k = 0; k < 8; k++ )
            V[k] = f[perm[k] * 8 + x];
        idct1_float(V);
        for ( int k = 0; k < 8; k++ )
            f[k * 8 + x] = V[k];
    }
    for ( int y = 0; y < 8; y++ ) {
        float V[8];
        for ( int k = 0; k < 8; k++ )
            V[k] = f[y * 8 + perm[k]];
        idct1_float(V);
        for ( int k = 0; k < 8; k++ ) {
            int p = (int)rintf(V[k] + 128.0f);
            out[y * 8 + k] = clamp8(p);
        }
    }
}

/* [ref: src/gpujpeg_dct_cpu.c:202-257 (int flavour, incl. +128/clamp/de-block :239-251);
 *       src/gpujpeg_dct_gpu.cu:681-727 (float flavour)] */
void orc_idct_plane(const int16_t* coef, int dw, int dh, const uint16_t inv[64], int flavour, uint8_t* plane)
{
    int bcx = dw / 8, bcy = dh / 8;
#pragma omp parallel for schedule(static)
    for ( int by = 0; by < bcy; by++ ) {
        for ( int bx = 0; bx < bcx; bx++ ) {
            const int16_t* c = coef + ((size_t)by * bcx + bx) * 64;
            uint8_t* d = plane + (size_t)by * 8 * dw + bx * 8;
            if ( flavour == ORC_IDCT_FLOAT_GPUREF ) {
                uint8_t o[64];
                orc_idct_float_block(c, inv, o);
                for ( int i = 0; i < 64; i++ )
                    d[(i / 8) * dw + (i % 8)] = o[i];
            }
            else {
                int16_t b[64];
                memcpy(b, c, sizeof b);
                orc_idct_int_block(b, inv);
                for ( int i = 0; i < 64; i++ ) {
                    int16_t v = (int16_t)(b[i] + 128);
                    d[(i / 8) * dw + (i % 8)] = (uint8_t)(v > 255 ? 255 : v < 0 ? 0 : v);
                }
            }
        }
    }
}

/* ------------------------------------------------------------------------------------------- */
/* Huffman encoding of one restart segment                                  */

struct bitw {
    uint8_t* p;
    uint32_t acc; /* pending bits, right aligned */
    int n;        /* number of pending bits (<8 between calls) */
};

/* MSB-first bit emission with 0xFF byte stuffing  [ref: src/gpujpeg_huffman_cpu_encoder.c:72-107] */
static inline void put_bits(struct bitw* w, unsigned code, int size)
{
    w->acc = (w->acc << size) | (code & ((1u << size) - 1u));
    w->n += size;
    while ( w->n >= 8 ) {
        uint8_t b = (uint8_t)(w->acc >> (w->n - 8));
        *w->p++ = b;
        if ( b == 0xFF ) *w->p++ = 0;
        w->n -= 8;
    }
}

static inline int bit_length(int v)
{
    int n = 0;
    while ( v ) {
        n++;
        v >>= 1;
    }
    return n;
}

struct enc_tab {
    uint16_t code[256];
    uint8_t size[256];
};
static struct enc_tab g_enc[2][2];
static void enc_tables_init(void)
{
#pragma omp critical(orc_enc_tables)
    {
        if ( !g_enc_ready ) {
            for ( int c = 0; c < 2; c++ )
                for ( int k = 0; k < 2; k++ )
                    orc_huff_encoder_table(c, k, g_enc[c][k].code, g_enc[c][k].size);
            g_enc_ready = 1;
        }
    }
}

/* one block  [ref: src/gpujpeg_huffman_cpu_encoder.c:135-227] */
static inline void encode_block(struct bitw* w, const int16_t* blk, int* pred, const struct enc_tab* dc,
                                const struct enc_tab* ac)
{
    int diff = blk[0] - *pred;
    *pred = blk[0];
    int mag = diff < 0 ? -diff : diff;
    int bits = diff < 0 ? diff - 1 : diff;
    int n = bit_length(mag);
    put_bits(w, dc->code[n], dc->size[n]);
    if ( n ) put_bits(w, (unsigned)bits, n);
    int run = 0;
    for ( int k = 1; k < 64; k++ ) {
        int v = blk[orc_zigzag_to_natural[k]];
        if ( v == 0 ) {
            run++;
            continue;
        }
        while ( run > 15 ) {
            put_bits(w, ac->code[0xF0], ac->size[0xF0]);
            run -= 16;
        }
        mag = v < 0 ? -v : v;
        bits = v < 0 ? v - 1 : v;
        n = bit_length(mag);
        int sym = (run << 4) + n;
        put_bits(w, ac->code[sym], ac->size[sym]);
        put_bits(w, (unsigned)bits, n);
        run = 0;
    }
    if ( run > 0 ) put_bits(w, ac->code[0], ac->size[0]);
}

/* segment end: pad with 1-bits to a byte boundary  [ref: src/gpujpeg_huffman_cpu_encoder.c:115-128] */
static inline void flush_bits(struct bitw* w)
{
    if ( w->n > 0 ) {
        put_bits(w, 0x7F, 7);
        w->acc = 0;
        w->n = 0;
    }
}

size_t orc_huff_encode_segment(const int16_t* coef, int nblocks, int cls, uint8_t* out)
{
    if ( !g_enc_ready ) enc_tables_init();
    struct bitw w = {out, 0, 0};
    int pred = 0; /* [ref: src/gpujpeg_huffman_cpu_encoder.c:361-364] predictor restarts per segment */
    for ( int b = 0; b < nblocks; b++ )
        encode_block(&w, coef + (size_t)b * 64, &pred, &g_enc[cls][0], &g_enc[cls][1]);
    flush_bits(&w);
    return (size_t)(w.p - out);
}

/* ------------------------------------------------------------------------------------------- */
/* header writer  [ref: src/gpujpeg_writer.c:120-156 (APP0), 282-303 (DQT), 319-356 (SOF0),
 *                 366-411 (DHT), 419-428 (DRI), 431-453 (COM), 456-518 (order)]                   */

static inline uint8_t* put8(uint8_t* p, int v) { *p++ = (uint8_t)v; return p; }
static inline uint8_t* put16(uint8_t* p, int v) { *p++ = (uint8_t)(v >> 8); *p++ = (uint8_t)v; return p; }
static inline uint8_t* putm(uint8_t* p, int m) { *p++ = 0xFF; *p++ = (uint8_t)m; return p; }

/* RGB-internal streams: Adobe APP14 (transform 0) instead of JFIF, component ids 'R','G','B', every component coded with
 * the luminance tables [ref: src/gpujpeg_writer.c:255-276, 305-313, 462-470; src/gpujpeg_common.c:689-692] */
static int g_rgb_internal = 0;
/* limited-range YCbCr internal colour spaces: SPIFF header naming the space (T.84 code 4 = BT.601, 1 = BT.709), a
 * second SOI behind the end-of-directory entry, and a "CS=ITU601" comment for BT.601
 * [ref: src/gpujpeg_writer.c:171-245, 462-466, 513-515] */
static int g_spiff_cs = 0;

/* table class of component c: luminance for the first and for a fourth (alpha) component, and for every component of an
 * RGB-internal JPEG [ref: src/gpujpeg_common.c:692-694] */
static int comp_class(int c) { return (c == 0 || c == 3 || g_rgb_internal) ? 0 : 1; }
/* component ids of SOF0 / SOS [ref: src/gpujpeg_writer.c:305-313] */
static int comp_ident(int c) { return g_rgb_internal ? "RGBA"[c] : c + 1; }

size_t orc_write_header(uint8_t* out, int w, int h, int quality, int rst, int comp_count)
{
    uint8_t raw[2][64];
    orc_quant_tables(quality, raw, NULL, NULL);
    uint8_t* p = out;
    p = putm(p, 0xD8);
    /* four components are described by a SPIFF header whatever the colour space [ref: src/gpujpeg_writer.c:458-460] */
    const int spiff_cs = g_spiff_cs ? g_spiff_cs : comp_count == 4 ? (g_rgb_internal ? 10 : 3) : 0;
    if ( spiff_cs ) {
        p = putm(p, 0xE8);
        p = put16(p, 32);
        memcpy(p, "SPIFF", 6);
        p += 6;
        p = put16(p, 0x100);
        p = put8(p, spiff_cs == 3 ? 1 : 0);   /* profile [ref: src/gpujpeg_writer.c:203] */
        p = put8(p, comp_count);
        p = put16(p, 0); p = put16(p, h);
        p = put16(p, 0); p = put16(p, w);
        p = put8(p, spiff_cs);
        p = put8(p, 8); p = put8(p, 5); p = put8(p, 0);
        p = put16(p, 0); p = put16(p, 1);
        p = put16(p, 0); p = put16(p, 1);
        p = putm(p, 0xE8);
        p = put16(p, 8);
        p = put16(p, 0); p = put16(p, 1);
        p = putm(p, 0xD8);
    }
    else if ( g_rgb_internal ) {
        p = putm(p, 0xEE);
        p = put16(p, 14);
        memcpy(p, "Adobe", 5);
        p += 5;
        p = put16(p, 100); p = put16(p, 0); p = put16(p, 0);
        p = put8(p, 0);
    }
    else {
        p = putm(p, 0xE0);
        p = put16(p, 16);
        memcpy(p, "JFIF", 5);
        p += 5;
        p = put8(p, 1); p = put8(p, 1); p = put8(p, 1);
        p = put16(p, 300); p = put16(p, 300);
        p = put8(p, 0); p = put8(p, 0);
    }
    int ntypes = (comp_count > 1 && !g_rgb_internal) ? 2 : 1;
    for ( int t = 0; t < ntypes; t++ ) {
        p = putm(p, 0xDB);
        p = put16(p, 67);
        p = put8(p, t);
        memcpy(p, raw[t], 64);
        p += 64;
    }
    p = putm(p, 0xC0);
    p = put16(p, 8 + 3 * comp_count);
    p = put8(p, 8);
    p = put16(p, h);
    p = put16(p, w);
    p = put8(p, comp_count);
    for ( int c = 0; c < comp_count; c++ ) {
        p = put8(p, comp_ident(c));
        p = put8(p, 0x11);
        p = put8(p, comp_class(c));
    }
    for ( int t = 0; t < ntypes; t++ ) {
        for ( int k = 0; k < 2; k++ ) {
            const uint8_t *bits, *vals;
            int n;
            orc_huff_spec(t, k, &bits, &vals, &n);
            p = putm(p, 0xC4);
            p = put16(p, n + 2 + 1 + 16);
            p = put8(p, (k << 4) | t);
            memcpy(p, bits + 1, 16);
            p += 16;
            memcpy(p, vals, n);
            p += n;
        }
    }
    p = putm(p, 0xDD);
    p = put16(p, 4);
    p = put16(p, rst);
    char com[64];
    int q = quality < 1 ? 1 : quality > 100 ? 100 : quality;
    int len = 0;
    {
        const char* pre = "CREATOR: GPUJPEG, quality = ";
        len = (int)strlen(pre);
        memcpy(com, pre, len);
        char num[8];
        int nd = 0;
        do {
            num[nd++] = (char)('0' + q % 10);
            q /= 10;
        } while ( q );
        while ( nd )
            com[len++] = num[--nd];
        com[len] = 0;
    }
    p = putm(p, 0xFE);
    p = put16(p, 2 + len + 1);
    memcpy(p, com, len + 1);
    p += len + 1;
    if ( g_spiff_cs == 4 ) {
        p = putm(p, 0xFE);
        p = put16(p, 12);
        memcpy(p, "CS=ITU601", 10);
        p += 10;
    }
    return (size_t)(p - out);
}

/* "segment info": in front of every SOS the encoder can place APP13 headers holding the byte position of every restart
 * segment inside the scan (and the scan's end), big-endian 32 bit, relative to the first byte behind the SOS header; a
 * header carries at most 65436 bytes of positions, so long tables are cut into several headers
 * [ref: src/gpujpeg_writer.c:553-599 (headers), :522-546 (positions), src/gpujpeg_encoder.c:575-621 (one position in
 * front of every segment, one more behind the scan's last segment, whose RSTn is dropped)] */
static int g_segment_info = 0;
void orc_set_segment_info(int on) { g_segment_info = on; }
#define ORC_MAX_HEADER_SIZE (65536 - 100)   /* [ref: src/gpujpeg_common_internal.h:91] */
#define ORC_MAX_SEGINFO_HEADERS 256
struct seginfo {
    uint8_t* block[ORC_MAX_SEGINFO_HEADERS];
    int count;
};
static uint8_t* write_segment_info_headers(uint8_t* p, int scan_index, int segment_count, struct seginfo* si)
{
    int data_size = (segment_count + 1) * 4;
    si->count = 0;
    while ( data_size > 0 && si->count < ORC_MAX_SEGINFO_HEADERS ) {
        const int header_size = data_size > ORC_MAX_HEADER_SIZE ? ORC_MAX_HEADER_SIZE : data_size;
        data_size -= header_size;
        p = putm(p, 0xED);
        p = put16(p, 3 + header_size);
        p = put8(p, scan_index);
        si->block[si->count++] = p;
        memset(p, 0, (size_t)header_size);
        p += header_size;
    }
    return p;
}
static void put_segment_position(struct seginfo* si, int index, size_t position)
{
    const int h = (index * 4) / ORC_MAX_HEADER_SIZE, d = (index * 4) % ORC_MAX_HEADER_SIZE;
    if ( h >= si->count ) return;
    si->block[h][d] = (uint8_t)(position >> 24);
    si->block[h][d + 1] = (uint8_t)(position >> 16);
    si->block[h][d + 2] = (uint8_t)(position >> 8);
    si->block[h][d + 3] = (uint8_t)position;
}

/* scan header  [ref: src/gpujpeg_writer.c:600-658] */
static uint8_t* write_sos(uint8_t* p, int interleaved, int comp_count, int scan_comp)
{
    p = putm(p, 0xDA);
    if ( interleaved ) {
        p = put16(p, 6 + 2 * comp_count);
        p = put8(p, comp_count);
        for ( int c = 0; c < comp_count; c++ ) {
            p = put8(p, comp_ident(c));
            p = put8(p, comp_class(c) ? 0x11 : 0x00);
        }
    }
    else {
        p = put16(p, 8);
        p = put8(p, 1);
        p = put8(p, comp_ident(scan_comp));
        p = put8(p, comp_class(scan_comp) ? 0x11 : 0x00);
    }
    p = put8(p, 0);
    p = put8(p, 0x3F);
    p = put8(p, 0);
    return p;
}

/* ------------------------------------------------------------------------------------------- */
/* whole-frame encode                                                                            */

/* [ref: src/gpujpeg_encoder.c:351-646 with the CPU Huffman path :511-534 /
 *       src/gpujpeg_huffman_cpu_encoder.c:296-376 for the scan/segment/RST structure] */
size_t orc_encode_rgb(const uint8_t* rgb, int w, int h, int pad, int quality, int rst, int interleaved,
                      int threads, uint8_t* out, int16_t* coef_out)
{
    if ( w <= 0 || h <= 0 || w > 65535 || h > 65535 || rst < 0 || rst > 65535 ) return 0;
    enc_tables_init();
#ifdef _OPENMP
    int saved_threads = omp_get_max_threads();
    omp_set_num_threads(threads > 1 ? threads : 1);
#else
    (void)threads;
#endif
    const int comps = 3;
    int dw = (w + 7) / 8 * 8, dh = (h + 7) / 8 * 8;
    size_t psz = (size_t)dw * dh;
    int nblk = (dw / 8) * (dh / 8);
    uint8_t raw[2][64];
    float fwd[2][64];
    orc_quant_tables(quality, raw, fwd, NULL);

    uint8_t* planes = (uint8_t*)scratch(0, 3 * psz);
    int16_t* coef = coef_out ? coef_out : (int16_t*)scratch(1, 3 * psz * sizeof(int16_t));
    orc_preprocess_rgb444(rgb, w, h, pad, planes, dw, dh);
    for ( int c = 0; c < comps; c++ )
        orc_fdct_quant_plane(planes + c * psz, dw, dh, fwd[c == 0 ? 0 : 1], coef + c * psz);

    uint8_t* p = out + orc_write_header(out, w, h, quality, rst, comps);

    int seg_mcu = rst > 0 ? rst : nblk;
    int nseg = (nblk + seg_mcu - 1) / seg_mcu;
    int nscan = interleaved ? 1 : comps;
    int blocks_per_mcu = interleaved ? comps : 1;
    size_t slot = (size_t)seg_mcu * blocks_per_mcu * 416 + 16;
    /* encode every segment of every scan independently (legal: segments share no state), then
     * concatenate in order with RSTn between segments of a scan */
    size_t total_seg = (size_t)nscan * nseg;
    size_t* seg_len = (size_t*)scratch(2, total_seg * sizeof(size_t));
    /* one worst-case slot per segment; only the used prefix of each slot is ever touched */
    uint8_t* tmp = (uint8_t*)scratch(3, total_seg * slot);
    if ( !seg_len || !tmp ) return 0;
#pragma omp parallel for schedule(dynamic, 16)
    for ( long long si = 0; si < (long long)total_seg; si++ ) {
        int scan = (int)(si / nseg), s = (int)(si % nseg);
        int first = s * seg_mcu;
        int cnt = nblk - first < seg_mcu ? nblk - first : seg_mcu;
        uint8_t* o = tmp + (size_t)si * slot;
        if ( !interleaved ) {
            seg_len[si] = orc_huff_encode_segment(coef + scan * psz + (size_t)first * 64, cnt, scan == 0 ? 0 : 1, o);
        }
        else {
            struct bitw bw = {o, 0, 0};
            int pred[3] = {0, 0, 0};
            for ( int m = 0; m < cnt; m++ )
                for ( int c = 0; c < comps; c++ )
                    encode_block(&bw, coef + c * psz + (size_t)(first + m) * 64, &pred[c],
                                 &g_enc[c == 0 ? 0 : 1][0], &g_enc[c == 0 ? 0 : 1][1]);
            flush_bits(&bw);
            seg_len[si] = (size_t)(bw.p - o);
        }
    }
    for ( int scan = 0; scan < nscan; scan++ ) {
        struct seginfo info;
        const int with_info = g_segment_info && rst > 0;
        if ( with_info ) p = write_segment_info_headers(p, scan, nseg, &info);
        p = write_sos(p, interleaved, comps, scan);
        const uint8_t* scan_start = p;
        for ( int s = 0; s < nseg; s++ ) {
            size_t si = (size_t)scan * nseg + s;
            if ( with_info ) put_segment_position(&info, s, (size_t)(p - scan_start));
            memcpy(p, tmp + si * slot, seg_len[si]);
            p += seg_len[si];
            if ( s + 1 < nseg ) p = putm(p, 0xD0 + (s & 7));
        }
        if ( with_info ) put_segment_position(&info, nseg, (size_t)(p - scan_start));
    }
    p = putm(p, 0xD9);
#ifdef _OPENMP
    omp_set_num_threads(saved_threads);
#endif
    return (size_t)(p - out);
}

/* ------------------------------------------------------------------------------------------- */
/* chroma subsampling: component geometry, point-sampling preprocessor, whole-frame encode          */

struct ogeo {
    int hs, vs;        /* sampling factors */
    int w, h;          /* samples carrying image data */
    int dw, dh;        /* allocated plane (multiple of the component's MCU size) */
    int bcx, bcy, nblk;
    size_t off;        /* offset of the plane (samples) == offset of its coefficients */
};

/* [ref: src/gpujpeg_common.c:671-736] */
static size_t ogeo_init(struct ogeo g[4], int comps, const int hs[4], const int vs[4], int w, int h, int interleaved,
                        int* max_hs, int* max_vs)
{
    int mh = 1, mv = 1;
    for ( int c = 0; c < comps; c++ ) {
        if ( hs[c] > mh ) mh = hs[c];
        if ( vs[c] > mv ) mv = vs[c];
    }
    size_t off = 0;
    for ( int c = 0; c < comps; c++ ) {
        int div_h = mh / hs[c], div_v = mv / vs[c];
        g[c].hs = hs[c];
        g[c].vs = vs[c];
        g[c].w = ((w + div_h - 1) / div_h * div_h) * hs[c] / mh;
        g[c].h = ((h + div_v - 1) / div_v * div_v) * vs[c] / mv;
        int mx = interleaved ? 8 * hs[c] : 8, my = interleaved ? 8 * vs[c] : 8;
        g[c].dw = (g[c].w + mx - 1) / mx * mx;
        g[c].dh = (g[c].h + my - 1) / my * my;
        g[c].bcx = g[c].dw / 8;
        g[c].bcy = g[c].dh / 8;
        g[c].nblk = g[c].bcx * g[c].bcy;
        g[c].off = off;
        off += (size_t)g[c].dw * g[c].dh;
    }
    *max_hs = mh;
    *max_vs = mv;
    return off;
}

/* block `i` (coding order) of MCU `m` of an interleaved scan -> pointer to its 64 coefficients
 * [ref: src/gpujpeg_common.c:1056-1085] */
static inline int16_t* mcu_block(int16_t* coef, const struct ogeo* g, int mcu_x, int m, int x, int y)
{
    int mx = m % mcu_x, my = m / mcu_x;
    size_t b = (size_t)(my * g->vs + y) * g->bcx + (size_t)mx * g->hs + x;
    return coef + g->off + b * 64;
}

/* RGB -> YCbCr planes; a component with divisor (dh,dv) keeps the sample of pixel (x,y) when x % dh == 0 and
 * y % dv == 0 -- no filtering [ref: src/gpujpeg_preprocessor.cu:50-64]; everything else stays 0 */
static void preprocess_rgb_ss(const uint8_t* rgb, int w, int h, int pad, uint8_t* planes, const struct ogeo g[3],
                              int max_hs, int max_vs, size_t total)
{
    memset(planes, 0, total);
    size_t pitch = (size_t)3 * w + pad;
#pragma omp parallel for schedule(static)
    for ( int y = 0; y < h; y++ ) {
        const uint8_t* s = rgb + (size_t)y * pitch;
        for ( int x = 0; x < w; x++ ) {
            int r = (int)s[3 * x + 0] * 256 / 255;
            int gg = (int)s[3 * x + 1] * 256 / 255;
            int b = (int)s[3 * x + 2] * 256 / 255;
            uint8_t v[3];
            v[0] = clamp8(((77 * r + 150 * gg + 29 * b + 128) >> 8) + 0);
            v[1] = clamp8(((-43 * r - 85 * gg + 128 * b + 128) >> 8) + 128);
            v[2] = clamp8(((128 * r - 107 * gg - 21 * b + 128) >> 8) + 128);
            for ( int c = 0; c < 3; c++ ) {
                int dh = max_hs / g[c].hs, dv = max_vs / g[c].vs;
                if ( x % dh || y % dv ) continue;
                planes[g[c].off + (size_t)(y / dv) * g[c].dw + x / dh] = v[c];
            }
        }
    }
}

/* planes -> RGB: every pixel takes the sample at (x / dh, y / dv) [ref: src/gpujpeg_postprocessor.cu:55-76] */
static void postprocess_rgb_ss(const uint8_t* planes, const struct ogeo g[3], int max_hs, int max_vs, uint8_t* rgb, int w,
                               int h, int pad)
{
    size_t pitch = (size_t)3 * w + pad;
#pragma omp parallel for schedule(static)
    for ( int y = 0; y < h; y++ ) {
        uint8_t* d = rgb + (size_t)y * pitch;
        for ( int x = 0; x < w; x++ ) {
            int v[3];
            for ( int c = 0; c < 3; c++ ) {
                int dh = max_hs / g[c].hs, dv = max_vs / g[c].vs;
                v[c] = planes[g[c].off + (size_t)(y / dv) * g[c].dw + x / dh];
            }
            int yy = (v[0] - 0) * 256 / 255;
            int cb = (v[1] - 128) * 256 / 255;
            int cr = (v[2] - 128) * 256 / 255;
            d[3 * x + 0] = clamp8((256 * yy + 0 * cb + 359 * cr + 128) >> 8);
            d[3 * x + 1] = clamp8((256 * yy - 88 * cb - 183 * cr + 128) >> 8);
            d[3 * x + 2] = clamp8((256 * yy + 454 * cb + 0 * cr + 128) >> 8);
        }
    }
}

/* [ref: src/gpujpeg_encoder.c:351-646; block order of interleaved MCUs src/gpujpeg_common.c:1056-1085;
 *       per-component scans src/gpujpeg_huffman_cpu_encoder.c:296-376] */
/* planes (already in the JPEG's component layout) -> file: FDCT + quantisation, header, scans.
 * [ref: src/gpujpeg_encoder.c:351-646; block order of interleaved MCUs src/gpujpeg_common.c:1056-1085;
 *       per-component scans src/gpujpeg_huffman_cpu_encoder.c:296-376] */
static size_t encode_from_planes(const uint8_t* planes, const struct ogeo g[4], int comps, const int hs[4], const int vs[4],
                                 int w, int h, int quality, int rst, int interleaved, uint8_t* out, int16_t* coef)
{
    uint8_t raw[2][64];
    float fwd[2][64];
    orc_quant_tables(quality, raw, fwd, NULL);
    for ( int c = 0; c < comps; c++ )
        orc_fdct_quant_plane(planes + g[c].off, g[c].dw, g[c].dh, fwd[comp_class(c)], coef + g[c].off);

    /* header: as orc_write_header, with the sampling factors patched into SOF0 */
    size_t hl = orc_write_header(out, w, h, quality, rst, comps);
    for ( size_t i = 0; i + 9 < hl; i++ )
        if ( out[i] == 0xFF && out[i + 1] == 0xC0 ) {
            for ( int c = 0; c < comps; c++ )
                out[i + 2 + 8 + 3 * c + 1] = (uint8_t)((hs[c] << 4) | vs[c]);
            break;
        }
    uint8_t* p = out + hl;

    if ( comps == 1 ) interleaved = 0;
    int nscan = interleaved ? 1 : comps;
    int bpm = 0;
    for ( int c = 0; c < comps; c++ )
        bpm += hs[c] * vs[c];
    int mcu_x = g[0].bcx / g[0].hs;
    int scan_mcus[4], scan_seg[4], seg_begin[5] = {0, 0, 0, 0, 0};
    int max_mcus = 0;
    for ( int k = 0; k < nscan; k++ ) {
        scan_mcus[k] = interleaved ? mcu_x * (g[0].bcy / g[0].vs) : g[k].nblk;
        if ( scan_mcus[k] > max_mcus ) max_mcus = scan_mcus[k];
    }
    int seg_mcu = rst > 0 ? rst : max_mcus;
    for ( int k = 0; k < nscan; k++ ) {
        /* rst == 0: one segment per scan [ref: src/gpujpeg_common.c:723-729] */
        scan_seg[k] = rst > 0 ? (scan_mcus[k] + seg_mcu - 1) / seg_mcu : 1;
        seg_begin[k + 1] = seg_begin[k] + scan_seg[k];
    }
    size_t total_seg = (size_t)seg_begin[nscan];
    size_t slot = (size_t)seg_mcu * (interleaved ? bpm : 1) * 416 + 16;
    size_t* seg_len = (size_t*)scratch(2, total_seg * sizeof(size_t));
    uint8_t* tmp = (uint8_t*)scratch(3, total_seg * slot);
    if ( !seg_len || !tmp ) return 0;
#pragma omp parallel for schedule(dynamic, 16)
    for ( long long si = 0; si < (long long)total_seg; si++ ) {
        int scan = 0;
        while ( scan + 1 < nscan && si >= seg_begin[scan + 1] )
            scan++;
        int s = (int)(si - seg_begin[scan]);
        int first = rst > 0 ? s * seg_mcu : 0;
        int cnt = rst > 0 ? (scan_mcus[scan] - first < seg_mcu ? scan_mcus[scan] - first : seg_mcu) : scan_mcus[scan];
        uint8_t* o = tmp + (size_t)si * slot;
        if ( !interleaved ) {
            seg_len[si] = orc_huff_encode_segment(coef + g[scan].off + (size_t)first * 64, cnt,
                                                  comp_class(scan), o);
        }
        else {
            struct bitw bw = {o, 0, 0};
            int pred[4] = {0, 0, 0, 0};
            for ( int m = first; m < first + cnt; m++ )
                for ( int c = 0; c < comps; c++ ) {
                    const int cls = comp_class(c);
                    for ( int y = 0; y < g[c].vs; y++ )
                        for ( int x = 0; x < g[c].hs; x++ )
                            encode_block(&bw, mcu_block(coef, &g[c], mcu_x, m, x, y), &pred[c], &g_enc[cls][0], &g_enc[cls][1]);
                }
            flush_bits(&bw);
            seg_len[si] = (size_t)(bw.p - o);
        }
    }
    for ( int scan = 0; scan < nscan; scan++ ) {
        struct seginfo info;
        const int with_info = g_segment_info && rst > 0;
        if ( with_info ) p = write_segment_info_headers(p, scan, scan_seg[scan], &info);
        p = write_sos(p, interleaved, comps, scan);
        const uint8_t* scan_start = p;
        for ( int s = 0; s < scan_seg[scan]; s++ ) {
            size_t si = (size_t)seg_begin[scan] + s;
            if ( with_info ) put_segment_position(&info, s, (size_t)(p - scan_start));
            memcpy(p, tmp + si * slot, seg_len[si]);
            p += seg_len[si];
            if ( s + 1 < scan_seg[scan] ) p = putm(p, 0xD0 + (s & 7));
        }
        if ( with_info ) put_segment_position(&info, scan_seg[scan], (size_t)(p - scan_start));
    }
    p = putm(p, 0xD9);
    return (size_t)(p - out);
}

size_t orc_encode_rgb_ss(const uint8_t* rgb, int w, int h, int pad, int quality, int rst, int interleaved, int lhs,
                         int lvs, int threads, uint8_t* out, int16_t* coef_out)
{
    if ( w <= 0 || h <= 0 || w > 65535 || h > 65535 || rst < 0 || rst > 65535 ) return 0;
    if ( lhs < 1 || lvs < 1 || lhs > 2 || lvs > 2 ) return 0;
    enc_tables_init();
#ifdef _OPENMP
    int saved_threads = omp_get_max_threads();
    omp_set_num_threads(threads > 1 ? threads : 1);
#else
    (void)threads;
#endif
    const int comps = 3;
    const int hs[4] = {lhs, 1, 1, 0}, vs[4] = {lvs, 1, 1, 0};
    struct ogeo g[4];
    int max_hs, max_vs;
    size_t total = ogeo_init(g, comps, hs, vs, w, h, interleaved, &max_hs, &max_vs);
    uint8_t* planes = (uint8_t*)scratch(0, total);
    int16_t* coef = coef_out ? coef_out : (int16_t*)scratch(1, total * sizeof(int16_t));
    preprocess_rgb_ss(rgb, w, h, pad, planes, g, max_hs, max_vs, total);
    size_t n = encode_from_planes(planes, g, comps, hs, vs, w, h, quality, rst, interleaved, out, coef);
#ifdef _OPENMP
    omp_set_num_threads(saved_threads);
#endif
    return n;
}

/* ------------------------------------------------------------------------------------------- */
/* raw pixel formats whose samples go into the JPEG without a colour transform (image colour space == internal
 * colour space): where the samples of each component live [ref: src/gpujpeg_preprocessor.cu:78-160 (loads),
 * :409-455 (planar copy), src/gpujpeg_common.c:2075-2107 (format descriptions)]                          */

struct rawcomp {
    size_t off;    /* byte offset of sample (0,0) */
    int xs;        /* bytes between horizontally adjacent samples */
    size_t pitch;  /* bytes between rows */
};

/* fmt: the reference's enum gpujpeg_pixel_format values (0 u8, 1 444-p012, 2 444-p0p1p2, 3 422-p1020, 4 422-p0p1p2,
 * 5 420-p0p1p2).  Returns the component count, 0 if unsupported. */
static int raw_layout(int fmt, int w, int h, int pad, struct rawcomp rc[3], int hs[4], int vs[4], size_t* size)
{
    int cw = (w + 1) / 2, ch = (h + 1) / 2;
    for ( int c = 0; c < 4; c++ )
        hs[c] = vs[c] = c < 3 ? 1 : 0;
    switch ( fmt ) {
        case 0:
            rc[0] = (struct rawcomp){0, 1, (size_t)w + pad};
            hs[1] = vs[1] = hs[2] = vs[2] = 0;
            *size = ((size_t)w + pad) * h;
            return 1;
        case 1:
            for ( int c = 0; c < 3; c++ )
                rc[c] = (struct rawcomp){(size_t)c, 3, (size_t)3 * w + pad};
            *size = ((size_t)3 * w + pad) * h;
            return 3;
        case 6:   /* 4444-u8-p0123 as a 3-component image: alpha ignored on input, 255 on output */
            for ( int c = 0; c < 3; c++ )
                rc[c] = (struct rawcomp){(size_t)c, 4, (size_t)4 * w + pad};
            *size = ((size_t)4 * w + pad) * h;
            return 3;
        case 2:
            for ( int c = 0; c < 3; c++ )
                rc[c] = (struct rawcomp){(size_t)c * ((size_t)w + pad) * h, 1, (size_t)w + pad};
            *size = 3 * ((size_t)w + pad) * h;
            return 3;
        case 3: {
            int we = (w + 1) & ~1;
            size_t pitch = (size_t)2 * we + pad;
            rc[0] = (struct rawcomp){1, 2, pitch};
            rc[1] = (struct rawcomp){0, 4, pitch};
            rc[2] = (struct rawcomp){2, 4, pitch};
            hs[0] = 2;
            *size = pitch * h;
            return 3;
        }
        case 4:
            rc[0] = (struct rawcomp){0, 1, (size_t)w + pad};
            rc[1] = (struct rawcomp){((size_t)w + pad) * h, 1, (size_t)cw + pad};
            rc[2] = (struct rawcomp){((size_t)w + pad) * h + ((size_t)cw + pad) * h, 1, (size_t)cw + pad};
            hs[0] = 2;
            *size = ((size_t)w + pad) * h + 2 * ((size_t)cw + pad) * h;
            return 3;
        case 5:
            rc[0] = (struct rawcomp){0, 1, (size_t)w + pad};
            rc[1] = (struct rawcomp){((size_t)w + pad) * h, 1, (size_t)cw + pad};
            rc[2] = (struct rawcomp){((size_t)w + pad) * h + ((size_t)cw + pad) * ch, 1, (size_t)cw + pad};
            hs[0] = vs[0] = 2;
            *size = ((size_t)w + pad) * h + 2 * ((size_t)cw + pad) * ch;
            return 3;
        default: return 0;
    }
}

size_t orc_raw_size(int fmt, int w, int h, int pad)
{
    struct rawcomp rc[3];
    int hs[4], vs[4];
    size_t size = 0;
    return raw_layout(fmt, w, h, pad, rc, hs, vs, &size) ? size : 0;
}

/* Encode an image whose samples already are the JPEG's components (Y or YCbCr in the internal colour space):
 * sampling follows the pixel format. */
size_t orc_encode_ycc(const uint8_t* raw, int w, int h, int pad, int fmt, int quality, int rst, int interleaved,
                      int threads, uint8_t* out, int16_t* coef_out)
{
    if ( w <= 0 || h <= 0 || w > 65535 || h > 65535 || rst < 0 || rst > 65535 ) return 0;
    struct rawcomp rc[3];
    int hs[4], vs[4];
    size_t size;
    const int comps = raw_layout(fmt, w, h, pad, rc, hs, vs, &size);
    if ( !comps ) return 0;
    enc_tables_init();
#ifdef _OPENMP
    int saved_threads = omp_get_max_threads();
    omp_set_num_threads(threads > 1 ? threads : 1);
#else
    (void)threads;
#endif
    if ( comps == 1 ) interleaved = 0;
    struct ogeo g[4];
    int max_hs, max_vs;
    size_t total = ogeo_init(g, comps, hs, vs, w, h, interleaved, &max_hs, &max_vs);
    uint8_t* planes = (uint8_t*)scratch(0, total);
    int16_t* coef = coef_out ? coef_out : (int16_t*)scratch(1, total * sizeof(int16_t));
    memset(planes, 0, total);
    for ( int c = 0; c < comps; c++ )
        for ( int y = 0; y < g[c].h; y++ )
            for ( int x = 0; x < g[c].w; x++ )
                planes[g[c].off + (size_t)y * g[c].dw + x] = raw[rc[c].off + (size_t)y * rc[c].pitch + (size_t)x * rc[c].xs];
    size_t n = encode_from_planes(planes, g, comps, hs, vs, w, h, quality, rst, interleaved, out, coef);
#ifdef _OPENMP
    omp_set_num_threads(saved_threads);
#endif
    return n;
}

/* ------------------------------------------------------------------------------------------- */
/* generic pre-/post-processing: any of the pixel formats above in any of the colour spaces RGB, YCbCr BT.601,
 * YCbCr BT.601 full range ("JPEG"), YCbCr BT.709, with any JPEG sampling -- the reference's per-pixel kernels
 * [ref: src/gpujpeg_preprocessor.cu:163-201, src/gpujpeg_postprocessor.cu:183-216, src/gpujpeg_colorspace.h:52-427] */

enum { CS_NONE = 0, CS_RGB = 1, CS_601 = 2, CS_601_256 = 3, CS_709 = 4 };

static const int k_to_rgb[5][9] = {{0}, {0}, {298, 0, 409, 298, -100, -208, 298, 516, 0},
                                   {256, 0, 359, 256, -88, -183, 256, 454, 0}, {298, 0, 459, 298, -55, -136, 298, 541, 0}};
static const int k_from_rgb[5][9] = {{0}, {0}, {66, 129, 25, -38, -74, 112, 112, -94, -18},
                                     {77, 150, 29, -43, -85, 128, 128, -107, -21}, {47, 157, 16, -26, -87, 112, 112, -102, -10}};
static const int k_base[5][3] = {{0, 0, 0}, {0, 0, 0}, {16, 128, 128}, {0, 128, 128}, {16, 128, 128}};

/* [ref: src/gpujpeg_colorspace.h:86-101] */
static void cs_to_rgb(int cs, int c[3])
{
    if ( cs == CS_RGB ) return;
    const int* m = k_to_rgb[cs];
    int r[3];
    for ( int i = 0; i < 3; i++ )
        r[i] = (c[i] - k_base[cs][i]) * 256 / 255;   /* C division: truncates toward zero */
    for ( int i = 0; i < 3; i++ )
        c[i] = clamp8((m[3 * i] * r[0] + m[3 * i + 1] * r[1] + m[3 * i + 2] * r[2] + 128) >> 8);
}
/* [ref: src/gpujpeg_colorspace.h:64-79] */
static void cs_from_rgb(int cs, int c[3])
{
    if ( cs == CS_RGB ) return;
    const int* m = k_from_rgb[cs];
    int r[3];
    for ( int i = 0; i < 3; i++ )
        r[i] = c[i] * 256 / 255;
    for ( int i = 0; i < 3; i++ )
        c[i] = clamp8(((m[3 * i] * r[0] + m[3 * i + 1] * r[1] + m[3 * i + 2] * r[2] + 128) >> 8) + k_base[cs][i]);
}
/* every YCbCr <-> YCbCr pair goes through RGB [ref: src/gpujpeg_colorspace.h:340-413] */
static void cs_transform(int from, int to, int c[3])
{
    if ( from == to || from == CS_NONE || to == CS_NONE ) return;
    cs_to_rgb(from, c);
    cs_from_rgb(to, c);
}

/* the sample triple the reference loads for pixel (x, y) [ref: src/gpujpeg_preprocessor.cu:88-160] */
/* enc_opt_flipped / enc_opt_channel_remap, dec_opt_flipped / dec_opt_channel_remap of the reference, for the generic
 * paths below (orc_encode_any*, orc_decode_any):
 *   channel remap  on the RAW image, pixel by pixel: out channel i = in channel (map >> 4i) & 15, 4 = 0xFF, 5 = 0x00
 *                  (__byte_perm of the packed pixel against 0x000000FF) [ref: src/gpujpeg_preprocessor.cu:488-514,
 *                  src/gpujpeg_encoder.c:662-698]; the encoder remaps before its colour transform, the decoder after
 *   flip           on the COMPONENT PLANES, each over its own padded height: row y <-> data_height - 1 - y
 *                  [ref: src/gpujpeg_preprocessor.cu:456-485]; the encoder flips after the preprocessor, the decoder before
 *                  the postprocessor [ref: src/gpujpeg_postprocessor.cu:447] */
static int g_flipped = 0;
static unsigned g_remap = 0;   /* (channel count << 24) | nibbles, 0 = none */
static int g_four_components = 0;   /* param.comp_count = 4: an RGBA image keeps its alpha as a fourth component */
void orc_set_four_components(int on) { g_four_components = on; }
void orc_set_flip_remap(int flipped, unsigned remap)
{
    g_flipped = flipped;
    g_remap = remap;
}
static void remap_channels(int c[4], int count)
{
    if ( !g_remap || (int)(g_remap >> 24) != count ) return;
    int in[4] = {c[0], c[1], c[2], c[3]};
    for ( int i = 0; i < count; i++ ) {
        const unsigned sel = (g_remap >> (4 * i)) & 15u;
        c[i] = sel < 4 ? in[sel] : sel == 4 ? 0xFF : 0;
    }
}
static void flip_plane(uint8_t* plane, int dw, int dh)
{
    for ( int y = 0; y < dh / 2; y++ )
        for ( int x = 0; x < dw; x++ ) {
            const uint8_t t = plane[(size_t)y * dw + x];
            plane[(size_t)y * dw + x] = plane[(size_t)(dh - 1 - y) * dw + x];
            plane[(size_t)(dh - 1 - y) * dw + x] = t;
        }
}

static void raw_load_pixel(const uint8_t* raw, int fmt, const struct rawcomp rc[3], const int fhs[4], const int fvs[4], int x,
                           int y, int c[3])
{
    if ( fmt == 0 ) {
        c[0] = raw[rc[0].off + (size_t)y * rc[0].pitch + x];
        c[1] = c[2] = 128;
        return;
    }
    for ( int k = 0; k < 3; k++ ) {
        int dh = fhs[0] / fhs[k], dv = fvs[0] / fvs[k];   /* luminance carries the format's maximum */
        c[k] = raw[rc[k].off + (size_t)(y / dv) * rc[k].pitch + (size_t)(x / dh) * rc[k].xs];
    }
}

size_t orc_encode_any2(const uint8_t* raw, int w, int h, int fmt, int cs, int internal, int quality, int rst, int interleaved,
                       int lhs, int lvs, int threads, uint8_t* out);
size_t orc_encode_any(const uint8_t* raw, int w, int h, int fmt, int cs, int quality, int rst, int interleaved, int lhs,
                      int lvs, int threads, uint8_t* out)
{
    return orc_encode_any2(raw, w, h, fmt, cs, CS_601_256, quality, rst, interleaved, lhs, lvs, threads, out);
}

/* internal = colour space of the JPEG's components: CS_601_256 (JFIF) or CS_RGB (Adobe APP14) */
size_t orc_encode_any2(const uint8_t* raw, int w, int h, int fmt, int cs, int internal, int quality, int rst, int interleaved,
                       int lhs, int lvs, int threads, uint8_t* out)
{
    if ( internal != CS_601_256 && internal != CS_RGB && internal != CS_601 && internal != CS_709 ) return 0;
    if ( w <= 0 || h <= 0 || w > 65535 || h > 65535 || rst < 0 || rst > 65535 ) return 0;
    struct rawcomp rc[3];
    int fhs[4], fvs[4];
    size_t size;
    const int fcomps = raw_layout(fmt, w, h, 0, rc, fhs, fvs, &size);
    if ( !fcomps || lhs < 1 || lhs > 2 || lvs < 1 || lvs > 2 ) return 0;
    enc_tables_init();
#ifdef _OPENMP
    int saved_threads = omp_get_max_threads();
    omp_set_num_threads(threads > 1 ? threads : 1);
#else
    (void)threads;
#endif
    /* comp_count = 4 with the four-sample pixel format: the alpha samples become a fourth component with the sampling of the
     * first, untouched by the colour transform [ref: src/gpujpeg_preprocessor.cu:131-138, 185-194] */
    const int comps = (g_four_components && fmt == 6) ? 4 : 3;
    const int hs[4] = {lhs, 1, 1, comps == 4 ? lhs : 0}, vs[4] = {lvs, 1, 1, comps == 4 ? lvs : 0};
    struct ogeo g[4];
    int max_hs, max_vs;
    size_t total = ogeo_init(g, comps, hs, vs, w, h, interleaved, &max_hs, &max_vs);
    uint8_t* planes = (uint8_t*)scratch(0, total);
    int16_t* coef = (int16_t*)scratch(1, total * sizeof(int16_t));
    memset(planes, 0, total);
    for ( int y = 0; y < h; y++ )
        for ( int x = 0; x < w; x++ ) {
            int c[4] = {0, 0, 0, 0};
            raw_load_pixel(raw, fmt, rc, fhs, fvs, x, y, c);
            if ( fmt == 6 ) c[3] = raw[rc[0].off + (size_t)y * rc[0].pitch + (size_t)x * 4 + 3];
            remap_channels(c, fmt == 6 ? 4 : fmt == 0 ? 1 : 3);
            cs_transform(cs, internal, c);
            for ( int k = 0; k < comps; k++ ) {
                int dh = max_hs / g[k].hs, dv = max_vs / g[k].vs;
                if ( x % dh || y % dv ) continue;
                planes[g[k].off + (size_t)(y / dv) * g[k].dw + x / dh] = (uint8_t)c[k];
            }
        }
    if ( g_flipped )
        for ( int k = 0; k < comps; k++ )
            flip_plane(planes + g[k].off, g[k].dw, g[k].dh);
    g_rgb_internal = internal == CS_RGB;
    g_spiff_cs = internal == CS_601 ? 4 : internal == CS_709 ? 1 : 0;
    size_t n = encode_from_planes(planes, g, comps, hs, vs, w, h, quality, rst, interleaved, out, coef);
    g_rgb_internal = g_spiff_cs = 0;
#ifdef _OPENMP
    omp_set_num_threads(saved_threads);
#endif
    return n;
}

/* ------------------------------------------------------------------------------------------- */
/* Huffman decoding of one restart segment (JPEG Annex F.2.2; behaviour on valid streams equals
 * [ref: src/gpujpeg_huffman_cpu_decoder.c:75-303])                                              */

struct dec_tab {
    int maxcode[18]; /* largest code of length l, -1 if none */
    int valptr[17];
    int mincode[17];
    const uint8_t* vals;
};

static void dec_table_build(struct dec_tab* t, const uint8_t* bits17, const uint8_t* vals)
{
    int p = 0;
    int code = 0;
    for ( int l = 1; l <= 16; l++ ) {
        if ( bits17[l] ) {
            t->valptr[l] = p;
            t->mincode[l] = code;
            p += bits17[l];
            code += bits17[l];
            t->maxcode[l] = code - 1;
        }
        else {
            t->maxcode[l] = -1;
        }
        code <<= 1;
    }
    t->maxcode[17] = 0x7FFFFFFF;
    t->vals = vals;
}

struct bitr {
    const uint8_t* p;
    const uint8_t* end;
    uint32_t acc;
    int n;
};

static inline int get_bit(struct bitr* r)
{
    if ( r->n == 0 ) {
        int b = 0; /* past the end: feed zeros (only reached on corrupt streams) */
        if ( r->p < r->end ) {
            b = *r->p++;
            if ( b == 0xFF && r->p < r->end && *r->p == 0 ) r->p++; /* drop the stuffed zero */
        }
        r->acc = (uint32_t)b;
        r->n = 8;
    }
    r->n--;
    return (int)((r->acc >> r->n) & 1u);
}
static inline int get_bits(struct bitr* r, int n)
{
    int v = 0;
    while ( n-- )
        v = (v << 1) | get_bit(r);
    return v;
}
static inline int decode_symbol(struct bitr* r, const struct dec_tab* t)
{
    int code = get_bit(r);
    int l = 1;
    while ( code > t->maxcode[l] ) {
        code = (code << 1) | get_bit(r);
        l++;
    }
    if ( l > 16 ) return 0; /* [ref: src/gpujpeg_huffman_cpu_decoder.c:155-159] garbage -> 0 */
    return t->vals[t->valptr[l] + code - t->mincode[l]];
}
/* [ref: src/gpujpeg_huffman_cpu_decoder.c:169-204] */
static inline int extend(int v, int n) { return v < (1 << (n - 1)) ? v - (1 << n) + 1 : v; }

int orc_huff_decode_segment(const uint8_t* data, size_t size, int nblocks, const uint8_t* dc_bits17,
                            const uint8_t* dc_vals, const uint8_t* ac_bits17, const uint8_t* ac_vals,
                            int16_t* coef)
{
    struct dec_tab dc, ac;
    dec_table_build(&dc, dc_bits17, dc_vals);
    dec_table_build(&ac, ac_bits17, ac_vals);
    struct bitr r = {data, data + size, 0, 0};
    int pred = 0;
    for ( int b = 0; b < nblocks; b++ ) {
        int16_t* blk = coef + (size_t)b * 64;
        memset(blk, 0, 64 * sizeof(int16_t));
        int s = decode_symbol(&r, &dc);
        int diff = s ? extend(get_bits(&r, s), s) : 0;
        pred += diff;
        blk[0] = (int16_t)pred;
        for ( int k = 1; k < 64; k++ ) {
            int rs = decode_symbol(&r, &ac);
            int run = rs >> 4, sz = rs & 15;
            if ( sz ) {
                k += run;
                int v = extend(get_bits(&r, sz), sz);
                if ( k < 64 ) blk[orc_zigzag_to_natural[k]] = (int16_t)v;
            }
            else {
                if ( run != 15 ) break;
                k += 15;
            }
        }
    }
    return 0;
}

/* ------------------------------------------------------------------------------------------- */
/* stream parsing + whole-frame decode                                                           */

struct parsed {
    int w, h, comps, rst;
    uint8_t qt[4][64];   /* zig-zag */
    int have_qt[4];
    uint8_t hbits[2][4][17];
    uint8_t hvals[2][4][256];
    int comp_id[4], comp_tq[4], comp_hv[4];
    int adobe_transform;   /* -1: no APP14 Adobe segment */
    int spiff_cs;          /* colour space code of a SPIFF header, 0 if none */
    int nscan;
    struct {
        int ncomp, comp[4], td[4], ta[4];
        size_t begin, end; /* entropy-coded data [begin,end) */
    } scan[4];
    size_t header_size;
};

static inline int rd16(const uint8_t* p) { return (p[0] << 8) | p[1]; }

/* marker walk  [ref: src/gpujpeg_reader.c:1619-1736 (loop), :681-730 DQT, :806-886 SOF0,
 * :920-986 DHT, :996-1027 DRI, :1256-1382 SOS] */
static int parse_stream(const uint8_t* j, size_t size, struct parsed* P)
{
    memset(P, 0, sizeof *P);
    P->adobe_transform = -1;
    size_t i = 0;
    if ( size < 4 || j[0] != 0xFF || j[1] != 0xD8 ) return -1;
    i = 2;
    while ( i + 2 <= size ) {
        if ( j[i] != 0xFF ) return -1;
        int m = j[i + 1];
        if ( m == 0xD9 ) return 0;
        if ( m == 0xFF ) { i++; continue; }
        if ( i + 4 > size ) return -1;
        int len = rd16(j + i + 2);
        const uint8_t* d = j + i + 4;
        if ( i + 2 + len > size ) return -1;
        if ( m == 0xDB ) {
            int off = 0;
            while ( off + 65 <= len - 2 ) {
                int pq = d[off] >> 4, tq = d[off] & 15;
                if ( pq != 0 || tq > 3 ) return -1;
                memcpy(P->qt[tq], d + off + 1, 64);
                P->have_qt[tq] = 1;
                off += 65;
            }
        }
        else if ( m == 0xC0 ) {
            if ( d[0] != 8 ) return -1;
            P->h = rd16(d + 1);
            P->w = rd16(d + 3);
            P->comps = d[5];
            if ( P->comps < 1 || P->comps > 4 ) return -1;
            for ( int c = 0; c < P->comps; c++ ) {
                P->comp_id[c] = d[6 + 3 * c];
                P->comp_hv[c] = d[7 + 3 * c];
                P->comp_tq[c] = d[8 + 3 * c];
            }
        }
        else if ( m == 0xC4 ) {
            int off = 0;
            while ( off + 17 <= len - 2 ) {
                int tc = d[off] >> 4, th = d[off] & 15;
                if ( tc > 1 || th > 3 ) return -1;
                int n = 0;
                P->hbits[tc][th][0] = 0;
                for ( int k = 1; k <= 16; k++ ) {
                    P->hbits[tc][th][k] = d[off + k];
                    n += d[off + k];
                }
                if ( n > 256 || off + 17 + n > len - 2 ) return -1;
                memcpy(P->hvals[tc][th], d + off + 17, n);
                off += 17 + n;
            }
        }
        else if ( m == 0xDD ) {
            P->rst = rd16(d);
        }
        else if ( m == 0xE8 && len == 32 && memcmp(d, "SPIFF", 6) == 0 && !P->spiff_cs ) {
            P->spiff_cs = d[18];   /* [ref: src/gpujpeg_reader.c:393-443] */
        }
        else if ( m == 0xEE && len >= 14 && memcmp(d, "Adobe", 5) == 0 ) {
            P->adobe_transform = d[11];   /* [ref: src/gpujpeg_reader.c:560-640] */
        }
        else if ( m == 0xDA ) {
            if ( P->nscan == 0 ) P->header_size = i;
            if ( P->nscan >= 4 ) return -1;
            int s = P->nscan++;
            P->scan[s].ncomp = d[0];
            for ( int c = 0; c < d[0]; c++ ) {
                int id = d[1 + 2 * c];
                int idx = -1;
                for ( int q = 0; q < P->comps; q++ )
                    if ( P->comp_id[q] == id ) idx = q;
                if ( idx < 0 ) return -1;
                P->scan[s].comp[c] = idx;
                P->scan[s].td[c] = d[2 + 2 * c] >> 4;
                P->scan[s].ta[c] = d[2 + 2 * c] & 15;
            }
            size_t b = i + 2 + len;
            size_t e = b;
            /* entropy-coded data runs until a marker that is neither RSTn nor a stuffed zero
             * [ref: src/gpujpeg_reader.c:1038-1155] */
            while ( e + 1 < size ) {
                const uint8_t* f = (const uint8_t*)memchr(j + e, 0xFF, size - 1 - e);
                if ( !f ) { e = size; break; }
                e = (size_t)(f - j);
                int mm = j[e + 1];
                if ( mm == 0 || (mm >= 0xD0 && mm <= 0xD7) ) { e += 2; continue; }
                if ( mm == 0xFF ) { e++; continue; }
                break;
            }
            P->scan[s].begin = b;
            P->scan[s].end = e;
            i = e;
            continue;
        }
        i += 2 + len;
    }
    return -1;
}

/* next marker at or behind i inside [i, e): position of its 0xFF, or e.  Stuffed zeros and fill bytes are skipped. */
static size_t next_marker(const uint8_t* j, size_t i, size_t e)
{
    while ( i + 1 < e ) {
        const uint8_t* f = (const uint8_t*)memchr(j + i, 0xFF, e - 1 - i);
        if ( !f ) break;
        i = (size_t)(f - j);
        if ( j[i + 1] == 0xFF ) { i++; continue; }   /* fill byte in front of a marker (T.81 B.1.1.2) */
        if ( j[i + 1] == 0x00 ) { i += 2; continue; }
        return i;
    }
    return e;
}

/* split one scan at its RST markers, resynchronising on a broken restart sequence as the reference's reader does
 * [ref: src/gpujpeg_reader.c:1038-1155]: restart markers must count RST0..RST7 cyclically; at a marker with another number
 * the current segment ends there, everything up to the next marker that carries the EXPECTED number is skipped, and the
 * data behind that marker becomes the next segment (numbered consecutively -- later segments move up); when no such
 * marker follows, the rest of the scan is dropped.  Returns the number of segments found (<= max_seg). */
static int split_scan(const uint8_t* j, size_t b, size_t e, size_t* seg_off, size_t* seg_len, int max_seg)
{
    int n = 0, prev = 7;   /* "RST0 - 1" */
    size_t start = b, i = b;
    for ( ;; ) {
        i = next_marker(j, i, e);
        if ( i >= e ) break;
        const int m = j[i + 1];
        if ( m < 0xD0 || m > 0xD7 ) break;   /* cannot happen: the scan ends at the first other marker */
        const int expected = (prev + 1) & 7;
        if ( n >= max_seg ) return -1;
        seg_off[n] = start;
        seg_len[n] = i - start;
        n++;
        if ( (m & 7) != expected ) {
            size_t k = next_marker(j, i + 2, e);
            while ( k < e && j[k + 1] != 0xD0 + expected )
                k = next_marker(j, k + 2, e);
            if ( k >= e ) return n;   /* nothing to resynchronise on: the rest of the scan is lost */
            i = k;
        }
        prev = expected;
        start = i + 2;
        i += 2;
    }
    if ( n >= max_seg ) return -1;
    seg_off[n] = start;
    seg_len[n] = e - start;
    return n + 1;
}

int orc_probe(const uint8_t* jpeg, size_t size, struct orc_stream_info* info)
{
    struct parsed P;
    if ( parse_stream(jpeg, size, &P) != 0 ) return -1;
    memset(info, 0, sizeof *info);
    info->width = P.w;
    info->height = P.h;
    info->comp_count = P.comps;
    info->restart_interval = P.rst;
    info->scan_count = P.nscan;
    info->interleaved = P.nscan == 1 && P.comps > 1;
    info->header_size = P.header_size;
    for ( int s = 0; s < P.nscan; s++ ) {
        info->scan_bytes[s] = P.scan[s].end - P.scan[s].begin;
        int n = 1;
        for ( size_t i = P.scan[s].begin; i + 1 < P.scan[s].end; i++ )
            if ( jpeg[i] == 0xFF && jpeg[i + 1] >= 0xD0 && jpeg[i + 1] <= 0xD7 ) n++;
        info->segment_count += n;
    }
    return 0;
}

/* decode one block of an interleaved MCU (JPEG F.2.2) */
static inline void decode_block(struct bitr* r, const struct dec_tab* dct, const struct dec_tab* act, int* pred, int16_t* blk)
{
    memset(blk, 0, 128);
    int sz = decode_symbol(r, dct);
    int diff = sz ? extend(get_bits(r, sz), sz) : 0;
    *pred += diff;
    blk[0] = (int16_t)*pred;
    for ( int k = 1; k < 64; k++ ) {
        int rs = decode_symbol(r, act);
        int run = rs >> 4, z = rs & 15;
        if ( z ) {
            k += run;
            int v = extend(get_bits(r, z), z);
            if ( k < 64 ) blk[orc_zigzag_to_natural[k]] = (int16_t)v;
        }
        else {
            if ( run != 15 ) break;
            k += 15;
        }
    }
}

/* Entropy-decode + IDCT a 1- or 3-component stream of any sampling into its component planes (scratch slot 7).
 * [ref: src/gpujpeg_huffman_cpu_decoder.c:305-420] */
static uint8_t* decode_to_planes(const struct parsed* P, const uint8_t* jpeg, int idct_flavour, struct ogeo g[4], int* max_hs,
                                 int* max_vs, int16_t* coef_out, int want_planes)
{
    const int comps = P->comps;
    int hs[4] = {0, 0, 0, 0}, vs[4] = {0, 0, 0, 0};
    for ( int c = 0; c < comps; c++ ) {
        hs[c] = comps == 1 ? 1 : P->comp_hv[c] >> 4;
        vs[c] = comps == 1 ? 1 : P->comp_hv[c] & 15;
        if ( hs[c] < 1 || vs[c] < 1 || hs[c] > 4 || vs[c] > 4 ) return NULL;
    }
    const int interleaved = comps > 1 && P->nscan == 1;
    if ( !interleaved && P->nscan != comps ) return NULL;
    size_t total = ogeo_init(g, comps, hs, vs, P->w, P->h, interleaved, max_hs, max_vs);
    for ( int c = 0; c < comps; c++ )
        if ( *max_hs % hs[c] || *max_vs % vs[c] ) return NULL;
    int16_t* coef = coef_out ? coef_out : (int16_t*)scratch(4, total * sizeof(int16_t));
    int mcu_x = g[0].bcx / g[0].hs;
    int rc = 0;
    for ( int s = 0; s < P->nscan && rc == 0; s++ ) {
        int mcus = interleaved ? mcu_x * (g[0].bcy / g[0].vs) : g[P->scan[s].comp[0]].nblk;
        int seg_mcu = P->rst > 0 ? P->rst : mcus;
        int nseg = (mcus + seg_mcu - 1) / seg_mcu;
        size_t* so = (size_t*)scratch(5, sizeof(size_t) * (nseg + 1));
        size_t* sl = (size_t*)scratch(6, sizeof(size_t) * (nseg + 1));
        int n = split_scan(jpeg, P->scan[s].begin, P->scan[s].end, so, sl, nseg + 1);
        if ( n < 1 || n > nseg ) { rc = -1; break; }
        /* segments lost to a broken restart sequence: their blocks stay zero, as in the reference (coefficient buffer
         * cleared before decoding, src/gpujpeg_decoder.c:301) */
        if ( n < nseg ) {
            if ( !interleaved ) memset(coef + g[P->scan[s].comp[0]].off, 0, (size_t)g[P->scan[s].comp[0]].nblk * 128);
            else memset(coef, 0, total * sizeof(int16_t));
        }
        const int nseg_found = n;
        if ( (interleaved && P->scan[s].ncomp != comps) || (!interleaved && P->scan[s].ncomp != 1) ) { rc = -1; break; }
        struct dec_tab dct[4], act[4];
        for ( int c = 0; c < P->scan[s].ncomp; c++ ) {
            dec_table_build(&dct[c], P->hbits[0][P->scan[s].td[c]], P->hvals[0][P->scan[s].td[c]]);
            dec_table_build(&act[c], P->hbits[1][P->scan[s].ta[c]], P->hvals[1][P->scan[s].ta[c]]);
        }
#pragma omp parallel for schedule(dynamic, 16)
        for ( int q = 0; q < nseg_found; q++ ) {
            int first = q * seg_mcu;
            int cnt = mcus - first < seg_mcu ? mcus - first : seg_mcu;
            struct bitr r = {jpeg + so[q], jpeg + so[q] + sl[q], 0, 0};
            int pred[4] = {0, 0, 0, 0};
            if ( !interleaved ) {
                const struct ogeo* k = &g[P->scan[s].comp[0]];
                for ( int m = first; m < first + cnt; m++ )
                    decode_block(&r, &dct[0], &act[0], &pred[0], coef + k->off + (size_t)m * 64);
            }
            else {
                for ( int m = first; m < first + cnt; m++ )
                    for ( int c = 0; c < comps; c++ ) {
                        const struct ogeo* k = &g[P->scan[s].comp[c]];
                        for ( int y = 0; y < k->vs; y++ )
                            for ( int x = 0; x < k->hs; x++ )
                                decode_block(&r, &dct[c], &act[c], &pred[c], mcu_block(coef, k, mcu_x, m, x, y));
                    }
            }
        }
    }
    if ( rc ) return NULL;
    uint8_t* planes = (uint8_t*)scratch(7, total);
    if ( want_planes ) {
        for ( int c = 0; c < comps; c++ ) {
            uint16_t inv[64];
            for ( int i = 0; i < 64; i++ )
                inv[orc_zigzag_to_natural[i]] = P->qt[P->comp_tq[c]][i];
            orc_idct_plane(coef + g[c].off, g[c].dw, g[c].dh, inv, idct_flavour, planes + g[c].off);
        }
    }
    return planes;
}

/* 3-component streams with chroma subsampling -> RGB [ref: src/gpujpeg_postprocessor.cu:55-76] */
static int decode_subsampled(const struct parsed* P, const uint8_t* jpeg, int idct_flavour, int threads, uint8_t* rgb,
                             int16_t* coef_out)
{
#ifdef _OPENMP
    int saved_threads = omp_get_max_threads();
    omp_set_num_threads(threads > 1 ? threads : 1);
#else
    (void)threads;
#endif
    struct ogeo g[4];
    int max_hs, max_vs;
    uint8_t* planes = decode_to_planes(P, jpeg, idct_flavour, g, &max_hs, &max_vs, coef_out, rgb != NULL);
    if ( planes && rgb ) postprocess_rgb_ss(planes, g, max_hs, max_vs, rgb, P->w, P->h, 0);
#ifdef _OPENMP
    omp_set_num_threads(saved_threads);
#endif
    return planes ? 0 : -1;
}

/* Decode to a raw pixel format WITHOUT colour transform (the samples of the JPEG's components as they are):
 * the format's sampling must be the stream's.  Returns 0 on success. */
int orc_decode_ycc(const uint8_t* jpeg, size_t size, int idct_flavour, int threads, int fmt, int pad, uint8_t* raw)
{
    struct parsed P;
    if ( parse_stream(jpeg, size, &P) != 0 ) return -1;
    struct rawcomp rc[3];
    int hs[4], vs[4];
    size_t rsize;
    const int comps = raw_layout(fmt, P.w, P.h, pad, rc, hs, vs, &rsize);
    if ( !comps || comps != P.comps ) return -1;
    for ( int c = 0; c < comps && comps > 1; c++ )
        if ( (P.comp_hv[c] >> 4) != hs[c] || (P.comp_hv[c] & 15) != vs[c] ) return -1;
#ifdef _OPENMP
    int saved_threads = omp_get_max_threads();
    omp_set_num_threads(threads > 1 ? threads : 1);
#else
    (void)threads;
#endif
    struct ogeo g[4];
    int max_hs, max_vs;
    uint8_t* planes = decode_to_planes(&P, jpeg, idct_flavour, g, &max_hs, &max_vs, NULL, 1);
    if ( planes ) {
        for ( int c = 0; c < comps; c++ )
            for ( int y = 0; y < g[c].h; y++ )
                for ( int x = 0; x < g[c].w; x++ )
                    raw[rc[c].off + (size_t)y * rc[c].pitch + (size_t)x * rc[c].xs] = planes[g[c].off + (size_t)y * g[c].dw + x];
    }
#ifdef _OPENMP
    omp_set_num_threads(saved_threads);
#endif
    return planes ? 0 : -1;
}

/* Decode to any pixel format / colour space: per pixel, the component samples at (x / dh, y / dv), colour transform
 * from the JPEG's YCbCr, store by the format's rule [ref: src/gpujpeg_postprocessor.cu:183-216,
 * src/gpujpeg_preprocessor_common.cuh:125-203].  Even widths only for the formats that share chroma horizontally. */
int orc_decode_any(const uint8_t* jpeg, size_t size, int idct_flavour, int threads, int fmt, int cs, uint8_t* raw)
{
    struct parsed P;
    if ( parse_stream(jpeg, size, &P) != 0 ) return -1;
    struct rawcomp rc[3];
    int fhs[4], fvs[4];
    size_t rsize;
    const int fcomps = raw_layout(fmt, P.w, P.h, 0, rc, fhs, fvs, &rsize);
    if ( !fcomps || (fcomps == 1 && P.comps != 1) ) return -1;
#ifdef _OPENMP
    int saved_threads = omp_get_max_threads();
    omp_set_num_threads(threads > 1 ? threads : 1);
#else
    (void)threads;
#endif
    /* colour space of the components: Adobe transform 0 or ids 'R','G','B' mean RGB [ref: src/gpujpeg_reader.c:264-640] */
    int stream_cs = (P.comps >= 3 && (P.adobe_transform == 0 || (P.comp_id[0] == 'R' && P.comp_id[1] == 'G' && P.comp_id[2] == 'B')))
                        ? CS_RGB : CS_601_256;
    if ( P.comps >= 3 && P.spiff_cs )
        stream_cs = P.spiff_cs == 1 ? CS_709 : P.spiff_cs == 4 ? CS_601 : P.spiff_cs == 10 ? CS_RGB : CS_601_256;
    struct ogeo g[4];
    int max_hs, max_vs;
    uint8_t* planes = decode_to_planes(&P, jpeg, idct_flavour, g, &max_hs, &max_vs, NULL, 1);
    if ( planes && g_flipped )
        for ( int k = 0; k < P.comps; k++ )
            flip_plane(planes + g[k].off, g[k].dw, g[k].dh);
    if ( planes ) {
        for ( int y = 0; y < P.h; y++ )
            for ( int x = 0; x < P.w; x++ ) {
                int c[4] = {0, 128, 128, 0xFF};
                for ( int k = 0; k < P.comps; k++ ) {
                    int dh = max_hs / g[k].hs, dv = max_vs / g[k].vs;
                    c[k] = planes[g[k].off + (size_t)(y / dv) * g[k].dw + x / dh];
                }
                if ( P.comps >= 3 ) cs_transform(stream_cs, cs, c);   /* (a fourth component passes through) */
                remap_channels(c, fmt == 6 ? 4 : fcomps == 1 ? 1 : 3);
                if ( fcomps == 1 ) {
                    raw[rc[0].off + (size_t)y * rc[0].pitch + x] = (uint8_t)c[0];
                    continue;
                }
                raw[rc[0].off + (size_t)y * rc[0].pitch + (size_t)x * rc[0].xs] = (uint8_t)c[0];
                if ( fmt == 6 ) raw[rc[0].off + (size_t)y * rc[0].pitch + (size_t)x * 4 + 3] = (uint8_t)c[3];   /* alpha: the fourth component, 0xFF without one (unless remapped) */
                const int dh = fhs[0], dv = fvs[0];   /* chroma of the format: every dh-th pixel of every dv-th row */
                if ( fmt == 3 ) {
                    /* U from even pixels, V from odd pixels [ref: src/gpujpeg_preprocessor_common.cuh:179-189] */
                    if ( x % 2 == 0 ) raw[rc[1].off + (size_t)y * rc[1].pitch + (size_t)(x / 2) * rc[1].xs] = (uint8_t)c[1];
                    else raw[rc[2].off + (size_t)y * rc[2].pitch + (size_t)(x / 2) * rc[2].xs] = (uint8_t)c[2];
                }
                else if ( x % dh == 0 && y % dv == 0 ) {
                    raw[rc[1].off + (size_t)(y / dv) * rc[1].pitch + (size_t)(x / dh) * rc[1].xs] = (uint8_t)c[1];
                    raw[rc[2].off + (size_t)(y / dv) * rc[2].pitch + (size_t)(x / dh) * rc[2].xs] = (uint8_t)c[2];
                }
            }
    }
#ifdef _OPENMP
    omp_set_num_threads(saved_threads);
#endif
    return planes ? 0 : -1;
}

/* Quantised coefficients of a stream of any component count (1, 3 or 4) and sampling: component after component, blocks
 * in raster order, natural order inside a block -- the layout the reference keeps them in.  Returns the number of
 * coefficients, 0 on error; coef == NULL: the count only. */
size_t orc_decode_coefficients(const uint8_t* jpeg, size_t size, int16_t* coef)
{
    struct parsed P;
    if ( parse_stream(jpeg, size, &P) != 0 ) return 0;
    struct ogeo g[4];
    int max_hs, max_vs;
    if ( !coef ) {
        int hs[4] = {0, 0, 0, 0}, vs[4] = {0, 0, 0, 0};
        for ( int c = 0; c < P.comps; c++ ) {
            hs[c] = P.comps == 1 ? 1 : P.comp_hv[c] >> 4;
            vs[c] = P.comps == 1 ? 1 : P.comp_hv[c] & 15;
            if ( hs[c] < 1 || vs[c] < 1 ) return 0;
        }
        return ogeo_init(g, P.comps, hs, vs, P.w, P.h, P.comps > 1 && P.nscan == 1, &max_hs, &max_vs);
    }
    if ( !decode_to_planes(&P, jpeg, 0, g, &max_hs, &max_vs, coef, 0) ) return 0;
    return (size_t)g[P.comps - 1].off + (size_t)g[P.comps - 1].nblk * 64;
}

/* [ref: src/gpujpeg_decoder.c:234-469 with the CPU Huffman path :275-295 and gpujpeg_idct_cpu] */
int orc_decode_rgb(const uint8_t* jpeg, size_t size, int idct_flavour, int threads, uint8_t* rgb, int* w,
                   int* h, int* comps, int16_t* coef_out)
{
    struct parsed P;
    if ( parse_stream(jpeg, size, &P) != 0 ) return -1;
    if ( w ) *w = P.w;
    if ( h ) *h = P.h;
    if ( comps ) *comps = P.comps;
    if ( !rgb && !coef_out ) return 0;
    if ( P.comps != 3 && P.comps != 1 ) return -1;
    if ( P.comps == 3 && (P.comp_hv[0] != 0x11 || P.comp_hv[1] != 0x11 || P.comp_hv[2] != 0x11) )
        return decode_subsampled(&P, jpeg, idct_flavour, threads, rgb, coef_out);
#ifdef _OPENMP
    int saved_threads = omp_get_max_threads();
    omp_set_num_threads(threads > 1 ? threads : 1);
#else
    (void)threads;
#endif
    int dw = (P.w + 7) / 8 * 8, dh = (P.h + 7) / 8 * 8;
    size_t psz = (size_t)dw * dh;
    int nblk = (dw / 8) * (dh / 8);
    int16_t* coef = coef_out ? coef_out : (int16_t*)scratch(4, P.comps * psz * sizeof(int16_t));
    int seg_mcu = P.rst > 0 ? P.rst : nblk;
    int nseg = (nblk + seg_mcu - 1) / seg_mcu;
    size_t* so = (size_t*)scratch(5, sizeof(size_t) * (nseg + 1));
    size_t* sl = (size_t*)scratch(6, sizeof(size_t) * (nseg + 1));
    int rc = 0;
    for ( int s = 0; s < P.nscan && rc == 0; s++ ) {
        int n = split_scan(jpeg, P.scan[s].begin, P.scan[s].end, so, sl, nseg + 1);
        if ( n < 1 || n > nseg ) { rc = -1; break; }
        if ( n < nseg ) {   /* segments lost to a broken restart sequence stay zero (see decode_to_planes) */
            if ( P.scan[s].ncomp == 1 ) memset(coef + P.scan[s].comp[0] * psz, 0, psz * sizeof(int16_t));
            else memset(coef, 0, P.comps * psz * sizeof(int16_t));
        }
        const int nseg_found = n;
        if ( P.scan[s].ncomp == 1 ) {
            int c = P.scan[s].comp[0];
            const uint8_t* db = P.hbits[0][P.scan[s].td[0]];
            const uint8_t* dv = P.hvals[0][P.scan[s].td[0]];
            const uint8_t* ab = P.hbits[1][P.scan[s].ta[0]];
            const uint8_t* av = P.hvals[1][P.scan[s].ta[0]];
#pragma omp parallel for schedule(dynamic, 16)
            for ( int g = 0; g < nseg_found; g++ ) {
                int first = g * seg_mcu;
                int cnt = nblk - first < seg_mcu ? nblk - first : seg_mcu;
                orc_huff_decode_segment(jpeg + so[g], sl[g], cnt, db, dv, ab, av,
                                        coef + c * psz + (size_t)first * 64);
            }
        }
        else {
            /* interleaved 4:4:4: MCU = one block of each component  [ref: src/gpujpeg_huffman_cpu_decoder.c:330-366] */
            struct dec_tab dct[4], act[4];
            for ( int c = 0; c < P.scan[s].ncomp; c++ ) {
                dec_table_build(&dct[c], P.hbits[0][P.scan[s].td[c]], P.hvals[0][P.scan[s].td[c]]);
                dec_table_build(&act[c], P.hbits[1][P.scan[s].ta[c]], P.hvals[1][P.scan[s].ta[c]]);
            }
            for ( int g = 0; g < nseg_found; g++ ) {
                int first = g * seg_mcu;
                int cnt = nblk - first < seg_mcu ? nblk - first : seg_mcu;
                struct bitr r = {jpeg + so[g], jpeg + so[g] + sl[g], 0, 0};
                int pred[4] = {0, 0, 0, 0};
                for ( int m = 0; m < cnt; m++ ) {
                    for ( int c = 0; c < P.scan[s].ncomp; c++ ) {
                        int16_t* blk = coef + P.scan[s].comp[c] * psz + (size_t)(first + m) * 64;
                        memset(blk, 0, 128);
                        int sz = decode_symbol(&r, &dct[c]);
                        int diff = sz ? extend(get_bits(&r, sz), sz) : 0;
                        pred[c] += diff;
                        blk[0] = (int16_t)pred[c];
                        for ( int k = 1; k < 64; k++ ) {
                            int rs = decode_symbol(&r, &act[c]);
                            int run = rs >> 4, z = rs & 15;
                            if ( z ) {
                                k += run;
                                int v = extend(get_bits(&r, z), z);
                                if ( k < 64 ) blk[orc_zigzag_to_natural[k]] = (int16_t)v;
                            }
                            else {
                                if ( run != 15 ) break;
                                k += 15;
                            }
                        }
                    }
                }
            }
        }
    }
    if ( rc == 0 && rgb ) {
        uint8_t* planes = (uint8_t*)scratch(7, P.comps * psz);
        for ( int c = 0; c < P.comps; c++ ) {
            uint16_t inv[64];
            for ( int i = 0; i < 64; i++ )
                inv[orc_zigzag_to_natural[i]] = P.qt[P.comp_tq[c]][i];
            orc_idct_plane(coef + c * psz, dw, dh, inv, idct_flavour, planes + c * psz);
        }
        if ( P.comps == 3 ) {
            orc_postprocess_rgb444(planes, dw, dh, rgb, P.w, P.h, 0);
        }
        else {
            for ( int y = 0; y < P.h; y++ )
                memcpy(rgb + (size_t)y * P.w, planes + (size_t)y * dw, P.w);
        }
    }
#ifdef _OPENMP
    omp_set_num_threads(saved_threads);
#endif
    return rc;
}
