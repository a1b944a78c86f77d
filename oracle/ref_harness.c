/*
 * ref_harness.c -- drives the REFERENCE's own CPU code (compiled in place from /root/reference/src,
 * never copied) so the oracle restatement can be pinned against it.  TEST INFRASTRUCTURE ONLY.
 *
 * Built by oracle/Makefile into oracle/_ref/libgpujpeg_refcpu.so together with the reference's
 *   gpujpeg_huffman_cpu_encoder.c, gpujpeg_huffman_cpu_decoder.c, gpujpeg_table.c,
 *   gpujpeg_dct_cpu.c, gpujpeg_writer.c, gpujpeg_exif.c, utils/pam.c, utils/y4m.c
 * The handful of CUDA-runtime symbols those files reference are satisfied by the host
 * stubs below (the library must run on machines without a GPU; no libcudart is linked).
 *
 * What is exercised here is exactly SURVEY.md section 8a rows a3, a4, a8, a9, a12, a14.
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "src/gpujpeg_dct_cpu.h"
#include "src/gpujpeg_decoder_internal.h"
#include "src/gpujpeg_encoder_internal.h"
#include "src/gpujpeg_huffman_cpu_decoder.h"
#include "src/gpujpeg_huffman_cpu_encoder.h"
#include "src/gpujpeg_marker.h"
#include "src/gpujpeg_table.h"
#include "src/gpujpeg_writer.h"
#include "src/gpujpeg_exif.h"

/* ---- host stand-ins for the CUDA runtime calls made by the reference's CPU-side files ---- */
cudaError_t cudaMemcpy(void* dst, const void* src, size_t count, enum cudaMemcpyKind kind)
{
    (void)kind;
    memcpy(dst, src, count);
    return cudaSuccess;
}
cudaError_t cudaGetLastError(void) { return cudaSuccess; }
const char* cudaGetErrorString(cudaError_t e) { (void)e; return "host stub"; }
cudaError_t cudaMallocHost(void** p, size_t size) { *p = malloc(size); return *p ? cudaSuccess : cudaErrorMemoryAllocation; }
cudaError_t cudaFreeHost(void* p) { free(p); return cudaSuccess; }
/* symbols referenced by gpujpeg_writer.c / gpujpeg_dct_cpu.c that are never reached here */
int gpujpeg_coder_allocate_cpu_huffman_buf(struct gpujpeg_coder* c) { (void)c; abort(); }
/* defined in the reference's gpujpeg_common.c (CUDA-bound, not compiled here); used by the messages of gpujpeg_exif.c */
const char* gj_fg_red = "";
const char* gj_fg_yellow = "";
const char* gj_term_reset = "";
const char* gpujpeg_color_space_get_name(enum gpujpeg_color_space cs) { (void)cs; return "(colour space)"; }

/* ---- geometry restated without CUDA allocations [ref: src/gpujpeg_common.c:676-865] ---- */
struct geom {
    struct gpujpeg_component comp[4];
    struct gpujpeg_segment* seg;
    int seg_count;
};

static int geom_init(struct geom* g, int w, int h, int comps, int rst, int interleaved, int lhs, int lvs, int16_t* coef)
{
    memset(g, 0, sizeof *g);
    const int max_h = lhs, max_v = lvs;   /* luminance (and a fourth, alpha, component) carries the maximum, chrominance is 1x1 */
    size_t off = 0;
    for ( int c = 0; c < comps; c++ ) {
        struct gpujpeg_component* k = &g->comp[c];
        const int sh = (c == 0 || c == 3) ? lhs : 1, sv = (c == 0 || c == 3) ? lvs : 1;
        const int div_h = max_h / sh, div_v = max_v / sv;
        k->type = (c == 0 || c == 3) ? GPUJPEG_COMPONENT_LUMINANCE : GPUJPEG_COMPONENT_CHROMINANCE;   /* [ref: src/gpujpeg_common.c:692-694] */
        k->sampling_factor.horizontal = sh;
        k->sampling_factor.vertical = sv;
        k->width = ((w + div_h - 1) / div_h * div_h) * sh / max_h;
        k->height = ((h + div_v - 1) / div_v * div_v) * sv / max_v;
        k->mcu_size_x = interleaved ? 8 * sh : 8;
        k->mcu_size_y = interleaved ? 8 * sv : 8;
        k->mcu_size = k->mcu_size_x * k->mcu_size_y;
        k->data_width = (k->width + k->mcu_size_x - 1) / k->mcu_size_x * k->mcu_size_x;
        k->data_height = (k->height + k->mcu_size_y - 1) / k->mcu_size_y * k->mcu_size_y;
        k->data_size = (size_t)k->data_width * k->data_height;
        k->mcu_count_x = k->data_width / k->mcu_size_x;
        k->mcu_count_y = k->data_height / k->mcu_size_y;
        k->mcu_count = k->mcu_count_x * k->mcu_count_y;
        k->segment_mcu_count = rst ? rst : k->mcu_count;
        k->segment_count = (k->mcu_count + k->segment_mcu_count - 1) / k->segment_mcu_count;
        k->data_quantized = coef + off;
        off += k->data_size;
    }
    int nscan = interleaved ? 1 : comps;
    g->seg_count = 0;
    for ( int s = 0; s < nscan; s++ )
        g->seg_count += g->comp[s].segment_count;
    g->seg = (struct gpujpeg_segment*)calloc(g->seg_count, sizeof(struct gpujpeg_segment));
    if ( !g->seg ) return -1;
    int i = 0;
    for ( int s = 0; s < nscan; s++ ) {
        for ( int j = 0; j < g->comp[s].segment_count; j++, i++ ) {
            g->seg[i].scan_index = s;
            g->seg[i].scan_segment_index = j;
            int left = g->comp[s].mcu_count - j * g->comp[s].segment_mcu_count;
            g->seg[i].mcu_count = left < g->comp[s].segment_mcu_count ? left : g->comp[s].segment_mcu_count;
        }
    }
    return 0;
}

/* ---- exported harness entry points ---- */

/* raw zig-zag table, forward float table and inverse natural-order table exactly as the reference
 * computes them [ref: src/gpujpeg_table.c:102-166] */
int ref_quant_tables(int quality, uint8_t raw[2][64], float fwd[2][64], uint16_t inv[2][64])
{
    for ( int t = 0; t < 2; t++ ) {
        struct gpujpeg_table_quantization q;
        memset(&q, 0, sizeof q);
        uint16_t dev_inv[64];
        q.d_table_forward = fwd[t];
        q.d_table = dev_inv;
        if ( gpujpeg_table_quantization_encoder_init(&q, (enum gpujpeg_component_type)t, quality) != 0 ) return -1;
        memcpy(raw[t], q.table_raw, 64);
        if ( gpujpeg_table_quantization_decoder_init(&q, (enum gpujpeg_component_type)t, quality) != 0 ) return -1;
        memcpy(inv[t], q.table, 64 * sizeof(uint16_t));
    }
    return 0;
}

/* code/size per symbol [ref: src/gpujpeg_table.c:264-343] */
int ref_huff_encoder_table(int cls, int kind, uint32_t code[256], uint8_t size[256])
{
    struct gpujpeg_table_huffman_encoder t;
    memset(&t, 0, sizeof t);
    gpujpeg_table_huffman_encoder_init(&t, (enum gpujpeg_component_type)cls, (enum gpujpeg_huffman_type)kind);
    for ( int i = 0; i < 256; i++ ) {
        code[i] = t.code[i];
        size[i] = (uint8_t)t.size[i];
    }
    return 0;
}

/* Whole file from quantised coefficients (3 planes, block-major natural order): reference header
 * writer + reference CPU Huffman encoder + EOI
 * [ref: src/gpujpeg_encoder.c:504-534, 626; src/gpujpeg_writer.c:456-518;
 *       src/gpujpeg_huffman_cpu_encoder.c:296-376] */
static int g_internal_rgb = 0;   /* components of an RGB-internal JPEG are all of the luminance type */
static int g_internal_cs = 0;    /* other internal colour space (enum gpujpeg_color_space), 0 = YCbCr JPEG */
static int g_header_type = 0;    /* forced header flavour (enum gpujpeg_header_type), 0 = by colour space */
void ref_set_header_type(int t) { g_header_type = t; }
/* orientation metadata and custom Exif tags handed to the reference writer [ref: src/gpujpeg_encoder.c:700-779] */
static struct gpujpeg_image_metadata g_metadata;
static struct gpujpeg_exif_tags* g_exif_tags = NULL;
void ref_set_orientation(int set, int rotation, int flip)
{
    g_metadata.vals[GPUJPEG_METADATA_ORIENTATION].set = set;
    g_metadata.vals[GPUJPEG_METADATA_ORIENTATION].orient.rotation = rotation;
    g_metadata.vals[GPUJPEG_METADATA_ORIENTATION].orient.flip = flip;
}
int ref_add_exif_tag(const char* cfg) { return gpujpeg_exif_add_tag(&g_exif_tags, cfg) ? 0 : -1; }
void ref_clear_exif_tags(void) { g_exif_tags = NULL; /* (the reference's destroy routine walks its lists wrongly; leak) */ }
/* the reference's Exif APP1 parser on a buffer that starts at the segment's length field; returns what it found:
 * bit 0 = orientation set, bits 8.. = rotation, bit 4 = flip */
int ref_exif_parse(uint8_t* seg, size_t size)
{
    struct gpujpeg_image_metadata m;
    memset(&m, 0, sizeof m);
    uint8_t* p = seg;
    gpujpeg_exif_parse(&p, seg + size, 0, &m);
    return (m.vals[GPUJPEG_METADATA_ORIENTATION].set ? 1 : 0) | (m.vals[GPUJPEG_METADATA_ORIENTATION].orient.flip ? 16 : 0) |
           (m.vals[GPUJPEG_METADATA_ORIENTATION].orient.rotation << 8);
}

size_t ref_encode_from_coef_ss(int16_t* coef, int w, int h, int comps, int quality, int rst, int interleaved, int lhs,
                               int lvs, uint8_t* out, size_t out_cap)
{
    struct gpujpeg_encoder* enc = (struct gpujpeg_encoder*)calloc(1, sizeof *enc);
    struct gpujpeg_writer wr;
    struct geom g;
    memset(&wr, 0, sizeof wr);
    if ( !enc || geom_init(&g, w, h, comps, rst, interleaved, lhs, lvs, coef) != 0 ) return 0;
    memset(&enc->coder.param, 0, sizeof enc->coder.param); /* same values gpujpeg_set_default_parameters gives */
    enc->coder.param.quality = quality;
    enc->coder.param.restart_interval = rst;
    enc->coder.param.interleaved = interleaved;
    enc->coder.param.comp_count = comps;
    for ( int c = 0; c < comps; c++ ) {
        enc->coder.param.sampling_factor[c].horizontal = (c == 0 || c == 3) ? lhs : 1;
        enc->coder.param.sampling_factor[c].vertical = (c == 0 || c == 3) ? lvs : 1;
    }
    enc->coder.param.color_space_internal = g_internal_rgb ? GPUJPEG_RGB : g_internal_cs ? (enum gpujpeg_color_space)g_internal_cs
                                                                                         : GPUJPEG_YCBCR_BT601_256LVLS;
    if ( g_internal_rgb )
        for ( int c = 0; c < comps; c++ )
            g.comp[c].type = GPUJPEG_COMPONENT_LUMINANCE;   /* [ref: src/gpujpeg_common.c:689-692] */
    memset(&enc->coder.param_image, 0, sizeof enc->coder.param_image);
    enc->coder.param_image.color_space = GPUJPEG_RGB;
    enc->coder.param_image.pixel_format = GPUJPEG_444_U8_P012;
    enc->coder.param_image.width = w;
    enc->coder.param_image.height = h;
    enc->coder.component = g.comp;
    enc->coder.segment = g.seg;
    enc->coder.segment_count = g.seg_count;
    enc->header_type = (enum gpujpeg_header_type)g_header_type;
    float fwd[64];
    uint16_t dinv[64];
    for ( int t = 0; t < 2; t++ ) {
        enc->table_quantization[t].d_table_forward = fwd;
        enc->table_quantization[t].d_table = dinv;
        gpujpeg_table_quantization_encoder_init(&enc->table_quantization[t], (enum gpujpeg_component_type)t, quality);
        for ( int k = 0; k < 2; k++ )
            gpujpeg_table_huffman_encoder_init(&enc->table_huffman[t][k], (enum gpujpeg_component_type)t,
                                               (enum gpujpeg_huffman_type)k);
    }
    wr.buffer = out;
    wr.buffer_current = out;
    wr.buffer_allocated_size = out_cap;
    wr.metadata = g_metadata;
    wr.exif_tags = g_exif_tags;
    enc->writer = &wr;
    gpujpeg_writer_write_header(enc);
    size_t n = 0;
    if ( gpujpeg_huffman_cpu_encoder_encode(enc) == 0 ) {
        gpujpeg_writer_emit_marker(enc->writer, GPUJPEG_MARKER_EOI);
        n = (size_t)(wr.buffer_current - wr.buffer);
    }
    free(g.seg);
    free(enc);
    return n;
}

/* the same for an RGB-internal JPEG (Adobe APP14 header, luminance tables for every component) */
size_t ref_encode_from_coef_rgb(int16_t* coef, int w, int h, int quality, int rst, int interleaved, int lhs, int lvs,
                                uint8_t* out, size_t out_cap)
{
    g_internal_rgb = 1;
    size_t n = ref_encode_from_coef_ss(coef, w, h, 3, quality, rst, interleaved, lhs, lvs, out, out_cap);
    g_internal_rgb = 0;
    return n;
}

/* RGB-internal with any component count (4: 'R','G','B','A' under a SPIFF header) */
size_t ref_encode_from_coef_rgb_n(int16_t* coef, int w, int h, int comps, int quality, int rst, int interleaved, uint8_t* out,
                                  size_t out_cap)
{
    g_internal_rgb = 1;
    size_t n = ref_encode_from_coef_ss(coef, w, h, comps, quality, rst, interleaved, 1, 1, out, out_cap);
    g_internal_rgb = 0;
    return n;
}

/* the same for the limited-range internal colour spaces (SPIFF header): cs = GPUJPEG_YCBCR_BT601 or _BT709 */
size_t ref_encode_from_coef_cs(int16_t* coef, int w, int h, int quality, int rst, int interleaved, int lhs, int lvs, int cs,
                               uint8_t* out, size_t out_cap)
{
    g_internal_cs = cs;
    size_t n = ref_encode_from_coef_ss(coef, w, h, 3, quality, rst, interleaved, lhs, lvs, out, out_cap);
    g_internal_cs = 0;
    return n;
}

size_t ref_encode_from_coef(int16_t* coef, int w, int h, int comps, int quality, int rst, int interleaved,
                            uint8_t* out, size_t out_cap)
{
    return ref_encode_from_coef_ss(coef, w, h, comps, quality, rst, interleaved, 1, 1, out, out_cap);
}

/* Decode entropy-coded data with the reference CPU Huffman decoder.  The caller supplies the
 * segment table (scan index, index in scan, byte offset, byte size -- markers stripped, stuffing
 * kept) exactly as the reference reader would build it [ref: src/gpujpeg_reader.c:1038-1155], and
 * the DHT contents per table id.  [ref: src/gpujpeg_huffman_cpu_decoder.c:371-425] */
int ref_huff_decode_ss(const uint8_t* data, size_t data_size, int w, int h, int comps, int rst, int interleaved, int lhs,
                       int lvs, int nseg, const int* seg_scan, const int* seg_index, const size_t* seg_off,
                       const size_t* seg_size, const uint8_t dht_bits[2][2][17], const uint8_t dht_vals[2][2][256],
                       int16_t* coef)
{
    struct gpujpeg_decoder* dec = (struct gpujpeg_decoder*)calloc(1, sizeof *dec);
    struct geom g;
    if ( !dec || geom_init(&g, w, h, comps, rst, interleaved, lhs, lvs, coef) != 0 ) return -1;
    if ( nseg != g.seg_count ) return -2;
    for ( int i = 0; i < nseg; i++ ) {
        g.seg[i].scan_index = seg_scan[i];
        g.seg[i].scan_segment_index = seg_index[i];
        g.seg[i].data_compressed_index = seg_off[i];
        g.seg[i].data_compressed_size = seg_size[i];
    }
    (void)data_size;
    dec->coder.param.comp_count = comps;
    dec->coder.param.interleaved = interleaved;
    dec->coder.param.restart_interval = rst;
    dec->coder.component = g.comp;
    dec->coder.segment = g.seg;
    dec->coder.segment_count = nseg;
    dec->segment_count = nseg;
    dec->coder.data_compressed = (uint8_t*)data;
    for ( int id = 0; id < 2; id++ ) {
        for ( int k = 0; k < 2; k++ ) { /* k: 0 = DC, 1 = AC; indexed [Th][Tc] as the reader does */
            struct gpujpeg_table_huffman_decoder* t = &dec->table_huffman[id][k];
            memcpy(t->bits, dht_bits[k][id], 17);
            memcpy(t->huffval, dht_vals[k][id], 256);
            gpujpeg_table_huffman_decoder_compute(t);
        }
    }
    for ( int c = 0; c < comps; c++ ) {
        dec->comp_table_huffman_map[c][GPUJPEG_HUFFMAN_DC] = (c == 0 || c == 3) ? 0 : 1;   /* what the reader takes from the SOS selectors; a fourth component codes like luminance */
        dec->comp_table_huffman_map[c][GPUJPEG_HUFFMAN_AC] = (c == 0 || c == 3) ? 0 : 1;
    }
    int rc = gpujpeg_huffman_cpu_decoder_decode(dec);
    free(g.seg);
    free(dec);
    return rc;
}

int ref_huff_decode(const uint8_t* data, size_t data_size, int w, int h, int comps, int rst, int interleaved,
                    int nseg, const int* seg_scan, const int* seg_index, const size_t* seg_off,
                    const size_t* seg_size, const uint8_t dht_bits[2][2][17], const uint8_t dht_vals[2][2][256],
                    int16_t* coef)
{
    return ref_huff_decode_ss(data, data_size, w, h, comps, rst, interleaved, 1, 1, nseg, seg_scan, seg_index, seg_off,
                              seg_size, dht_bits, dht_vals, coef);
}

/* dequantise + integer IDCT of one block, in place, result before the +128 level shift
 * [ref: src/gpujpeg_dct_cpu.c:178-199] */
void ref_idct_block(int16_t blk[64], const uint16_t inv[64])
{
    static int init = 0;
    if ( !init ) {
        gpujpeg_idct_cpu_init();
        init = 1;
    }
    int16_t t[64];
    memcpy(t, inv, sizeof t);
    gpujpeg_idct_cpu_perform(blk, t);
}
