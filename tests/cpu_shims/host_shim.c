/* Exposes a few internal host functions of the product (tables, writer, reader) to the CPU tests.
 * Compiled by tests/_shims.py together with gj_tables.c and gj_codestream.c (no CUDA involved). */
#include <string.h>

#include "../../gpujpeg_b200/csrc/gj_internal.h"

static int g_header_type = 0;   /* GPUJPEG_HEADER_DEFAULT; shim_set_header_type forces a flavour (enc_hdr option) */
void shim_set_header_type(int t) { g_header_type = t; }
void shim_set_alpha_sampling(int on) { (void)on; }   /* (kept for the tests' symmetry: the fourth component always follows the first) */
static struct gj_header_extras g_extras;   /* orientation and user Exif tags of the headers composed below */
void shim_set_orientation(int set, int rotation, int flip)
{
    g_extras.metadata.vals[GPUJPEG_METADATA_ORIENTATION].set = set != 0;
    g_extras.metadata.vals[GPUJPEG_METADATA_ORIENTATION].orient.rotation = (unsigned)rotation & 3u;
    g_extras.metadata.vals[GPUJPEG_METADATA_ORIENTATION].orient.flip = flip != 0;
}
int shim_add_exif_tag(const char* cfg) { return gj_exif_add_tag((struct gj_exif_tags**)&g_extras.exif_tags, cfg); }
void shim_clear_exif_tags(void)
{
    gj_exif_tags_destroy((struct gj_exif_tags*)g_extras.exif_tags);
    g_extras.exif_tags = NULL;
}

int shim_header(int width, int height, int quality, int rst, int interleaved, unsigned char* out)
{
    struct gpujpeg_parameters p;
    struct gpujpeg_image_parameters pi;
    memset(&p, 0, sizeof p);
    memset(&pi, 0, sizeof pi);
    p.quality = quality;
    p.restart_interval = rst;
    p.interleaved = interleaved;
    p.comp_count = 3;
    for ( int c = 0; c < 4; c++ )
        p.sampling_factor[c].horizontal = p.sampling_factor[c].vertical = 1;
    p.color_space_internal = GPUJPEG_YCBCR_BT601_256LVLS;
    pi.width = width;
    pi.height = height;
    pi.color_space = GPUJPEG_RGB;
    pi.pixel_format = GPUJPEG_444_U8_P012;
    uint8_t raw[2][64];
    struct gj_huff_spec spec[2][2];
    for ( int t = 0; t < 2; t++ ) {
        gj_quant_raw(t, quality, raw[t]);
        for ( int k = 0; k < 2; k++ )
            gj_huff_spec_default(t, k, &spec[t][k]);
    }
    size_t n = gj_write_header(out, &p, &pi, raw, spec, (enum gpujpeg_header_type)g_header_type, &g_extras);
    n += gj_write_sos(out + n, &p, 0);
    return (int)n;
}

void shim_forward_table_zz(int cls, int quality, float* fwd_zz, unsigned char* raw_zz)
{
    gj_quant_raw(cls, quality, raw_zz);
    gj_quant_forward_zz(raw_zz, fwd_zz);
}

void shim_enc_lut(int cls, unsigned* ac256, unsigned* dc16)
{
    struct gj_huff_spec dc, ac;
    struct gj_enc_lut lut;
    gj_huff_spec_default(cls, 0, &dc);
    gj_huff_spec_default(cls, 1, &ac);
    gj_enc_lut_build(&dc, &ac, &lut);
    memcpy(ac256, lut.ac, sizeof lut.ac);
    memcpy(dc16, lut.dc, sizeof lut.dc);
}

/* decode one symbol from a 16-bit peek with the device LUT logic (mirrors decode_symbol in gj_huffman.cu) */
int shim_dec_lut_symbol(int cls, int kind, unsigned peek16, int* len)
{
    struct gj_huff_spec spec;
    struct gj_dec_lut t;
    gj_huff_spec_default(cls, kind, &spec);
    if ( gj_dec_lut_build(&spec, &t) ) return -1;
    const unsigned e = t.look[peek16 >> (16 - GJ_DEC_LOOK_BITS)];
    if ( e & 15u ) {
        *len = (int)(e & 15u);
        return (int)(e >> 4);
    }
    int l = GJ_DEC_LOOK_BITS + 1;
    while ( l <= 16 && peek16 >= t.maxcode[l] ) l++;
    if ( l > 16 ) {
        *len = 16;
        return -2;
    }
    *len = l;
    return t.vals[((int)(peek16 >> (16 - l)) + t.valoff[l]) & 255];
}

/* decode one symbol from a 16-bit peek with the self-synchronising decoder's table (mirrors lookup / long_code in
 * gj_huffdec.cu): returns the symbol's zig-zag advance, *total = code length + value size, *size = value size, *how =
 * 1 first-level entry, 2 second-level table, 3 canonical search needed (entry 0) */
int shim_dec_fast_symbol(int cls, int kind, unsigned peek16, int* total, int* size, int* how)
{
    struct gj_huff_spec spec;
    static struct gj_dec_fast f;
    gj_huff_spec_default(cls, kind, &spec);
    gj_dec_fast_build(&spec, kind, &f);
    unsigned e = f.e[peek16 >> (16 - GJ_DEC_FAST_BITS)];
    *how = 1;
    if ( (e & GJ_DEC_FAST_TOTAL_MASK) == 0 ) {
        *how = 2;
        if ( e ) {
            if ( (e & 127u) < 1 || (e & 127u) > GJ_DEC_FAST_SUBS ) return -2;
            e = f.sub[(e & 127u) - 1][peek16 & ((1u << (16 - GJ_DEC_FAST_BITS)) - 1u)];
        }
        if ( e == 0 ) {
            *how = 3;
            return -1;
        }
    }
    *total = (int)((e >> GJ_DEC_FAST_TOTAL_SHIFT) & 31u);
    *size = (int)(e >> GJ_DEC_FAST_SIZE_SHIFT);
    return (int)(e & 127u);
}

/* parse + split; returns number of segments, fills a few fields */
int shim_parse(const unsigned char* data, size_t size, int* info /*[8]*/, unsigned* seg_off, unsigned* seg_len, int max_seg)
{
    struct gj_stream s;
    if ( gj_reader_parse(data, size, &s, 0) ) return -1;
    info[0] = s.width;
    info[1] = s.height;
    info[2] = s.comp_count;
    info[3] = s.restart_interval;
    info[4] = s.scan_count;
    info[5] = s.interleaved;
    info[6] = (int)s.header_size;
    info[7] = (int)s.color_space;
    return gj_reader_split(data, &s, seg_off, seg_len, max_seg);
}

/* header flavour and orientation metadata the reader finds: out = {header type, orientation set, rotation, flip, colour space} */
int shim_parse_meta(const unsigned char* data, size_t size, int* out /*[5]*/)
{
    struct gj_stream s;
    if ( gj_reader_parse(data, size, &s, 0) ) return -1;
    out[0] = (int)s.header_type;
    out[1] = (int)s.metadata.vals[GPUJPEG_METADATA_ORIENTATION].set;
    out[2] = (int)s.metadata.vals[GPUJPEG_METADATA_ORIENTATION].orient.rotation;
    out[3] = (int)s.metadata.vals[GPUJPEG_METADATA_ORIENTATION].orient.flip;
    out[4] = (int)s.color_space;
    return 0;
}

int shim_geometry(int width, int height, int rst, int interleaved, long* out /*[8]*/)
{
    struct gpujpeg_parameters p;
    struct gpujpeg_image_parameters pi;
    memset(&p, 0, sizeof p);
    memset(&pi, 0, sizeof pi);
    p.restart_interval = rst;
    p.interleaved = interleaved;
    p.comp_count = 3;
    pi.width = width;
    pi.height = height;
    struct gj_geometry g;
    gj_geometry_init(&g, &p, &pi);
    out[0] = g.bcx; out[1] = g.bcy; out[2] = g.nblk; out[3] = g.seg_per_scan; out[4] = g.seg_count;
    out[5] = g.scan_count; out[6] = (long)g.slot_stride; out[7] = (long)g.coef_count;
    return 0;
}

/* geometry with chroma subsampling: per component {bcx, bcy, blk_off}, then per scan segment counts,
 * then {seg_count, seg_mcu, bpm, mcu_x, coef_count, slot_stride} */
int shim_geometry_ss(int width, int height, int rst, int interleaved, int lhs, int lvs, long* out /*[24]*/)
{
    struct gpujpeg_parameters p;
    struct gpujpeg_image_parameters pi;
    memset(&p, 0, sizeof p);
    memset(&pi, 0, sizeof pi);
    p.restart_interval = rst;
    p.interleaved = interleaved;
    p.comp_count = 3;
    for ( int c = 0; c < 3; c++ ) {
        p.sampling_factor[c].horizontal = (uint8_t)(c == 0 ? lhs : 1);
        p.sampling_factor[c].vertical = (uint8_t)(c == 0 ? lvs : 1);
    }
    pi.width = width;
    pi.height = height;
    struct gj_geometry g;
    if ( gj_geometry_init(&g, &p, &pi) ) return -1;
    for ( int c = 0; c < 3; c++ ) {
        out[3 * c] = g.comp[c].bcx;
        out[3 * c + 1] = g.comp[c].bcy;
        out[3 * c + 2] = g.comp[c].blk_off;
    }
    for ( int k = 0; k < 4; k++ )
        out[9 + k] = g.lay.scan_seg_begin[k + 1] - g.lay.scan_seg_begin[k];
    out[13] = g.seg_count; out[14] = g.seg_mcu; out[15] = g.lay.bpm; out[16] = g.lay.mcu_x;
    out[17] = (long)g.coef_count; out[18] = (long)g.slot_stride; out[19] = g.lay.simple;
    for ( int i = 0; i < 4; i++ )
        out[20 + i] = i < g.lay.bpm ? g.lay.idx_pred[i] : 0;
    return 0;
}

/* raw image layout of a pixel format: out = {comp_count, size, then per component off, pitch, xs, sampling h, v} */
int shim_raw_layout(int fmt, int width, int height, int pad, long* out /*[17]*/)
{
    struct gpujpeg_image_parameters pi;
    memset(&pi, 0, sizeof pi);
    pi.width = width;
    pi.height = height;
    pi.width_padding = pad;
    pi.pixel_format = (enum gpujpeg_pixel_format)fmt;
    struct gj_raw_layout l;
    if ( gj_raw_layout_init(&l, &pi) ) return -1;
    out[0] = l.comp_count;
    out[1] = (long)l.size;
    for ( int c = 0; c < 3; c++ ) {
        out[2 + 5 * c] = (long)l.comp[c].off;
        out[3 + 5 * c] = (long)l.comp[c].pitch;
        out[4 + 5 * c] = l.comp[c].xs;
        out[5 + 5 * c] = l.sampling[c].horizontal;
        out[6 + 5 * c] = l.sampling[c].vertical;
    }
    return 0;
}

/* header + first SOS for any component count, sampling and internal colour space (JFIF / Adobe / SPIFF flavours) */
int shim_header2(int width, int height, int quality, int rst, int interleaved, int comps, int lhs, int lvs, int internal,
                 unsigned char* out)
{
    struct gpujpeg_parameters p;
    struct gpujpeg_image_parameters pi;
    memset(&p, 0, sizeof p);
    memset(&pi, 0, sizeof pi);
    p.quality = quality;
    p.restart_interval = rst;
    p.interleaved = interleaved;
    p.comp_count = comps;
    for ( int c = 0; c < comps; c++ ) {   /* a fourth component (alpha) takes the first one's sampling */
        p.sampling_factor[c].horizontal = (uint8_t)((c == 0 || c == 3) ? lhs : 1);
        p.sampling_factor[c].vertical = (uint8_t)((c == 0 || c == 3) ? lvs : 1);
    }
    p.color_space_internal = (enum gpujpeg_color_space)internal;
    pi.width = width;
    pi.height = height;
    uint8_t raw[2][64];
    struct gj_huff_spec spec[2][2];
    for ( int t = 0; t < 2; t++ ) {
        gj_quant_raw(t, quality, raw[t]);
        for ( int k = 0; k < 2; k++ )
            gj_huff_spec_default(t, k, &spec[t][k]);
    }
    size_t n = gj_write_header(out, &p, &pi, raw, spec, (enum gpujpeg_header_type)g_header_type, &g_extras);
    n += gj_write_sos(out + n, &p, 0);
    return (int)n;
}
