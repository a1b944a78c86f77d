/*
 * gj_imageio.c -- image files next to the codec: Netpbm (PGM / PPM / PNM raw variants, PAM) and YUV4MPEG2 single frames.
 *
 * What the reference does through its delegates [ref: src/utils/image_delegate.c:149-340 (probe / load / save rules),
 * src/utils/pam.c:47-294, src/utils/y4m.c:42-206] restated for this library: a file is described once by
 * gj_imgfile_probe() from its first bytes (header parsed out of a memory buffer, no stdio scanning), loaded with a
 * caller-supplied allocator (the public API passes the pinned-memory one, tests pass malloc) and written with headers
 * that are byte-for-byte the reference's, so that files made by either library are interchangeable.
 * Host code only; nothing here touches the device.  BMP / PNG / TGA / GIF (third-party stb code in the reference) are
 * not taken by this build.
 */
#include <ctype.h>
#include <errno.h>
#include <stdlib.h>
#include <string.h>

#include "gj_internal.h"

enum { HEAD_MAX = 4096 };   /* longest header looked at (comments included) */

/* ---- a cursor over the first bytes of the file ---- */
struct cursor {
    const char* p;
    const char* end;
};
static int cur_peek(const struct cursor* c) { return c->p < c->end ? (unsigned char)*c->p : -1; }
static void cur_skip_space(struct cursor* c)
{
    while ( c->p < c->end && isspace((unsigned char)*c->p) ) c->p++;
}
/* decimal integer with optional sign, leading white space skipped (scanf %d); 0 = none there */
static int cur_int(struct cursor* c, int* out)
{
    cur_skip_space(c);
    const char* q = c->p;
    int neg = 0;
    if ( q < c->end && (*q == '-' || *q == '+') ) neg = *q++ == '-';
    if ( q >= c->end || !isdigit((unsigned char)*q) ) return 0;
    long long v = 0;
    while ( q < c->end && isdigit((unsigned char)*q) ) {
        if ( v < (1ll << 40) ) v = v * 10 + (*q - '0');
        q++;
    }
    if ( v > 0x7FFFFFFF ) v = 0x7FFFFFFF;
    *out = (int)(neg ? -v : v);
    c->p = q;
    return 1;
}
/* next white-space separated word (scanf " %128s"); 0 at the end of the buffer */
static int cur_word(struct cursor* c, char* out, size_t cap)
{
    cur_skip_space(c);
    size_t n = 0;
    while ( c->p < c->end && !isspace((unsigned char)*c->p) ) {
        if ( n + 1 < cap ) out[n++] = *c->p;
        c->p++;
    }
    out[n] = '\0';
    return n != 0;
}

/* ---- Netpbm ---- */
/* "P5" / "P6": width, height, maxval in decimal, '#' comments between them, then exactly one '\n' before the samples
 * [ref: src/utils/pam.c:90-143]; plain (ASCII) variants and bitmaps are refused */
static int parse_pnm(struct cursor* c, char id, struct gj_imgfile* f)
{
    switch ( id ) {
        case '5': f->channels = 1; break;
        case '6': f->channels = 3; break;
        case '1': case '2': case '3':
            fprintf(stderr, "Plain (ASCII) PNM are not supported, input is P%c\n", id);
            return -1;
        case '4':
            fprintf(stderr, "Bitmap PBM (P4) images are not supported\n");
            return -1;
        default:
            fprintf(stderr, "Wrong PNM type P%c\n", id);
            return -1;
    }
    int* const item[3] = {&f->width, &f->height, &f->maxval};
    int have = 0;
    while ( have < 3 ) {
        if ( cur_int(c, item[have]) ) {
            have++;
            continue;
        }
        if ( cur_peek(c) != '#' ) break;   /* (white space was skipped by cur_int) */
        while ( c->p < c->end && *c->p != '\n' ) c->p++;
    }
    if ( have < 3 ) {
        fprintf(stderr, "Problem parsing PNM header, number of hdr items successfully read: %d\n", have);
        return -1;
    }
    if ( cur_peek(c) != '\n' ) {
        fprintf(stderr, "PNM maximal value isn't immediately followed by <NL>\n");
        return -1;
    }
    c->p++;
    return 0;
}

/* "P7": KEY value lines up to ENDHDR; TUPLTYPE is not needed, DEPTH tells the pixel format [ref: src/utils/pam.c:47-79] */
static int parse_pam(struct cursor* c, struct gj_imgfile* f)
{
    while ( c->p < c->end ) {
        const char* nl = memchr(c->p, '\n', (size_t)(c->end - c->p));
        if ( !nl ) break;
        const char* line = c->p;
        const size_t len = (size_t)(nl - line);
        c->p = nl + 1;
        if ( len == 6 && memcmp(line, "ENDHDR", 6) == 0 ) return 0;
        if ( len == 0 || line[0] == '#' ) {
            if ( len == 0 ) break;   /* a line without a key ends the header in the reference as well */
            continue;
        }
        const char* sp = memchr(line, ' ', len);
        if ( !sp ) break;
        const size_t klen = (size_t)(sp - line);
        const int val = atoi(sp + 1);   /* stops at the newline */
        if ( klen == 5 && memcmp(line, "WIDTH", 5) == 0 ) f->width = val;
        else if ( klen == 6 && memcmp(line, "HEIGHT", 6) == 0 ) f->height = val;
        else if ( klen == 5 && memcmp(line, "DEPTH", 5) == 0 ) f->channels = val;
        else if ( klen == 6 && memcmp(line, "MAXVAL", 6) == 0 ) f->maxval = val;
        else if ( klen == 8 && memcmp(line, "TUPLTYPE", 8) == 0 ) { /* implied by DEPTH */ }
        else fprintf(stderr, "unrecognized key %.*s in PAM header\n", (int)klen, line);
    }
    return 0;   /* the size checks of the caller catch a header that ended early */
}

/* ---- YUV4MPEG2 ---- */
/* chroma tag: "mono[N]", "444alpha", "<sss>[p<N>][suffix]" [ref: src/utils/y4m.c:42-60] */
static int y4m_chroma(const char* tag, struct gj_imgfile* f)
{
    f->bitdepth = 8;
    if ( strcmp(tag, "444alpha") == 0 ) {
        f->subsampling = 4444;
        return 0;
    }
    if ( strncmp(tag, "mono", 4) == 0 ) {
        f->subsampling = 400;
        if ( isdigit((unsigned char)tag[4]) ) f->bitdepth = atoi(tag + 4);
        return 0;
    }
    char* e = NULL;
    const long ss = strtol(tag, &e, 10);
    if ( e == tag ) {
        fprintf(stderr, "Y4M: unable to parse chroma type\n");
        return -1;
    }
    f->subsampling = (int)ss;
    if ( *e == 'p' && isdigit((unsigned char)e[1]) ) f->bitdepth = atoi(e + 1);
    return 0;
}
static size_t y4m_payload(const struct gj_imgfile* f)
{
    const size_t w = (size_t)f->width, h = (size_t)f->height, cw = (w + 1) / 2, ch = (h + 1) / 2;
    size_t n;
    switch ( f->subsampling ) {   /* [ref: src/utils/y4m.c:62-76] */
        case 400: n = w * h; break;
        case 420: n = w * h + 2 * cw * ch; break;
        case 422: n = w * h + 2 * cw * h; break;
        case 444: n = w * h * 3; break;
        case 4444: n = w * h * 4; break;
        default:
            fprintf(stderr, "Unsupported subsampling '%d'\n", f->subsampling);
            return 0;
    }
    return n * (f->bitdepth > 8 ? 2 : 1);
}
/* stream header words up to FRAME, which a newline must follow directly [ref: src/utils/y4m.c:102-139] */
static int parse_y4m(struct cursor* c, struct gj_imgfile* f)
{
    char word[129];
    while ( cur_word(c, word, sizeof word) && strcmp(word, "FRAME") != 0 ) {
        switch ( word[0] ) {
            case 'W': f->width = atoi(word + 1); break;
            case 'H': f->height = atoi(word + 1); break;
            case 'C':
                if ( y4m_chroma(word + 1, f) ) return -1;
                break;
            case 'X':
                if ( strcmp(word, "XCOLORRANGE=LIMITED") == 0 ) f->limited = 1;
                break;
            default: break;   /* frame rate, interlacing, aspect: not needed */
        }
    }
    if ( cur_peek(c) != '\n' ) return -1;
    c->p++;
    return 0;
}

/* Describes a PGM / PPM / PNM / PAM / Y4M file from its header.  0 on success (f->data_offset and f->data_bytes then
 * say where the samples are), -1 with a message otherwise. */
int gj_imgfile_probe(const char* filename, struct gj_imgfile* f)
{
    memset(f, 0, sizeof *f);
    FILE* file = fopen(filename, "rb");
    if ( !file ) {
        fprintf(stderr, "Failed to open %s: %s\n", filename, strerror(errno));
        return -1;
    }
    char head[HEAD_MAX];
    const size_t got = fread(head, 1, sizeof head, file);
    fclose(file);
    struct cursor c = {head, head + got};
    int rc = -1;
    if ( got >= 10 && memcmp(head, "YUV4MPEG2", 9) == 0 && isspace((unsigned char)head[9]) ) {
        f->kind = GJ_IMGFILE_Y4M;
        c.p += 9;
        rc = parse_y4m(&c, f);
        if ( rc == 0 ) {
            if ( f->width > 0 && f->height > 0 && f->width <= 65535 && f->height <= 65535 ) {
                f->data_bytes = y4m_payload(f);
                if ( f->data_bytes == 0 ) rc = -1;
            }
        }
        else fprintf(stderr, "File '%s' doesn't seem to be valid Y4M.\n", filename);
    }
    else if ( got >= 3 && head[0] == 'P' && head[1] == '7' && head[2] == '\n' ) {
        f->kind = GJ_IMGFILE_PAM;
        c.p += 3;
        rc = parse_pam(&c, f);
    }
    else if ( got >= 3 && head[0] == 'P' && isspace((unsigned char)head[2]) ) {
        f->kind = GJ_IMGFILE_PNM;
        c.p += 3;
        rc = parse_pnm(&c, head[1], f);
    }
    else {
        fprintf(stderr, "File '%s' doesn't seem to be valid PAM, PNM or Y4M.\n", filename);
    }
    if ( rc ) return -1;
    /* nothing a JPEG could hold is larger (and the byte counts below stay far from overflowing) */
    if ( f->width > 65535 || f->height > 65535 || f->channels > 4 ) {
        fprintf(stderr, "Image %s is too large (%dx%d, %d channels)!\n", filename, f->width, f->height, f->channels);
        return -1;
    }
    if ( f->kind == GJ_IMGFILE_Y4M && (f->width <= 0 || f->height <= 0) ) {
        fprintf(stderr, "Unspecified/incorrect size %dx%d!\n", f->width, f->height);
        return -1;
    }
    if ( f->kind != GJ_IMGFILE_Y4M ) {   /* [ref: src/utils/pam.c:193-205] */
        if ( f->width <= 0 || f->height <= 0 ) {
            fprintf(stderr, "Unspecified/incorrect size %dx%d!\n", f->width, f->height);
            rc = -1;
        }
        if ( f->channels <= 0 ) {
            fprintf(stderr, "Unspecified/incorrect channel count %d!\n", f->channels);
            rc = -1;
        }
        if ( f->maxval <= 0 || f->maxval > 65535 ) {
            fprintf(stderr, "Unspecified/incorrect maximal value %d!\n", f->maxval);
            rc = -1;
        }
        if ( rc ) return -1;
        f->data_bytes = (size_t)f->channels * (size_t)f->width * (size_t)f->height * (f->maxval > 255 ? 2 : 1);
    }
    f->data_offset = (size_t)(c.p - head);
    return 0;
}

/* The image parameters a probed file stands for [ref: src/utils/image_delegate.c:170-205 (PAM/PNM), :253-298 (Y4M)];
 * only 8-bit samples are taken, as in the reference. */
int gj_imgfile_params(const char* filename, const struct gj_imgfile* f, struct gpujpeg_image_parameters* pi)
{
    if ( f->kind == GJ_IMGFILE_Y4M ) {
        pi->width = f->width;
        pi->height = f->height;
        if ( f->bitdepth != 8 ) {
            GJ_ERR("Currently only 8-bit Y4M pictures are supported but %s has %d bits!\n", filename, f->bitdepth);
            return GPUJPEG_ERROR;
        }
        switch ( f->subsampling ) {
            case 400: pi->pixel_format = GPUJPEG_U8; break;
            case 420: pi->pixel_format = GPUJPEG_420_U8_P0P1P2; break;
            case 422: pi->pixel_format = GPUJPEG_422_U8_P0P1P2; break;
            case 444: pi->pixel_format = GPUJPEG_444_U8_P0P1P2; break;
            case 4444: GJ_ERR("[y4m] Planar YCbCr with alpha is not currently supported!\n"); return -1;
            default: GJ_ERR("Unknown subsamplig in Y4M!\n"); return GPUJPEG_ERROR;
        }
        pi->color_space = f->limited ? GPUJPEG_YCBCR_BT601 : GPUJPEG_YCBCR_BT601_256LVLS;
        return 0;
    }
    if ( f->maxval != 255 ) {
        GJ_ERR("PAM/PNM image %s reports %d levels but only 255 are currently supported!\n", filename, f->maxval);
        return GPUJPEG_ERROR;
    }
    pi->width = f->width;
    pi->height = f->height;
    pi->color_space = GPUJPEG_RGB;
    switch ( f->channels ) {
        case 4: pi->pixel_format = GPUJPEG_4444_U8_P0123; break;
        case 3: pi->pixel_format = GPUJPEG_444_U8_P012; break;
        case 1:
            pi->color_space = GPUJPEG_YCBCR_BT601_256LVLS;
            pi->pixel_format = GPUJPEG_U8;
            break;
        default: GJ_ERR("Unsupported PAM/PNM component count %d!\n", f->channels); return GPUJPEG_ERROR;
    }
    return 0;
}

/* Loads the samples of a PGM / PPM / PNM / PAM / Y4M file into a buffer obtained from `alloc`. */
int gj_imgfile_load(const char* filename, uint8_t** image, size_t* image_size, void* (*alloc)(size_t))
{
    struct gj_imgfile f;
    if ( gj_imgfile_probe(filename, &f) ) return -1;
    if ( f.kind != GJ_IMGFILE_Y4M && f.maxval != 255 ) {
        GJ_ERR("PAM/PNM image %s reports %d levels but only 255 are currently supported!\n", filename, f.maxval);
        return -1;
    }
    FILE* file = fopen(filename, "rb");
    if ( !file ) {
        fprintf(stderr, "Failed to open %s: %s\n", filename, strerror(errno));
        return -1;
    }
    uint8_t* data = (uint8_t*)alloc(f.data_bytes);
    if ( !data ) {
        fprintf(stderr, "Failed to allocate data!\n");
        fclose(file);
        return -1;
    }
    errno = 0;
    size_t got = 0;
    if ( fseek(file, (long)f.data_offset, SEEK_SET) == 0 ) got = fread(data, 1, f.data_bytes, file);
    fclose(file);
    if ( got != f.data_bytes ) {
        fprintf(stderr, "Unable to load image data from file %s - read %zu B, expected %zu B: %s\n", filename, got,
                f.data_bytes, errno ? strerror(errno) : "EOF");
        /* the buffer stays with the caller's allocator, as in the reference */
        return -1;
    }
    *image = data;
    *image_size = f.data_bytes;
    return 0;
}

/* ---- writers: headers byte-for-byte the reference's [ref: src/utils/pam.c:225-294, src/utils/y4m.c:158-206] ---- */
static int write_file(const char* filename, const char* header, const uint8_t* data, size_t len, const char* what)
{
    errno = 0;
    FILE* file = fopen(filename, "wb");
    if ( !file ) {
        fprintf(stderr, "Failed to open %s for writing: %s\n", filename, strerror(errno));
        return -1;
    }
    fputs(header, file);
    const size_t put = fwrite(data, 1, len, file);
    if ( put != len ) fprintf(stderr, "Unable to write %s data - length %zd, written %zd: %s\n", what, len, put, strerror(errno));
    fclose(file);
    return put == len ? 0 : -1;
}

/* PAM (pnm == 0) or PNM (pnm != 0) from a packed image without subsampling [ref: src/utils/image_delegate.c:207-250] */
int gj_imgfile_save_pam(const char* filename, const struct gpujpeg_image_parameters* pi, const uint8_t* data, int pnm)
{
    if ( pi->pixel_format != GPUJPEG_U8 && pi->color_space != GPUJPEG_RGB ) {
        GJ_ERR("Wrong color space %s for PAM!\n", gpujpeg_color_space_get_name(pi->color_space));
        return -1;
    }
    int depth;
    switch ( pi->pixel_format ) {
        case GPUJPEG_U8: depth = 1; break;
        case GPUJPEG_444_U8_P012: depth = 3; break;
        case GPUJPEG_4444_U8_P0123: depth = 4; break;
        default:
            GJ_ERR("Wrong pixel format %s for PAM/PNM! Only packed formats without subsampling are supported.\n",
                   gpujpeg_pixel_format_get_name(pi->pixel_format));
            return -1;
    }
    char header[160];
    if ( pnm ) {
        if ( depth != 1 && depth != 3 ) {
            fprintf(stderr, "Only 1 or 3 channels supported for PNM!\n");
            /* the reference has created the (empty) file by now; so does this build */
            FILE* file = fopen(filename, "wb");
            if ( file ) fclose(file);
            return -1;
        }
        snprintf(header, sizeof header, "P%d\n%u %u\n%d\n", depth == 1 ? 5 : 6, (unsigned)pi->width, (unsigned)pi->height, 255);
    }
    else {
        static const char* const tuple[5] = {"INVALID", "GRAYSCALE", "GRAYSCALE_ALPHA", "RGB", "RGB_ALPHA"};
        snprintf(header, sizeof header, "P7\nWIDTH %u\nHEIGHT %u\nDEPTH %d\nMAXVAL %d\nTUPLTYPE %s\nENDHDR\n", (unsigned)pi->width,
                 (unsigned)pi->height, depth, 255, tuple[depth]);
    }
    return write_file(filename, header, data, (size_t)pi->width * (size_t)pi->height * (size_t)depth, "PAM/PNM");
}

/* one YUV4MPEG2 frame from a planar YCbCr image [ref: src/utils/image_delegate.c:309-340] */
int gj_imgfile_save_y4m(const char* filename, const struct gpujpeg_image_parameters* pi, const uint8_t* data)
{
    if ( pi->color_space == GPUJPEG_RGB ) {
        GJ_ERR("Y4M cannot use RGB colorspace!\n");
        return -1;
    }
    struct gj_imgfile f;
    memset(&f, 0, sizeof f);
    f.width = pi->width;
    f.height = pi->height;
    f.bitdepth = 8;
    const char* chroma;
    switch ( pi->pixel_format ) {
        case GPUJPEG_U8: f.subsampling = 400; chroma = "mono"; break;
        case GPUJPEG_420_U8_P0P1P2: f.subsampling = 420; chroma = "420"; break;
        case GPUJPEG_422_U8_P0P1P2: f.subsampling = 422; chroma = "422"; break;
        case GPUJPEG_444_U8_P0P1P2: f.subsampling = 444; chroma = "444"; break;
        default:
            GJ_ERR("Wrong pixel format %s for Y4M! Only planar formats are supported.\n",
                   gpujpeg_pixel_format_get_name(pi->pixel_format));
            return -1;
    }
    /* everything but full-range BT.601 is announced as limited range, as the reference does */
    const int limited = pi->color_space != GPUJPEG_YCBCR_BT601_256LVLS;
    char header[160];
    snprintf(header, sizeof header, "YUV4MPEG2 W%d H%d F25:1 Ip A0:0 C%s XCOLORRANGE=%s\nFRAME\n", pi->width, pi->height, chroma,
             limited ? "LIMITED" : "FULL");
    return write_file(filename, header, data, y4m_payload(&f), "Y4M");
}
