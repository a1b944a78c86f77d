/*
 * gj_exif.c -- the Exif flavour of the file header (APP1 "Exif\0\0" + big-endian TIFF structure) and the reader's look
 * into foreign Exif headers for the orientation.  Host code only.
 *
 * Restates what the reference's writer puts into the segment [ref: src/gpujpeg_exif.c:286-438]: IFD0 with Orientation,
 * X/YResolution 72/1, ResolutionUnit inches, DateTime (now, local time), YCbCrPositioning centred and the pointer to the
 * Exif IFD; the Exif IFD with ExifVersion "0230", ComponentsConfiguration YCbCr, FlashpixVersion "0100", ColorSpace sRGB,
 * PixelX/YDimension.  User tags (enc_exif_tag option, [ref: src/gpujpeg_exif.c:440-590]) are added to the IFD their number
 * belongs to and replace the default of the same number.
 *
 * Built differently from the reference (fields are collected first, every IFD is laid out by one routine), but the bytes
 * are the reference's, including three things a reader of the file may stumble over and which are kept for parity:
 *   (1) values longer than 4 bytes are placed behind the IFD in the order defaults-then-user-tags, not in tag order;
 *   (2) the Exif IFD pointer is the end of the value area as it was when the pointer field was composed, so a user tag of
 *       IFD0 with a long value moves the Exif IFD away from where the pointer says [ref: src/gpujpeg_exif.c:253-262];
 *   (3) a user tag that replaces a default of the Exif IFD still leaves six default slots, the last default twice
 *       [ref: src/gpujpeg_exif.c:384-386: the shortened count is not the one passed on].
 */
#include <ctype.h>
#include <stdlib.h>
#include <string.h>
#include <strings.h>
#include <time.h>

#include "gj_internal.h"

enum { T_BYTE = 1, T_ASCII = 2, T_SHORT = 3, T_LONG = 4, T_RATIONAL = 5, T_UNDEFINED = 7, T_SLONG = 9, T_SRATIONAL = 10 };
enum { ID_ORIENTATION = 0x112, ID_EXIF_IFD = 0x8769, FIRST_PRIVATE_ID = 0x827A };

static const struct { int type; int unit; const char* name; } k_types[] = {   /* unit: bytes per number */
    {T_BYTE, 1, "BYTE"}, {T_ASCII, 1, "ASCII"}, {T_SHORT, 2, "SHORT"}, {T_LONG, 4, "LONG"}, {T_RATIONAL, 4, "RATIONAL"},
    {T_UNDEFINED, 1, "UNDEFINED"}, {T_SLONG, 4, "SLONG"}, {T_SRATIONAL, 4, "SRATIONAL"},
};
static int type_unit(int type)
{
    for ( unsigned i = 0; i < sizeof k_types / sizeof k_types[0]; i++ )
        if ( k_types[i].type == type ) return k_types[i].unit;
    return 0;
}
static int type_is_rational(int type) { return type == T_RATIONAL || type == T_SRATIONAL; }
static int type_is_bytes(int type) { return type == T_ASCII || type == T_UNDEFINED; }

/* tags that can be given by name [ref: src/gpujpeg_exif.c:117-142]; "Sofware" is the reference's spelling */
static const struct { uint16_t id; int type; const char* name; } k_names[] = {
    {0x112, T_SHORT, "Orientation"}, {0x11A, T_RATIONAL, "XResolution"}, {0x11B, T_RATIONAL, "YResolution"},
    {0x128, T_SHORT, "ResolutionUnit"}, {0x131, T_ASCII, "Sofware"}, {0x131, T_ASCII, "Software"}, {0x132, T_ASCII, "DateTime"},
    {0x13E, T_RATIONAL, "WhitePoint"}, {0x213, T_SHORT, "YCbCrPositioning"}, {ID_EXIF_IFD, T_LONG, "Exif IFD Pointer"},
    {0x9000, T_UNDEFINED, "ExifVersion"}, {0x9101, T_UNDEFINED, "ComponentConfiguration"}, {0xA000, T_UNDEFINED, "FlashPixVersion"},
    {0xA001, T_SHORT, "ColorSpace"}, {0xA002, T_SHORT, "PixelXDimension"}, {0xA003, T_SHORT, "PixelYDimension"},
};

/* one field of an IFD: `bytes` is the value as it stands in the file (big-endian numbers) */
struct field {
    uint16_t id, type;
    uint32_t count;
    uint8_t* bytes;
    uint32_t nbytes;
};
struct gj_exif_tags {
    struct field* f[2];   /* [0] IFD0, [1] Exif IFD */
    int n[2];
    size_t payload;       /* bytes of all values */
};

static void put32(uint8_t* p, uint32_t v) { p[0] = (uint8_t)(v >> 24); p[1] = (uint8_t)(v >> 16); p[2] = (uint8_t)(v >> 8); p[3] = (uint8_t)v; }
static void put16(uint8_t* p, unsigned v) { p[0] = (uint8_t)(v >> 8); p[1] = (uint8_t)v; }

/* numbers -> the file's bytes */
static void pack_numbers(uint8_t* out, const uint32_t* v, int n, int unit)
{
    for ( int i = 0; i < n; i++ )
        for ( int b = 0; b < unit; b++ )
            *out++ = (uint8_t)(v[i] >> (8 * (unit - 1 - b)));
}

void gj_exif_tags_destroy(struct gj_exif_tags* t)
{
    if ( !t ) return;
    for ( int k = 0; k < 2; k++ ) {
        for ( int i = 0; i < t->n[k]; i++ )
            free(t->f[k][i].bytes);
        free(t->f[k]);
    }
    free(t);
}
size_t gj_exif_tags_bytes(const struct gj_exif_tags* t) { return t ? t->payload + 12u * (size_t)(t->n[0] + t->n[1]) : 0; }

static void exif_usage(void)
{
    printf("Exif value syntax:\n"
           "\t" GPUJPEG_ENC_OPT_EXIF_TAG "=<ID>:<type>=<value>\n"
           "\t" GPUJPEG_ENC_OPT_EXIF_TAG "=<name>=<value>\n"
           "\t\tname must be a tag name known to GPUJPEG\n\n"
           "If mulitple numeric values required, separate with a comma; rationals are in format num/den.\n"
           "UNDEFINED and ASCII should be raw strings.\n\nrecognized tag name (type):\n");
    for ( unsigned i = 0; i < sizeof k_names / sizeof k_names[0]; i++ )
        for ( unsigned j = 0; j < sizeof k_types / sizeof k_types[0]; j++ )
            if ( k_types[j].type == k_names[i].type ) printf("\t- %s (%s)\n", k_names[i].name, k_types[j].name);
}

/* "<ID>:<type>=<value>" or "<name>=<value>" -> one more user tag; 0 on success [ref: src/gpujpeg_exif.c:490-590] */
int gj_exif_add_tag(struct gj_exif_tags** tags, const char* cfg)
{
    if ( strcmp(cfg, "help") == 0 ) {
        exif_usage();
        return -1;
    }
    const char* p = cfg;
    long id = 0;
    int type = 0;
    if ( isdigit((unsigned char)*p) ) {
        char* e = NULL;
        id = strtol(p, &e, 0);
        if ( *e != ':' ) {
            GJ_ERR("Error parsing Exif tag ID or missing type!\n");
            return -1;
        }
        p = e + 1;
        size_t best = 0;
        for ( unsigned i = 0; i < sizeof k_types / sizeof k_types[0]; i++ ) {
            const size_t len = strlen(k_types[i].name);
            if ( strncasecmp(p, k_types[i].name, len) == 0 && best == 0 ) {   /* first match in type-number order, as the reference */
                type = k_types[i].type;
                best = len;
            }
        }
        p += best;
        if ( type == 0 ) {
            GJ_ERR("Error parsing Exif tag type!\n");
            return -1;
        }
        if ( *p != '=' ) {
            GJ_ERR("Error parsing Exif - missing value!\n");
            return -1;
        }
    }
    else {
        for ( unsigned i = 0; i < sizeof k_names / sizeof k_names[0]; i++ ) {
            const size_t len = strlen(k_names[i].name);
            if ( strncasecmp(p, k_names[i].name, len) == 0 && p[len] == '=' ) {
                id = k_names[i].id;
                type = k_names[i].type;
                p += len;
                break;
            }
        }
        if ( type == 0 || *p != '=' ) {
            GJ_ERR("[Exif] Wrong tag name or missing value!\n");
            return -1;
        }
    }
    p++;
    struct field f = {(uint16_t)id, (uint16_t)type, 0, NULL, 0};
    if ( type_is_bytes(type) ) {
        const size_t len = strlen(p);
        f.count = (uint32_t)(len + (type == T_ASCII ? 1 : 0));   /* the string's terminator is part of an ASCII value */
        f.bytes = (uint8_t*)calloc(len + 1, 1);
        if ( !f.bytes ) return -1;
        memcpy(f.bytes, p, len);
        f.nbytes = f.count;
    }
    else {
        uint32_t* v = NULL;
        int n = 0;
        const int per = type_is_rational(type) ? 2 : 1;
        char* e = (char*)p;
        do {
            if ( *e == ',' ) e++;
            uint32_t* nv = (uint32_t*)realloc(v, (size_t)(n + per) * sizeof *v);
            if ( !nv ) {
                free(v);
                return -1;
            }
            v = nv;
            v[n++] = (uint32_t)strtoull(e, &e, 0);
            if ( per == 2 ) {
                if ( *e != '/' ) GJ_ERR("[Exif] Malformed rational, expected '/', got '%c'!\n", *e);
                if ( *e ) e++;
                v[n++] = (uint32_t)strtoull(e, &e, 0);
            }
        } while ( *e == ',' );
        if ( *e != '\0' ) {
            free(v);
            GJ_ERR("Trainling data in Exif value: %s\n", e);
            return -1;
        }
        const int unit = type_unit(type);
        f.count = (uint32_t)(n / per);
        f.nbytes = (uint32_t)(n * unit);
        f.bytes = (uint8_t*)malloc(f.nbytes ? f.nbytes : 1);
        if ( !f.bytes ) {
            free(v);
            return -1;
        }
        pack_numbers(f.bytes, v, n, unit);
        free(v);
    }
    if ( !*tags ) *tags = (struct gj_exif_tags*)calloc(1, sizeof **tags);
    struct gj_exif_tags* t = *tags;
    if ( !t || t->payload + f.nbytes + 12u * (size_t)(t->n[0] + t->n[1] + 1) > 60000 ) {
        if ( t ) GJ_ERR("[Exif] The tags do not fit into one APP1 segment.\n");
        free(f.bytes);
        return -1;
    }
    const int k = id < FIRST_PRIVATE_ID ? 0 : 1;
    struct field* nf = (struct field*)realloc(t->f[k], (size_t)(t->n[k] + 1) * sizeof *nf);
    if ( !nf ) {
        free(f.bytes);
        return -1;
    }
    t->f[k] = nf;
    nf[t->n[k]++] = f;
    t->payload += f.nbytes;
    return 0;
}

/* Lays out one IFD at `at`: field count, the 12-byte fields sorted by tag number, "no next IFD", then the long values in
 * the order the fields were given.  Returns the first byte behind the value area. */
static uint8_t* emit_ifd(const uint8_t* tiff, uint8_t* at, const struct field* f, int n)
{
    uint8_t* entry = at + 2;
    uint8_t* values = entry + 12 * n + 4;
    put16(at, (unsigned)n);
    for ( int i = 0; i < n; i++, entry += 12 ) {
        put16(entry, f[i].id);
        put16(entry + 2, f[i].type);
        put32(entry + 4, f[i].count);
        memset(entry + 8, 0, 4);
        if ( f[i].id == ID_EXIF_IFD && f[i].bytes == NULL ) put32(entry + 8, (uint32_t)(values - tiff));   /* quirk (2) */
        else if ( f[i].nbytes <= 4 ) memcpy(entry + 8, f[i].bytes, f[i].nbytes);
        else {
            put32(entry + 8, (uint32_t)(values - tiff));
            memcpy(values, f[i].bytes, f[i].nbytes);
            values += f[i].nbytes;
        }
    }
    /* stable insertion sort of the fields by tag number */
    uint8_t* first = at + 2;
    for ( int i = 1; i < n; i++ ) {
        uint8_t tmp[12];
        memcpy(tmp, first + 12 * i, 12);
        const int id = tmp[0] << 8 | tmp[1];
        int j = i;
        while ( j > 0 && (first[12 * (j - 1)] << 8 | first[12 * (j - 1) + 1]) > id ) {
            memcpy(first + 12 * j, first + 12 * (j - 1), 12);
            j--;
        }
        memcpy(first + 12 * j, tmp, 12);
    }
    put32(first + 12 * n, 0);
    return values;
}

/* defaults without the ones a user tag replaces, then the user tags.  `keep_slots` (quirk (3)): the list is shortened in
 * place, but as many fields as there are defaults are taken from it -- the slots behind the shortened list still hold what
 * was there before the fields in front of them moved up */
static int collect(struct field* out, const struct field* defaults, int ndef, const struct field* user, int nuser, int keep_slots)
{
    struct field d[8];
    int n = ndef;
    memcpy(d, defaults, (size_t)ndef * sizeof *d);
    for ( int u = 0; u < nuser; u++ )
        for ( int j = 0; j < n; j++ )
            if ( d[j].id == user[u].id ) {
                memmove(d + j, d + j + 1, (size_t)(n - j - 1) * sizeof *d);
                n--;
                break;
            }
    if ( keep_slots ) n = ndef;
    memcpy(out, d, (size_t)n * sizeof *d);
    if ( nuser ) memcpy(out + n, user, (size_t)nuser * sizeof *user);
    return n + nuser;
}

unsigned gj_exif_orientation_code(const struct gpujpeg_orientation* o)
{
    /* Exif orientation 1..8 <-> (clockwise quarter turns, mirrored) [ref: src/gpujpeg_exif.c:158-168] */
    static const uint8_t code[4][2] = {{1, 2}, {6, 5}, {3, 4}, {8, 7}};
    return code[o->rotation & 3][o->flip & 1];
}

/* The APP1 segment (marker included); DateTime is the current local time. */
size_t gj_exif_write(uint8_t* out, const struct gpujpeg_parameters* param, const struct gpujpeg_image_parameters* pi,
                     const struct gpujpeg_image_metadata* metadata, const struct gj_exif_tags* tags)
{
    if ( param->color_space_internal != GPUJPEG_YCBCR_BT601_256LVLS )
        GJ_WARN("[Exif] Color space %s currently not recorded, assumed %s (report)\n",
                gpujpeg_color_space_get_name(param->color_space_internal), gpujpeg_color_space_get_name(GPUJPEG_YCBCR_BT601_256LVLS));
    uint8_t* p = out;
    *p++ = 0xFF;
    *p++ = 0xE1;
    uint8_t* const length_at = p;
    p += 2;
    memcpy(p, "Exif\0", 6);   /* identifier and one byte of padding */
    p += 6;
    uint8_t* const tiff = p;
    memcpy(p, "MM", 2);       /* big-endian */
    put16(p + 2, 0x002A);
    put32(p + 4, 8);          /* IFD0 follows */
    p += 8;

    char date_time[20] = "    :  :     :  :  ";
    time_t now = time(NULL);
    struct tm tm;
    if ( localtime_r(&now, &tm) ) strftime(date_time, sizeof date_time, "%Y:%m:%d %H:%M:%S", &tm);
    unsigned orientation = 1;
    if ( metadata && metadata->vals[GPUJPEG_METADATA_ORIENTATION].set )
        orientation = gj_exif_orientation_code(&metadata->vals[GPUJPEG_METADATA_ORIENTATION].orient);
    uint8_t v_orient[2], v_res[8], v_unit[2], v_pos[2], v_cs[2], v_w[2], v_h[2];
    put16(v_orient, orientation);
    put32(v_res, 72);
    put32(v_res + 4, 1);
    put16(v_unit, 2);   /* inches */
    put16(v_pos, 1);    /* centred */
    put16(v_cs, 1);     /* sRGB */
    put16(v_w, (unsigned)pi->width & 0xFFFFu);
    put16(v_h, (unsigned)pi->height & 0xFFFFu);
    const struct field ifd0[7] = {
        {ID_ORIENTATION, T_SHORT, 1, v_orient, 2}, {0x11A, T_RATIONAL, 1, v_res, 8}, {0x11B, T_RATIONAL, 1, v_res, 8},
        {0x128, T_SHORT, 1, v_unit, 2}, {0x132, T_ASCII, 20, (uint8_t*)date_time, 20}, {0x213, T_SHORT, 1, v_pos, 2},
        {ID_EXIF_IFD, T_LONG, 1, NULL, 4},
    };
    const struct field exif[6] = {
        {0x9000, T_UNDEFINED, 4, (uint8_t*)"0230", 4}, {0x9101, T_UNDEFINED, 4, (uint8_t*)"\1\2\3\0", 4},
        {0xA000, T_UNDEFINED, 4, (uint8_t*)"0100", 4}, {0xA001, T_SHORT, 1, v_cs, 2}, {0xA002, T_SHORT, 1, v_w, 2},
        {0xA003, T_SHORT, 1, v_h, 2},
    };
    const int nu0 = tags ? tags->n[0] : 0, nu1 = tags ? tags->n[1] : 0;
    struct field* all = (struct field*)malloc((size_t)(8 + (nu0 > nu1 ? nu0 : nu1)) * sizeof *all);
    if ( !all ) return 0;
    int n = collect(all, ifd0, 7, tags ? tags->f[0] : NULL, nu0, 0);
    p = emit_ifd(tiff, p, all, n);
    n = collect(all, exif, 6, tags ? tags->f[1] : NULL, nu1, 1);
    p = emit_ifd(tiff, p, all, n);
    free(all);
    put16(length_at, (unsigned)(p - length_at));
    return (size_t)(p - out);
}

/* ---- reader: the orientation out of IFD0 [ref: src/gpujpeg_exif.c:640-764] ---- */
static unsigned rd16(const uint8_t* p, int le) { return le ? (unsigned)(p[1] << 8 | p[0]) : (unsigned)(p[0] << 8 | p[1]); }
static uint32_t rd32(const uint8_t* p, int le)
{
    return le ? (uint32_t)p[3] << 24 | (uint32_t)p[2] << 16 | (uint32_t)p[1] << 8 | p[0]
              : (uint32_t)p[0] << 24 | (uint32_t)p[1] << 16 | (uint32_t)p[2] << 8 | p[3];
}

/* `seg` = the APP1 payload behind the length field ("Exif\0\0" ...), `len` bytes of it; `file_end` bounds IFD0, which the
 * reference lets lie anywhere in the rest of the file */
void gj_exif_parse(const uint8_t* seg, size_t len, const uint8_t* file_end, int verbose, struct gpujpeg_image_metadata* metadata)
{
    if ( len + 2 < 18 ) {
        GJ_WARN("Insufficient Exif header length %u!\n", (unsigned)(len + 2));
        return;
    }
    const uint8_t* tiff = seg + 6;
    int le;
    if ( tiff[0] == 'I' && tiff[1] == 'I' ) le = 1;
    else if ( tiff[0] == 'M' && tiff[1] == 'M' ) le = 0;
    else {
        GJ_WARN("Unexpected endianity!\n");
        return;
    }
    GJ_DEBUG(verbose, "%s endian Exif detected.\n", le ? "Little" : "Big");
    if ( rd16(tiff + 2, le) != 0x002A ) {
        GJ_WARN("Wrong TIFF tag, expected 0x%04x!\n", 0x002A);
        return;
    }
    const uint32_t off = rd32(tiff + 4, le);
    if ( (size_t)(file_end - tiff) < (size_t)off + 2 ) {
        GJ_WARN("Unexpected end of file!\n");
        return;
    }
    const uint8_t* p = tiff + off;
    const unsigned n = rd16(p, le);
    p += 2;
    if ( (size_t)(file_end - p) < (size_t)n * 12 ) {
        GJ_WARN("[Exif] Insufficient space to hold %u IFD0 items!\n", n);
        return;
    }
    GJ_DEBUG(verbose, "Found %u IFD0 items.\n", n);
    for ( unsigned i = 0; i < n; i++, p += 12 ) {
        const unsigned id = rd16(p, le), type = rd16(p + 2, le);
        uint32_t val = rd32(p + 8, le);
        /* a number shorter than the slot stands at its left end: for big-endian files it is shifted down by its own
         * width, as the reference does it (right for SHORT, which is what the orientation is) */
        if ( !le && (type == T_BYTE || type == T_SHORT) ) val >>= 8 * type_unit((int)type);
        GJ_DEBUG(verbose, "[Exif] Found IFD0 tag %#x type %u: count=%u, value/offset=%#x\n", id, type, rd32(p + 4, le), val);
        if ( id != ID_ORIENTATION ) continue;
        if ( val == 0 || val > 8 ) {
            GJ_WARN("[Exif] Flawed orientation value %d! Should be 1-8...\n", (int)val);
            continue;
        }
        static const uint8_t rot[8] = {0, 0, 2, 2, 1, 1, 3, 3}, flip[8] = {0, 1, 0, 1, 1, 0, 1, 0};
        metadata->vals[GPUJPEG_METADATA_ORIENTATION].orient.rotation = rot[val - 1];
        metadata->vals[GPUJPEG_METADATA_ORIENTATION].orient.flip = flip[val - 1];
        metadata->vals[GPUJPEG_METADATA_ORIENTATION].set = 1;
    }
}
