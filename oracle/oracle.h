/*
 * oracle.h -- CPU restatement of the reference's JPEG hot path.  TEST INFRASTRUCTURE ONLY.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load this library.  Nothing under gpujpeg_b200/ links, imports or calls it.
 *
 * Parity status: the reference's own tests pin no bytes, so this oracle is
 * pinned against the reference *code* instead:
 *   - Huffman encode/decode, integer IDCT, quantisation/Huffman tables and the header writer are
 *     checked byte-for-byte against the reference's own C sources compiled in place into
 *     oracle/_ref/libgpujpeg_refcpu.so (tests/test_oracle_vs_ref.py, golden fixtures in tests/golden/).
 *   - colour transforms and the float forward DCT have no CPU implementation in the reference;
 *     they restate the CUDA kernels' arithmetic and are checked on the GPU box against the
 *     reference GPU library compiled into oracle/_ref/libgpujpeg_refgpu.so (tests/test_ref_gpu.py).
 */
#ifndef GJ_ORACLE_H
#define GJ_ORACLE_H
#include <stddef.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

enum { ORC_IDCT_INT = 0, ORC_IDCT_FLOAT_GPUREF = 1 };

/* zig-zag -> natural index map (JPEG Annex A, fig. A.6) */
extern const uint8_t orc_zigzag_to_natural[64];

/* ---- synthetic inputs ---- */
void orc_gen_random(uint8_t* data, size_t len, int seed);                 /* reference LCG */
void orc_gen_gradient(uint8_t* data, int width, int height, int bpp);     /* .tst default  */
void orc_gen_photo(uint8_t* data, int width, int height, int seed);       /* S-photo, RGB  */

/* ---- tables ---- */
/* raw: zig-zag order u8; fwd: float table indexed [x*8+y] as the FDCT consumes it; inv: natural-order u16 */
void orc_quant_tables(int quality, uint8_t raw[2][64], float fwd[2][64], uint16_t inv[2][64]);
/* Annex K Huffman spec: cls 0 = luminance, 1 = chrominance; kind 0 = DC, 1 = AC */
void orc_huff_spec(int cls, int kind, const uint8_t** bits17, const uint8_t** vals, int* nvals);
/* test hook: replace the table of (cls, kind) in header writer and encoder; bits17 == NULL restores Annex K */
void orc_set_huffman_override(int cls, int kind, const uint8_t* bits17, const uint8_t* vals, int nvals);
/* code/size per symbol (encoder view) */
void orc_huff_encoder_table(int cls, int kind, uint16_t code[256], uint8_t size[256]);

/* ---- stages (4:4:4, 8-bit) ---- */
/* RGB interleaved (row pitch 3*w+pad) -> 3 planes of dw*dh (padding zero-filled) */
void orc_preprocess_rgb444(const uint8_t* rgb, int w, int h, int pad, uint8_t* planes, int dw, int dh);
/* planes -> RGB interleaved */
void orc_postprocess_rgb444(const uint8_t* planes, int dw, int dh, uint8_t* rgb, int w, int h, int pad);
/* one plane -> int16 coefficients, block-major, natural order inside the block */
void orc_fdct_quant_plane(const uint8_t* plane, int dw, int dh, const float fwd[64], int16_t* coef);
/* int16 coefficients (block-major natural) -> plane */
void orc_idct_plane(const int16_t* coef, int dw, int dh, const uint16_t inv[64], int flavour, uint8_t* plane);
/* single block helpers */
void orc_fdct_quant_block(const uint8_t* px, int stride, const float fwd[64], int16_t out[64]);
void orc_idct_int_block(int16_t blk[64], const uint16_t inv[64]);   /* in place, result before +128 */
void orc_idct_float_block(const int16_t in[64], const uint16_t inv[64], uint8_t out[64]);

/* Huffman-encode `nblocks` consecutive blocks (natural order, 64 int16 each) as ONE restart segment:
 * DC predictor starts at 0, output is padded with 1-bits and byte-stuffed, no marker appended.
 * Returns number of bytes written (out must hold >= nblocks*416+8). */
size_t orc_huff_encode_segment(const int16_t* coef, int nblocks, int cls, uint8_t* out);
/* Huffman-decode one restart segment (stuffed bytes, no marker) into nblocks blocks. 0 on success. */
int orc_huff_decode_segment(const uint8_t* data, size_t size, int nblocks,
                            const uint8_t* dc_bits17, const uint8_t* dc_vals,
                            const uint8_t* ac_bits17, const uint8_t* ac_vals, int16_t* coef);

/* ---- whole frames ---- */
/* JPEG header up to (not including) the first SOS; returns length. */
size_t orc_write_header(uint8_t* out, int w, int h, int quality, int rst, int comp_count);
/* Encode RGB 4:4:4 u8 interleaved -> baseline JPEG (non-interleaved scans when interleaved==0).
 * threads<=1: sequential.  Returns bytes written, 0 on error.  out must hold 1000+w*h*3*2 bytes.
 * If coef_out != NULL it receives all quantised coefficients (3 planes, block-major natural). */
size_t orc_encode_rgb(const uint8_t* rgb, int w, int h, int pad, int quality, int rst, int interleaved,
                      int threads, uint8_t* out, int16_t* coef_out);
/* Same with chroma subsampling: luminance sampling factors lhs x lvs in {1,2} (2x2 = 4:2:0, 2x1 = 4:2:2,
 * 1x2 = 4:4:0), chrominance 1x1; chroma is point-sampled as the reference's preprocessor does.
 * coef_out layout: component after component, each block-major natural over ITS block grid. */
size_t orc_encode_rgb_ss(const uint8_t* rgb, int w, int h, int pad, int quality, int rst, int interleaved, int lhs,
                         int lvs, int threads, uint8_t* out, int16_t* coef_out);
/* Raw formats whose samples enter the JPEG without a colour transform (image colour space == internal colour space):
 * fmt = the reference's enum gpujpeg_pixel_format value (0 u8, 1 444-u8-p012, 2 444-u8-p0p1p2, 3 422-u8-p1020,
 * 4 422-u8-p0p1p2, 5 420-u8-p0p1p2); the JPEG's sampling is the format's.  orc_raw_size gives the buffer size. */
size_t orc_raw_size(int fmt, int w, int h, int pad);
size_t orc_encode_ycc(const uint8_t* raw, int w, int h, int pad, int fmt, int quality, int rst, int interleaved,
                      int threads, uint8_t* out, int16_t* coef_out);
int orc_decode_ycc(const uint8_t* jpeg, size_t size, int idct_flavour, int threads, int fmt, int pad, uint8_t* raw);
/* Generic path: any of those pixel formats in colour space cs (the reference's enum gpujpeg_color_space value: 0 none,
 * 1 RGB, 2 BT.601, 3 BT.601 full range = JPEG, 4 BT.709), transformed per pixel to / from the JPEG's YCbCr, with any
 * luminance sampling lhs x lvs in {1,2}.  Widths must be even for the formats that share chroma horizontally. */
size_t orc_encode_any(const uint8_t* raw, int w, int h, int fmt, int cs, int quality, int rst, int interleaved, int lhs,
                      int lvs, int threads, uint8_t* out);
int orc_decode_any(const uint8_t* jpeg, size_t size, int idct_flavour, int threads, int fmt, int cs, uint8_t* raw);
/* the same with the colour space of the JPEG's components chosen: 3 = YCbCr JPEG (JFIF header) or 1 = RGB (Adobe APP14
 * header, component ids 'R','G','B', luminance tables for every component); orc_decode_any detects it from the stream */
size_t orc_encode_any2(const uint8_t* raw, int w, int h, int fmt, int cs, int internal, int quality, int rst, int interleaved,
                       int lhs, int lvs, int threads, uint8_t* out);
/* comp_count = 4 of the reference's parameters: a 4444-u8-p0123 image given to orc_encode_any* keeps its alpha samples as a
 * fourth component (luminance tables, the first component's sampling, SPIFF header) [ref: src/gpujpeg_common.c:692-694,
 * src/gpujpeg_writer.c:458-460]; orc_decode_any hands a fourth component out as alpha */
void orc_set_four_components(int on);
/* quantised coefficients of a stream with 1, 3 or 4 components: returns their number (coef == NULL: count only), 0 on error */
size_t orc_decode_coefficients(const uint8_t* jpeg, size_t size, int16_t* coef);
/* Decode a baseline JPEG produced by this codec family (3 comp, any of the above samplings, or 1 comp) to RGB/gray u8.
 * Returns 0 on success; fills w,h,comps.  rgb may be NULL to probe. coef_out optional. */
int orc_decode_rgb(const uint8_t* jpeg, size_t size, int idct_flavour, int threads, uint8_t* rgb,
                   int* w, int* h, int* comps, int16_t* coef_out);

/* stream structure probe used by tests: counts and offsets of restart segments */
struct orc_stream_info {
    int width, height, comp_count, restart_interval, interleaved, scan_count, segment_count;
    int quality_guess;
    size_t header_size;      /* bytes before first SOS */
    size_t scan_bytes[4];    /* entropy-coded bytes (incl. RST markers) per scan */
};
int orc_probe(const uint8_t* jpeg, size_t size, struct orc_stream_info* info);

#ifdef __cplusplus
}
#endif
#endif
