// Host build of the per-thread device arithmetic (gpujpeg_b200/csrc/gj_device.cuh) so the build
// container, which has no GPU, can check it against the oracle.  Test infrastructure only: the
// product never runs this code on the CPU.  Build: g++ -O2 -ffp-contract=off -shared -fPIC.
#include <cstdint>
#include <cstring>

#include "../../gpujpeg_b200/csrc/gj_device.cuh"

extern "C" {

// exhaustive check of the float colour transform against the integer definition; returns mismatches
long km_check_rgb_to_ycbcr_exhaustive(int r_lo, int r_hi)
{
    long bad = 0;
    for ( int R = r_lo; R < r_hi; R++ )
        for ( int G = 0; G < 256; G++ )
            for ( int B = 0; B < 256; B++ ) {
                const int r = R * 256 / 255, g = G * 256 / 255, b = B * 256 / 255;
                auto c8 = [](int v) { return v < 0 ? 0 : v > 255 ? 255 : v; };
                const int Y = c8(((77 * r + 150 * g + 29 * b + 128) >> 8) + 0);
                const int Cb = c8(((-43 * r - 85 * g + 128 * b + 128) >> 8) + 128);
                const int Cr = c8(((128 * r - 107 * g - 21 * b + 128) >> 8) + 128);
                float y, cb, cr;
                gj_rgb_to_ycbcr((float)R, (float)G, (float)B, y, cb, cr);
                if ( y != (float)Y || cb != (float)Cb || cr != (float)Cr ) bad++;
            }
    return bad;
}

long km_check_ycbcr_to_rgb_exhaustive(int y_lo, int y_hi)
{
    long bad = 0;
    for ( int Y = y_lo; Y < y_hi; Y++ )
        for ( int Cb = 0; Cb < 256; Cb++ )
            for ( int Cr = 0; Cr < 256; Cr++ ) {
                const int y = (Y - 0) * 256 / 255, cb = (Cb - 128) * 256 / 255, cr = (Cr - 128) * 256 / 255;
                auto c8 = [](int v) { return v < 0 ? 0 : v > 255 ? 255 : v; };
                const int R = c8((256 * y + 0 * cb + 359 * cr + 128) >> 8);
                const int G = c8((256 * y - 88 * cb - 183 * cr + 128) >> 8);
                const int B = c8((256 * y + 454 * cb + 0 * cr + 128) >> 8);
                int r, g, b;
                gj_ycbcr_to_rgb(Y, Cb, Cr, r, g, b);
                if ( r != R || g != G || b != B ) bad++;
            }
    return bad;
}

// plane (u8, dw x dh) -> coefficients in NATURAL order, block-major, using the kernel's own block
// routine and a zig-zag ordered forward table exactly as K1 does
void km_fdct_quant_plane(const uint8_t* plane, int dw, int dh, const float* fwd_zz, int16_t* coef)
{
    const int bcx = dw / 8, bcy = dh / 8;
    for ( int by = 0; by < bcy; by++ )
        for ( int bx = 0; bx < bcx; bx++ ) {
            float v[64];
            for ( int i = 0; i < 64; i++ )
                v[i] = (float)plane[(size_t)(by * 8 + i / 8) * dw + bx * 8 + i % 8];
            gj_fdct_block(v);
            int16_t* out = coef + ((size_t)by * bcx + bx) * 64;
            for ( int k = 0; k < 64; k++ ) {
                const int n = gj_zz2nat(k);
                out[n] = (int16_t)(gj_quant_bits(v[n], fwd_zz[k]) & 0xFFFFu);   // the kernel's own rounding (no conversion instruction)
            }
        }
}

// coefficients (natural order) -> plane with the kernel's integer / float IDCT, q in zig-zag order
void km_idct_plane(const int16_t* coef, int dw, int dh, const uint16_t* q_zz, int flavour, uint8_t* plane)
{
    const int bcx = dw / 8, bcy = dh / 8;
    for ( int by = 0; by < bcy; by++ )
        for ( int bx = 0; bx < bcx; bx++ ) {
            const int16_t* in = coef + ((size_t)by * bcx + bx) * 64;
            uint8_t px[64];
            if ( flavour == 0 ) {
                int v[64];
                for ( int k = 0; k < 64; k++ ) {
                    const int n = gj_zz2nat(k);
                    v[n] = gj_s16((int)in[n] * (int)(short)q_zz[k]);
                }
                int u[64];
                for ( int i = 0; i < 64; i++ )
                    u[i] = v[i];
                gj_idct_int_block(v);      // the reference's own formulation ...
                gj_idct_int_block_px(u);   // ... and the fused-level-shift variant the kernel runs
                for ( int i = 0; i < 64; i++ ) {
                    px[i] = (uint8_t)gj_clamp8(u[i]);
                    if ( px[i] != (uint8_t)gj_clamp8(gj_s16(v[i] + 128)) ) px[i] ^= 0x55;   // force a test failure
                }
            }
            else {
                float f[64];
                for ( int k = 0; k < 64; k++ ) {
                    const int n = gj_zz2nat(k);
                    f[n] = (float)((int)in[n] * (int)q_zz[k]);
                }
                gj_idct_float_block(f);
                for ( int i = 0; i < 64; i++ )
                    px[i] = (uint8_t)gj_clamp8(GJ_RINT(GJ_FADD(f[i], 128.0f)));
            }
            for ( int i = 0; i < 64; i++ )
                plane[(size_t)(by * 8 + i / 8) * dw + bx * 8 + i % 8] = px[i];
        }
}

int km_zigzag_tables_consistent(void)
{
    for ( int k = 0; k < 64; k++ )
        if ( gj_nat2zz(gj_zz2nat(k)) != k ) return 0;
    return 1;
}

int km_category(int v) { return gj_category(v); }
unsigned km_value_bits(int v, int size) { return gj_value_bits(v, size); }
int km_extend(int bits, int size) { return gj_extend(bits, size); }
}
