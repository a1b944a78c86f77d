/* A C caller of the library, written against <libgpujpeg/gpujpeg.h> exactly as an application of the reference
 * would be (compare the reference's examples/encode_minimal.c and decode_minimal.c): encodes a gradient image as
 * 4:2:0 interleaved JPEG, decodes it again, writes both results.
 *   roundtrip <width> <height> <out.jpg> <out.rgb>      exit 0 ok, 3 no CUDA device, 1 error */
#include <libgpujpeg/gpujpeg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

int main(int argc, char** argv)
{
    if ( argc < 5 ) return 1;
    const int w = atoi(argv[1]), h = atoi(argv[2]);
    if ( gpujpeg_init_device(0, 0) != 0 ) return 3;

    struct gpujpeg_parameters param;
    gpujpeg_set_default_parameters(&param);
    param.quality = 80;
    param.interleaved = 1;
    param.restart_interval = 4;
    gpujpeg_parameters_chroma_subsampling(&param, GPUJPEG_SUBSAMPLING_420);
    struct gpujpeg_image_parameters pi;
    gpujpeg_image_set_default_parameters(&pi);
    pi.width = w;
    pi.height = h;

    uint8_t* rgb = (uint8_t*)malloc((size_t)w * h * 3);
    for ( int y = 0; y < h; y++ )
        memset(rgb + (size_t)y * w * 3, y * 255 / h, (size_t)w * 3);

    struct gpujpeg_encoder* enc = gpujpeg_encoder_create(NULL);
    if ( !enc ) return 3;
    struct gpujpeg_encoder_input in = gpujpeg_encoder_input_image(rgb);
    uint8_t* jpeg = NULL;
    size_t jpeg_size = 0;
    if ( gpujpeg_encoder_encode(enc, &param, &pi, &in, &jpeg, &jpeg_size) != 0 ) return 1;
    FILE* f = fopen(argv[3], "wb");
    if ( !f || fwrite(jpeg, 1, jpeg_size, f) != jpeg_size ) return 1;
    fclose(f);

    struct gpujpeg_decoder* dec = gpujpeg_decoder_create(NULL);
    if ( !dec ) return 1;
    struct gpujpeg_decoder_output out;
    gpujpeg_decoder_output_set_default(&out);
    if ( gpujpeg_decoder_decode(dec, jpeg, jpeg_size, &out) != 0 ) return 1;
    if ( out.param_image.width != w || out.param_image.height != h || out.data_size != (size_t)w * h * 3 ) return 1;
    f = fopen(argv[4], "wb");
    if ( !f || fwrite(out.data, 1, out.data_size, f) != out.data_size ) return 1;
    fclose(f);

    struct gpujpeg_image_parameters info_pi;
    struct gpujpeg_parameters info_p;
    int segments = 0;
    gpujpeg_set_default_parameters(&info_p);
    if ( gpujpeg_decoder_get_image_info(jpeg, jpeg_size, &info_pi, &info_p, &segments) != 0 ) return 1;
    printf("%dx%d %s %d segments, %zu bytes\n", info_pi.width, info_pi.height,
           gpujpeg_subsampling_get_name(info_p.comp_count, info_p.sampling_factor), segments, jpeg_size);

    gpujpeg_decoder_destroy(dec);
    gpujpeg_encoder_destroy(enc);
    free(rgb);
    return 0;
}
