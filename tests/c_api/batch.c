/* A plain C caller of the multi-GPU batch extension (include/gpujpegx.h): encodes `count` gradient frames with one
 * worker per entry of the device list given on the command line, decodes them again, and writes every stream and every
 * decoded frame to files.  usage: batch <w> <h> <count> <outdir> <dev> [<dev> ...] */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <gpujpegx.h>
#include <libgpujpeg/gpujpeg.h>

int main(int argc, char** argv)
{
    if ( argc < 6 ) return 2;
    const int w = atoi(argv[1]), h = atoi(argv[2]), count = atoi(argv[3]);
    const char* dir = argv[4];
    int devices[16], n = 0;
    for ( int i = 5; i < argc && n < 16; i++ )
        devices[n++] = atoi(argv[i]);
    struct gpujpegx_batch* b = gpujpegx_batch_create(devices, n);
    if ( !b ) {
        fprintf(stderr, "no batch coder (no CUDA device?)\n");
        return 3;
    }
    struct gpujpeg_parameters p;
    struct gpujpeg_image_parameters pi;
    gpujpeg_set_default_parameters(&p);
    gpujpeg_image_set_default_parameters(&pi);
    p.quality = 80;
    p.restart_interval = 4;
    pi.width = w;
    pi.height = h;
    const size_t raw = gpujpeg_image_calculate_size(&pi);
    uint8_t** img = calloc((size_t)count, sizeof *img);
    uint8_t** out = calloc((size_t)count, sizeof *out);
    uint8_t** jpeg = calloc((size_t)count, sizeof *jpeg);
    size_t* size = calloc((size_t)count, sizeof *size);
    for ( int f = 0; f < count; f++ ) {
        img[f] = malloc(raw);
        out[f] = malloc(raw);
        for ( int y = 0; y < h; y++ )
            for ( int x = 0; x < w; x++ )
                for ( int c = 0; c < 3; c++ )
                    img[f][((size_t)y * w + x) * 3 + c] = (uint8_t)((y * 255 / h + 17 * f + 40 * c + (x & 7)) & 255);
    }
    int rc = gpujpegx_batch_encode(b, &p, &pi, (const uint8_t* const*)img, count, GPUJPEGX_HOST, jpeg, size);
    if ( rc == 0 ) rc = gpujpegx_batch_decode(b, (const uint8_t* const*)jpeg, size, count, out, GPUJPEGX_HOST);
    for ( int f = 0; f < count && rc == 0; f++ ) {
        char path[512];
        snprintf(path, sizeof path, "%s/%d.jpg", dir, f);
        FILE* fp = fopen(path, "wb");
        fwrite(jpeg[f], 1, size[f], fp);
        fclose(fp);
        snprintf(path, sizeof path, "%s/%d.rgb", dir, f);
        fp = fopen(path, "wb");
        fwrite(out[f], 1, raw, fp);
        fclose(fp);
    }
    printf("%d frames on %d workers, owner of frame 1 = device %d, rc %d\n", count, gpujpegx_batch_device_count(b),
           gpujpegx_batch_owner(b, 1), rc);
    gpujpegx_batch_destroy(b);
    return rc ? 1 : 0;
}
