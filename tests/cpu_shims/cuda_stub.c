/* Host stand-ins for the thin CUDA wrappers of gj_cuda_util.cu, so that gj_common.c + gj_imageio.c (image file helpers,
 * name tables, parameter defaults) can be exercised by the CPU tests.  Test infrastructure: never part of the product. */
#include <stdlib.h>
#include <string.h>

#include "../../gpujpeg_b200/csrc/gj_internal.h"

int gj_cuda_malloc(void** p, size_t size) { *p = malloc(size); return *p ? 0 : -1; }
int gj_cuda_free(void* p) { free(p); return 0; }
int gj_cuda_malloc_host(void** p, size_t size) { *p = malloc(size); return *p ? 0 : -1; }
int gj_cuda_free_host(void* p) { free(p); return 0; }
int gj_cuda_memcpy_h2d_async(void* dst, const void* src, size_t size, gj_stream_t s) { (void)s; memcpy(dst, src, size); return 0; }
int gj_cuda_stream_sync(gj_stream_t s) { (void)s; return 0; }
const char* gj_cuda_last_error(void) { return "host stub"; }
int gj_cuda_device_count(void) { return 0; }
int gj_cuda_device_props(int dev, struct gpujpeg_device_info* info) { (void)dev; (void)info; return -1; }
int gj_cuda_set_device(int dev) { (void)dev; return -1; }
void gj_cuda_device_reset(void) {}
