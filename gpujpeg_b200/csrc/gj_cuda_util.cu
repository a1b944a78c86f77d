/*
 * gj_cuda_util.cu -- the few CUDA-runtime calls the plain-C host files need, behind C wrappers
 * (keeps cuda_runtime.h out of the C sources; [ref: src/gpujpeg_util.h:48-62] for the error style).
 */
#include <cuda_runtime.h>
#include <stdio.h>
#include <string.h>

#include "gj_internal.h"

static const char* g_last = "";
static int chk(cudaError_t e)
{
    if ( e == cudaSuccess ) return 0;
    g_last = cudaGetErrorString(e);
    return -1;
}

extern "C" {

int gj_cuda_malloc(void** p, size_t size) { return chk(cudaMalloc(p, size)); }
int gj_cuda_free(void* p) { return p ? chk(cudaFree(p)) : 0; }
int gj_cuda_malloc_host(void** p, size_t size) { return chk(cudaMallocHost(p, size)); }
int gj_cuda_free_host(void* p) { return p ? chk(cudaFreeHost(p)) : 0; }
int gj_cuda_memcpy_h2d_async(void* dst, const void* src, size_t size, gj_stream_t s)
{
    return chk(cudaMemcpyAsync(dst, src, size, cudaMemcpyHostToDevice, s));
}
int gj_cuda_memcpy_d2h_async(void* dst, const void* src, size_t size, gj_stream_t s)
{
    return chk(cudaMemcpyAsync(dst, src, size, cudaMemcpyDeviceToHost, s));
}
int gj_cuda_memcpy_d2d_async(void* dst, const void* src, size_t size, gj_stream_t s)
{
    return chk(cudaMemcpyAsync(dst, src, size, cudaMemcpyDeviceToDevice, s));
}
int gj_cuda_memset_async(void* dst, int v, size_t size, gj_stream_t s) { return chk(cudaMemsetAsync(dst, v, size, s)); }
int gj_cuda_stream_sync(gj_stream_t s) { return chk(cudaStreamSynchronize(s)); }
int gj_cuda_pointer_is_device(const void* p)
{
    cudaPointerAttributes a;
    if ( cudaPointerGetAttributes(&a, p) != cudaSuccess ) {
        cudaGetLastError();
        return 0;
    }
    return a.type == cudaMemoryTypeDevice || a.type == cudaMemoryTypeManaged;
}
const char* gj_cuda_last_error(void)
{
    const cudaError_t e = cudaGetLastError();
    if ( e != cudaSuccess ) g_last = cudaGetErrorString(e);
    return g_last;
}

int gj_timer_create(struct gj_timer* t)
{
    t->armed = 0;
    if ( chk(cudaEventCreate((cudaEvent_t*)&t->start)) ) return -1;
    return chk(cudaEventCreate((cudaEvent_t*)&t->stop));
}
void gj_timer_destroy(struct gj_timer* t)
{
    if ( t->start ) cudaEventDestroy((cudaEvent_t)t->start);
    if ( t->stop ) cudaEventDestroy((cudaEvent_t)t->stop);
    t->start = t->stop = NULL;
}
void gj_timer_start(struct gj_timer* t, gj_stream_t s)
{
    if ( t->start ) cudaEventRecord((cudaEvent_t)t->start, s);
    t->armed = 1;
}
void gj_timer_stop(struct gj_timer* t, gj_stream_t s)
{
    if ( t->stop ) cudaEventRecord((cudaEvent_t)t->stop, s);
    t->armed = 2;
}
double gj_timer_ms(struct gj_timer* t)
{
    if ( t->armed != 2 ) return 0.0;
    float ms = 0.f;
    if ( cudaEventSynchronize((cudaEvent_t)t->stop) != cudaSuccess ) return 0.0;
    if ( cudaEventElapsedTime(&ms, (cudaEvent_t)t->start, (cudaEvent_t)t->stop) != cudaSuccess ) {
        cudaGetLastError();
        return 0.0;
    }
    return ms;
}

int gj_cuda_device_count(void)
{
    int n = 0;
    if ( cudaGetDeviceCount(&n) != cudaSuccess ) {
        g_last = cudaGetErrorString(cudaGetLastError());
        return -1;
    }
    return n;
}
int gj_cuda_device_props(int dev, struct gpujpeg_device_info* info)
{
    cudaDeviceProp p;
    if ( chk(cudaGetDeviceProperties(&p, dev)) ) return -1;
    memset(info, 0, sizeof *info);
    info->id = dev;
    strncpy(info->name, p.name, sizeof info->name - 1);
    info->cc_major = p.major;
    info->cc_minor = p.minor;
    info->global_memory = p.totalGlobalMem;
    info->constant_memory = p.totalConstMem;
    info->shared_memory = p.sharedMemPerBlock;
    info->register_count = p.regsPerBlock;
    info->multiprocessor_count = p.multiProcessorCount;
    return 0;
}
int gj_cuda_set_device(int dev) { return chk(cudaSetDevice(dev)); }
int gj_cuda_get_device(void)
{
    int d = -1;
    if ( cudaGetDevice(&d) != cudaSuccess ) {
        cudaGetLastError();
        return -1;
    }
    return d;
}
void gj_cuda_device_reset(void) { cudaDeviceReset(); }
}

extern "C" int gj_cuda_sm_count(void)
{
    int dev = 0, n = 0;
    if ( cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n < 1 )
        return 1;
    return n;
}

extern "C" int gj_cuda_stream_create(gj_stream_t* s)
{
    cudaStream_t st;
    if ( cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking) != cudaSuccess ) return -1;
    *s = st;
    return 0;
}
extern "C" void gj_cuda_stream_destroy(gj_stream_t s) { cudaStreamDestroy(s); }
extern "C" int gj_cuda_event_create(void** ev)
{
    cudaEvent_t e;
    if ( cudaEventCreateWithFlags(&e, cudaEventDisableTiming) != cudaSuccess ) return -1;
    *ev = e;
    return 0;
}
extern "C" void gj_cuda_event_destroy(void* ev) { if ( ev ) cudaEventDestroy((cudaEvent_t)ev); }
extern "C" int gj_cuda_event_record(void* ev, gj_stream_t s) { return cudaEventRecord((cudaEvent_t)ev, s) == cudaSuccess ? 0 : -1; }
extern "C" int gj_cuda_stream_wait_event(gj_stream_t s, void* ev) { return cudaStreamWaitEvent(s, (cudaEvent_t)ev, 0) == cudaSuccess ? 0 : -1; }
extern "C" int gj_cuda_enable_peer(int peer)
{
    const cudaError_t e = cudaDeviceEnablePeerAccess(peer, 0);
    if ( e == cudaSuccess || e == cudaErrorPeerAccessAlreadyEnabled ) {
        cudaGetLastError();
        return 0;
    }
    cudaGetLastError();
    return -1;
}
extern "C" int gj_cuda_memcpy_peer_async(void* dst, int dst_dev, const void* src, int src_dev, size_t size, gj_stream_t s)
{
    return cudaMemcpyPeerAsync(dst, dst_dev, src, src_dev, size, s) == cudaSuccess ? 0 : -1;
}
