/*
 * gj_batch.c -- gpujpegx_batch_*: a batch of independent frames sharded over the GPUs of one box.  Host C, pthreads.
 *
 * One persistent worker per device: host thread + CUDA stream + encoder + decoder, exactly the reference's concurrency
 * contract (one coder per thread and stream, instances share nothing: test/misc/mt_encode.c:12-43;
 * gpujpeg_init_device, src/gpujpeg_common.c:219-288).  Frame f belongs to worker f mod N.  The
 * data path has no collective: the only traffic between devices is the optional scatter of raw frames from / gather of
 * decoded frames to the first device (GPUJPEGX_DEVICE_FIRST), done with peer copies on the owner's stream so that the
 * copy of frame f + N overlaps nothing it depends on.
 */
#include <pthread.h>
#include <stdlib.h>
#include <string.h>

#include "../../include/gpujpegx.h"
#include "gj_internal.h"

enum { JOB_NONE = 0, JOB_ENCODE = 1, JOB_DECODE = 2, JOB_EXIT = 3 };

struct worker {
    struct gpujpegx_batch* batch;
    int index, device;
    pthread_t thread;
    int started;
    gj_stream_t stream;
    struct gpujpeg_encoder* enc;
    struct gpujpeg_decoder* dec;
    uint8_t* d_stage; size_t d_stage_size;   /* frames moved from / to the first device */
    int init_error;
};

struct gpujpegx_batch {
    int n;
    struct worker* w;
    pthread_mutex_t mu;
    pthread_cond_t cv_job, cv_done;
    unsigned long long generation;   /* bumped for every job */
    int job, pending, failed;
    /* the current job */
    const struct gpujpeg_parameters* param;
    const struct gpujpeg_image_parameters* param_image;
    const uint8_t* const* in;
    const size_t* in_sizes;
    uint8_t* const* out;
    int count;
    enum gpujpegx_location where;
    /* encoder results, batch-owned */
    uint8_t** jpeg; size_t* jpeg_cap; size_t* jpeg_size; int jpeg_slots;
    double last_ms;
};

static int grow_stage(struct worker* w, size_t size)
{
    if ( w->d_stage_size >= size ) return 0;
    gj_cuda_free(w->d_stage);
    w->d_stage = NULL;
    w->d_stage_size = 0;
    if ( gj_cuda_malloc((void**)&w->d_stage, size) ) return -1;
    w->d_stage_size = size;
    return 0;
}

static int encode_frame(struct worker* w, int f)
{
    struct gpujpegx_batch* b = w->batch;
    struct gpujpeg_encoder_input in;
    const uint8_t* src = b->in[f];
    struct gpujpeg_image_parameters pi = *b->param_image;
    const size_t raw = gpujpeg_image_calculate_size(&pi);
    if ( b->where == GPUJPEGX_HOST ) {
        gpujpeg_encoder_input_set_image(&in, (uint8_t*)src);
    }
    else if ( b->where == GPUJPEGX_DEVICE_OWNER || w->index == 0 ) {
        gpujpeg_encoder_input_set_gpu_image(&in, (uint8_t*)src);
    }
    else {   /* the frame lives on the first device: fetch it over the peer link, on this worker's stream */
        if ( grow_stage(w, raw) || gj_cuda_memcpy_peer_async(w->d_stage, w->device, src, b->w[0].device, raw, w->stream) ) return -1;
        gpujpeg_encoder_input_set_gpu_image(&in, w->d_stage);
    }
    uint8_t* jpeg = NULL;
    size_t size = 0;
    if ( gpujpeg_encoder_encode(w->enc, b->param, b->param_image, &in, &jpeg, &size) ) return -1;
    if ( b->jpeg_cap[f] < size ) {
        free(b->jpeg[f]);
        b->jpeg[f] = (uint8_t*)malloc(size + size / 8 + 4096);
        b->jpeg_cap[f] = b->jpeg[f] ? size + size / 8 + 4096 : 0;
        if ( !b->jpeg[f] ) return -1;
    }
    memcpy(b->jpeg[f], jpeg, size);   /* the encoder's buffer is reused by its next frame */
    b->jpeg_size[f] = size;
    return 0;
}

static int decode_frame(struct worker* w, int f)
{
    struct gpujpegx_batch* b = w->batch;
    struct gpujpeg_decoder_output out;
    if ( b->where == GPUJPEGX_HOST ) {
        gpujpeg_decoder_output_set_custom(&out, b->out[f]);
        return gpujpeg_decoder_decode(w->dec, (uint8_t*)b->in[f], b->in_sizes[f], &out) ? -1 : 0;
    }
    if ( b->where == GPUJPEGX_DEVICE_OWNER || w->index == 0 ) {
        gpujpeg_decoder_output_set_custom_cuda(&out, b->out[f]);
        return gpujpeg_decoder_decode(w->dec, (uint8_t*)b->in[f], b->in_sizes[f], &out) ? -1 : 0;
    }
    /* decode here, then hand the pixels to the first device over the peer link */
    gpujpeg_decoder_output_set_cuda_buffer(&out);
    if ( gpujpeg_decoder_decode(w->dec, (uint8_t*)b->in[f], b->in_sizes[f], &out) ) return -1;
    if ( gj_cuda_memcpy_peer_async(b->out[f], b->w[0].device, out.data, w->device, out.data_size, w->stream) ||
         gj_cuda_stream_sync(w->stream) )
        return -1;
    return 0;
}

static void* worker_main(void* arg)
{
    struct worker* w = (struct worker*)arg;
    struct gpujpegx_batch* b = w->batch;
    /* [ref: src/gpujpeg_common.c:219-288 gpujpeg_init_device] the device is a property of the calling thread */
    if ( gj_cuda_set_device(w->device) || gj_cuda_stream_create(&w->stream) ) w->init_error = 1;
    if ( !w->init_error ) {
        for ( int k = 0; k < b->n; k++ )
            if ( k != w->index ) gj_cuda_enable_peer(b->w[k].device);   /* best effort: without it peer copies are staged by the driver */
        w->enc = gpujpeg_encoder_create((cudaStream_t)w->stream);
        w->dec = gpujpeg_decoder_create((cudaStream_t)w->stream);
        if ( !w->enc || !w->dec ) w->init_error = 1;
        else gpujpeg_encoder_set_option(w->enc, GPUJPEG_ENC_OPT_OUT, GPUJPEG_ENC_OUT_VAL_PINNED);
    }
    unsigned long long seen = 0;
    pthread_mutex_lock(&b->mu);
    b->pending--;   /* initialisation done */
    pthread_cond_broadcast(&b->cv_done);
    for ( ;; ) {
        while ( b->generation == seen )
            pthread_cond_wait(&b->cv_job, &b->mu);
        seen = b->generation;
        const int job = b->job;
        pthread_mutex_unlock(&b->mu);
        int failed = 0;
        if ( job == JOB_ENCODE || job == JOB_DECODE ) {
            for ( int f = w->index; f < b->count && !failed; f += b->n )
                failed = w->init_error || (job == JOB_ENCODE ? encode_frame(w, f) : decode_frame(w, f));
        }
        pthread_mutex_lock(&b->mu);
        if ( failed ) b->failed = 1;
        b->pending--;
        pthread_cond_broadcast(&b->cv_done);
        if ( job == JOB_EXIT ) break;
    }
    pthread_mutex_unlock(&b->mu);
    if ( w->enc ) gpujpeg_encoder_destroy(w->enc);
    if ( w->dec ) gpujpeg_decoder_destroy(w->dec);
    gj_cuda_free(w->d_stage);
    if ( w->stream ) gj_cuda_stream_destroy(w->stream);
    return NULL;
}

/* hands a job to every worker and waits for all of them */
static int run_job(struct gpujpegx_batch* b, int job)
{
    const double t0 = gpujpeg_get_time();
    pthread_mutex_lock(&b->mu);
    b->job = job;
    b->failed = 0;
    b->pending = b->n;
    b->generation++;
    pthread_cond_broadcast(&b->cv_job);
    while ( b->pending > 0 )
        pthread_cond_wait(&b->cv_done, &b->mu);
    const int failed = b->failed;
    pthread_mutex_unlock(&b->mu);
    b->last_ms = (gpujpeg_get_time() - t0) * 1000.0;
    return failed ? -1 : 0;
}

struct gpujpegx_batch* gpujpegx_batch_create(const int* devices, int device_count)
{
    const int visible = gj_cuda_device_count();
    if ( visible < 1 ) {
        GJ_ERR("No CUDA device for the batch coder.\n");
        return NULL;
    }
    if ( !devices || device_count < 1 ) device_count = visible;
    struct gpujpegx_batch* b = (struct gpujpegx_batch*)calloc(1, sizeof *b);
    if ( !b ) return NULL;
    b->n = device_count;
    b->w = (struct worker*)calloc((size_t)device_count, sizeof *b->w);
    pthread_mutex_init(&b->mu, NULL);
    pthread_cond_init(&b->cv_job, NULL);
    pthread_cond_init(&b->cv_done, NULL);
    if ( !b->w ) {
        gpujpegx_batch_destroy(b);
        return NULL;
    }
    b->pending = device_count;
    for ( int i = 0; i < device_count; i++ ) {
        b->w[i].batch = b;
        b->w[i].index = i;
        b->w[i].device = devices ? devices[i] : i;
        if ( b->w[i].device < 0 || b->w[i].device >= visible || pthread_create(&b->w[i].thread, NULL, worker_main, &b->w[i]) ) {
            pthread_mutex_lock(&b->mu);
            b->pending--;
            b->w[i].init_error = 1;
            pthread_mutex_unlock(&b->mu);
            continue;
        }
        b->w[i].started = 1;
    }
    pthread_mutex_lock(&b->mu);
    while ( b->pending > 0 )
        pthread_cond_wait(&b->cv_done, &b->mu);
    pthread_mutex_unlock(&b->mu);
    for ( int i = 0; i < device_count; i++ ) {
        if ( b->w[i].init_error ) {
            GJ_ERR("Batch coder: device %d cannot be initialised: %s\n", b->w[i].device, gj_cuda_last_error());
            gpujpegx_batch_destroy(b);
            return NULL;
        }
    }
    return b;
}

void gpujpegx_batch_destroy(struct gpujpegx_batch* b)
{
    if ( !b ) return;
    if ( b->w ) {
        int started = 0;
        for ( int i = 0; i < b->n; i++ )
            started += b->w[i].started;
        if ( started ) {
            pthread_mutex_lock(&b->mu);
            b->job = JOB_EXIT;
            b->pending = started;
            b->generation++;
            pthread_cond_broadcast(&b->cv_job);
            pthread_mutex_unlock(&b->mu);
            for ( int i = 0; i < b->n; i++ )
                if ( b->w[i].started ) pthread_join(b->w[i].thread, NULL);
        }
        free(b->w);
    }
    for ( int f = 0; f < b->jpeg_slots; f++ )
        free(b->jpeg[f]);
    free(b->jpeg);
    free(b->jpeg_cap);
    free(b->jpeg_size);
    pthread_mutex_destroy(&b->mu);
    pthread_cond_destroy(&b->cv_job);
    pthread_cond_destroy(&b->cv_done);
    free(b);
}

int gpujpegx_batch_device_count(const struct gpujpegx_batch* b) { return b ? b->n : 0; }
int gpujpegx_batch_owner(const struct gpujpegx_batch* b, int frame) { return b && frame >= 0 ? b->w[frame % b->n].device : -1; }
double gpujpegx_batch_last_ms(const struct gpujpegx_batch* b) { return b ? b->last_ms : 0.0; }

int gpujpegx_batch_encode(struct gpujpegx_batch* b, const struct gpujpeg_parameters* param,
                          const struct gpujpeg_image_parameters* param_image, const uint8_t* const* images, int count,
                          enum gpujpegx_location where, uint8_t** jpegs, size_t* sizes)
{
    if ( !b || !param || !param_image || !images || count < 0 || !jpegs || !sizes ) return -1;
    if ( count > b->jpeg_slots ) {
        uint8_t** j = (uint8_t**)realloc(b->jpeg, (size_t)count * sizeof *j);
        if ( j ) b->jpeg = j;
        size_t* c = (size_t*)realloc(b->jpeg_cap, (size_t)count * sizeof *c);
        if ( c ) b->jpeg_cap = c;
        size_t* s = (size_t*)realloc(b->jpeg_size, (size_t)count * sizeof *s);
        if ( s ) b->jpeg_size = s;
        if ( !j || !c || !s ) return -1;
        for ( int f = b->jpeg_slots; f < count; f++ ) {
            b->jpeg[f] = NULL;
            b->jpeg_cap[f] = b->jpeg_size[f] = 0;
        }
        b->jpeg_slots = count;
    }
    b->param = param;
    b->param_image = param_image;
    b->in = images;
    b->count = count;
    b->where = where;
    const int rc = run_job(b, JOB_ENCODE);
    for ( int f = 0; f < count; f++ ) {
        jpegs[f] = rc ? NULL : b->jpeg[f];
        sizes[f] = rc ? 0 : b->jpeg_size[f];
    }
    return rc;
}

int gpujpegx_batch_decode(struct gpujpegx_batch* b, const uint8_t* const* jpegs, const size_t* sizes, int count,
                          uint8_t* const* outputs, enum gpujpegx_location where)
{
    if ( !b || !jpegs || !sizes || count < 0 || !outputs ) return -1;
    b->in = jpegs;
    b->in_sizes = sizes;
    b->out = outputs;
    b->count = count;
    b->where = where;
    return run_job(b, JOB_DECODE);
}
