/* cornetto_b200/host/pipeline.c -- parse / GPU / print pipeline shared by telofind and sdust.
 *
 * The reference handles one record at a time on one thread (kseq_read -> scan -> printf).  Here
 * the calling thread only parses: it fills pinned batches, each owned by a worker thread that has
 * its own GPU context, and a worker runs the C-ABI call and formats the text of its batch while
 * the parser is already filling the next one.  Workers are visited round-robin and a worker's
 * previous output is written before its batch buffer is refilled, so stdout is in batch order.
 * With $CORNETTO_GPUS=N the workers use N devices (independent shards, no device-to-device
 * traffic); with one device two workers still overlap parsing with H2D + kernels + formatting. */
#include <pthread.h>

#include "cornetto.h"

typedef struct {
    pthread_t       th;
    pthread_mutex_t mu;
    pthread_cond_t  cv;
    int             state;      /* 0 idle (output, if any, ready), 1 has work, 2 quit */
    int             device;
    corn_ctx_t     *ctx;
    rec_batch_t    *batch;
    outbuf_t        out;
    batch_fn        fn;
    void           *arg;
} worker_t;

static void *worker_main(void *p)
{
    worker_t *w = (worker_t *)p;
    int r = corn_gpu_init(w->device, &w->ctx);
    if (r != CORN_OK) {
        CORN_ERROR("cannot initialise the GPU (device %d): %s", w->device, corn_gpu_strerror(r));
        exit(EXIT_FAILURE);
    }
    for (;;) {
        pthread_mutex_lock(&w->mu);
        while (w->state == 0) pthread_cond_wait(&w->cv, &w->mu);
        const int st = w->state;
        pthread_mutex_unlock(&w->mu);
        if (st == 2) break;
        /* (batch buffers stay pageable: corn_gpu_upload stages the copy through a small page-locked ring) */
        w->fn(w->ctx, w->batch, &w->out, w->arg);
        pthread_mutex_lock(&w->mu);
        w->state = 0;
        pthread_cond_broadcast(&w->cv);
        pthread_mutex_unlock(&w->mu);
    }
    if (!cornetto_fast_exit()) corn_gpu_destroy(w->ctx);
    return NULL;
}

static void wait_idle(worker_t *w)
{
    pthread_mutex_lock(&w->mu);
    while (w->state != 0) pthread_cond_wait(&w->cv, &w->mu);
    pthread_mutex_unlock(&w->mu);
}

void run_batch_pipeline(fastx_t *fx, const char *path, batch_fn fn, void *arg)
{
    int n_gpus = 1;
    const char *e = getenv("CORNETTO_GPUS");
    if (e && atoi(e) > 0) n_gpus = atoi(e);
    if (n_gpus > 1) {                                   /* (the single-GPU path leaves all CUDA start-up to the workers, */
        const int avail = corn_gpu_device_count();      /*  so that parsing overlaps the seconds the driver needs)       */
        if (avail <= 0) {
            CORN_ERROR("cannot initialise the GPU: %s", corn_gpu_strerror(avail < 0 ? avail : CORN_E_NOGPU));
            exit(EXIT_FAILURE);
        }
        if (n_gpus > avail) n_gpus = avail;
    }
    const int n_workers = n_gpus > 1 ? n_gpus : 2;     /* one device: two workers double-buffer */
    const char *dev0 = getenv("CORNETTO_GPU");
    const int base_dev = (n_gpus == 1 && dev0) ? atoi(dev0) : 0;

    const uint64_t cap = cornetto_batch_capacity(path, n_gpus);
    uint64_t max_rec = cap / 64 + 16;
    if (max_rec > (1u << 23)) max_rec = 1u << 23;

    worker_t *w = (worker_t *)calloc((size_t)n_workers, sizeof(worker_t));
    CORN_MALLOC_CHK(w);
    for (int i = 0; i < n_workers; ++i) {
        pthread_mutex_init(&w[i].mu, NULL);
        pthread_cond_init(&w[i].cv, NULL);
        w[i].device = base_dev + (i % n_gpus);
        w[i].fn = fn; w[i].arg = arg;
        w[i].batch = NULL;                              /* allocated on first use: small inputs touch one worker only */
        outbuf_init(&w[i].out, NULL);
        if (pthread_create(&w[i].th, NULL, worker_main, &w[i]) != 0) { CORN_ERROR("%s", "pthread_create failed"); exit(EXIT_FAILURE); }
    }
    int dispatched = 0, written = 0;                    /* batches handed to workers / outputs written, in order */
    for (;;) {
        worker_t *x = &w[dispatched % n_workers];
        wait_idle(x);
        if (dispatched - n_workers >= written) { outbuf_write(&x->out, cornetto_pipeline_out()); written = dispatched - n_workers + 1; }
        if (!x->batch) x->batch = rec_batch_create(cap, (uint32_t)max_rec);
        if (rec_batch_fill(x->batch, fx) == 0) break;
        const int input_done = x->batch->eof;
        x->batch->seq = cornetto_next_batch_seq();
        pthread_mutex_lock(&x->mu);
        x->state = 1;
        pthread_cond_broadcast(&x->cv);
        pthread_mutex_unlock(&x->mu);
        ++dispatched;
        if (input_done) break;
    }
    for (; written < dispatched; ++written) {           /* drain, oldest first */
        worker_t *y = &w[written % n_workers];
        wait_idle(y);
        outbuf_write(&y->out, cornetto_pipeline_out());
    }
    for (int i = 0; i < n_workers; ++i) {
        pthread_mutex_lock(&w[i].mu);
        w[i].state = 2;
        pthread_cond_broadcast(&w[i].cv);
        pthread_mutex_unlock(&w[i].mu);
        pthread_join(w[i].th, NULL);
        outbuf_free(&w[i].out);
        if (w[i].batch && !cornetto_fast_exit()) rec_batch_destroy(w[i].batch);
    }
    fflush(cornetto_pipeline_out());
    free(w);
}
