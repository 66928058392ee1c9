// cornetto_b200/csrc/sdust_api.cu -- the host-callable sdust library interface of the reference,
// src/sdust/sdust.h:16-21, on top of corn_gpu_sdust():
//     uint64_t *sdust(void *km, const uint8_t *seq, int l_seq, int T, int W, int *n);         (:16)
//     sdust_buf_t *sdust_buf_init(void *km);  void sdust_buf_destroy(sdust_buf_t *buf);          (:19-20)
//     const uint64_t *sdust_core(const uint8_t *seq, int l_seq, int T, int W, int *n, sdust_buf_t *buf);   (:21)
// Same argument meaning and ownership rules as src/sdust/sdust.c:130-171: l_seq < 0 means strlen(seq); every call
// starts from a fresh state (:134-137); sdust_core()'s result belongs to the buffer and lives until the next call
// on it or sdust_buf_destroy(); sdust()'s result is malloc()ed and the caller free()s it (the reference's kalloc.h
// maps kmalloc/kfree onto malloc/free and ignores `km`; so does this).  Intervals are start<<32 | finish (:99).
// One sequence per call is one record per batch: fine for the occasional caller, while bulk users should batch
// records through corn_gpu_sdust() (what `cornetto sdust` does).
//
// The reference cannot fail; this can (no device, out of memory): then NULL is returned and *n = -1.
#include <mutex>

#include "corn_internal.cuh"

struct sdust_buf_s {
    void          *km;
    corn_hbatch_t *hb;
    uint64_t       hb_cap;
    uint64_t      *res;       // result of the last sdust_core() on this buffer
    size_t         res_cap;
};

namespace {
std::mutex  g_mu;             // one shared context: calls from several threads are serialised
corn_ctx_t *g_ctx;

corn_ctx_t *shared_ctx()
{
    if (!g_ctx && corn_gpu_init(-1, &g_ctx) != CORN_OK) g_ctx = NULL;
    return g_ctx;
}
}  // namespace

extern "C" sdust_buf_t *sdust_buf_init(void *km)
{
    sdust_buf_t *buf = (sdust_buf_t *)calloc(1, sizeof(sdust_buf_t));
    if (buf) buf->km = km;
    return buf;
}

extern "C" void sdust_buf_destroy(sdust_buf_t *buf)
{
    if (!buf) return;
    if (buf->hb) corn_hbatch_destroy(buf->hb);
    free(buf->res);
    free(buf);
}

extern "C" const uint64_t *sdust_core(const uint8_t *seq, int l_seq, int T, int W, int *n, sdust_buf_t *buf)
{
    if (n) *n = -1;
    if (!seq || !buf || !n) return NULL;
    if (l_seq < 0) l_seq = (int)strlen((const char *)seq);
    std::lock_guard<std::mutex> lock(g_mu);
    corn_ctx_t *ctx = shared_ctx();
    if (!ctx) return NULL;
    const uint64_t need = (uint64_t)l_seq + 4 * CORN_ALIGN;
    if (!buf->hb || buf->hb_cap < need) {
        if (buf->hb) corn_hbatch_destroy(buf->hb);
        buf->hb = NULL;
        buf->hb_cap = need < (1u << 16) ? (1u << 16) : need + need / 4;
        if (corn_hbatch_create(buf->hb_cap, 1, &buf->hb) != CORN_OK) { buf->hb = NULL; buf->hb_cap = 0; return NULL; }
    }
    corn_hbatch_reset(buf->hb);
    if (corn_hbatch_add(buf->hb, seq, (uint64_t)l_seq) != CORN_OK) return NULL;
    corn_batch_t view;
    corn_hbatch_view(buf->hb, &view);
    corn_intervals_t iv;
    if (corn_gpu_sdust(ctx, &view, T, W, &iv) != CORN_OK) return NULL;
    if (iv.n_iv > 0x7FFFFFFFull) { corn_gpu_intervals_free(&iv); return NULL; }
    if (iv.n_iv + 1 > buf->res_cap) {
        free(buf->res);
        buf->res_cap = (size_t)iv.n_iv + 1 + (size_t)iv.n_iv / 2;
        buf->res = (uint64_t *)malloc(buf->res_cap * sizeof(uint64_t));
        if (!buf->res) { buf->res_cap = 0; corn_gpu_intervals_free(&iv); return NULL; }
    }
    if (iv.n_iv) memcpy(buf->res, iv.iv, (size_t)iv.n_iv * sizeof(uint64_t));
    *n = (int)iv.n_iv;
    corn_gpu_intervals_free(&iv);
    return buf->res;
}

extern "C" uint64_t *sdust(void *km, const uint8_t *seq, int l_seq, int T, int W, int *n)
{
    sdust_buf_t *buf = sdust_buf_init(km);
    if (!buf) { if (n) *n = -1; return NULL; }
    uint64_t *ret = (uint64_t *)sdust_core(seq, l_seq, T, W, n, buf);
    if (ret) buf->res = NULL;          // handed to the caller (src/sdust/sdust.c:168: buf->res.a = 0)
    sdust_buf_destroy(buf);
    return ret;
}
