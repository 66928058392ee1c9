// cornetto_b200/csrc/sdust_wide.cu -- symmetric DUST for window sizes the hot kernel does not cover.
//
// `sdust -w W` takes any W in the reference (src/sdust/sdust.c:186-189).  csrc/sdust.cu is built around the
// default W = 64 (byte counters, two or four 32-position window blocks, state in shared memory) and serves
// W <= 128.  This file is the generic instance for 128 < W <= 1024: the same chunked execution (exact mid-record
// start, save events owned by time, seam fold -- sdust_core.cuh), one thread per chunk, every routine in its
// serial form, 16-bit counters and 64-bit slots, per-thread state in global memory (it stays in L1/L2).  It is
// not tuned: nobody runs sdust with such windows on a whole genome, and the reference itself needs
// O(W) to O(W^2) work per base inside low-complexity sequence there.
#include "corn_internal.cuh"
#define SD_WIDE
#include "sdust_core.cuh"

using namespace sd_wide;

namespace {

struct WideFetch {
    const uint8_t *seq;
    __device__ __forceinline__ uint8_t operator()(int i) const { return __ldg(seq + i); }
};

__global__ void __launch_bounds__(128) k_sdust_scan_wide(const corn_sdust_wide_params P)
{
    const uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= P.n_chunks) return;
    const uint32_t rec = corn_upper_bound(P.chunk_base, P.n_rec, j) - 1;
    const uint32_t k = j - P.chunk_base[rec];
    const int len = (int)P.rec_len[rec];
    const int c0 = (int)k * P.C, c1 = min(len, c0 + P.C);
    // per-chunk state row: slots (8-byte aligned) | cw | cv | ring
    uint8_t *row = P.state + (size_t)j * P.state_stride;
    sd_mem m;
    m.slot = (sd_slot_t *)row;
    m.cw = (sd_cnt_t *)(row + (size_t)P.W * sizeof(sd_slot_t));
    m.cv = m.cw + 64;
    m.ring = (uint8_t *)(m.cv + 64);
    m.pitch = 0;
    sd_sink sink;
    sd_sink_init(sink, P.slots + (size_t)j * P.cap, P.cap);
    WideFetch fetch;
    fetch.seq = P.seq + P.rec_off[rec];
    sd_run_chunk(fetch, len, c0, c1, P.T, P.W, m, sink);
    P.cnt[j] = sink.n;
    if (sink.overflow) atomicAdd(P.err, 1u);
}

}  // namespace

size_t corn_sdust_wide_state_stride(int W)
{
    return (((size_t)W * sizeof(sd_slot_t) + 128 * sizeof(sd_cnt_t) + (size_t)W + 15) / 16) * 16;
}

int corn_sdust_wide_scan(corn_ctx *ctx, const corn_sdust_wide_params &P)
{
    if (P.W <= 128 || P.W > SD_MAX_W) return corn_set_err(ctx, CORN_E_ARG, "wide sdust instance: W = %d outside (128, %d]", P.W, SD_MAX_W);
    k_sdust_scan_wide<<<(P.n_chunks + 127) / 128, 128, 0, ctx->stream>>>(P);
    corn_count_launch(ctx);
    CORN_LAUNCH_CHECK(ctx);
    return CORN_OK;
}
