// cornetto_b200/csrc/scan.cu -- exclusive prefix sums over the small tables of the sparse phase
// (per-tile event counts, per-record chunk counts, per-block window counts).  These tables hold
// 10^4..10^6 entries, so a plain reduce / scan-of-sums / scan three-kernel scheme is enough.
#include "corn_internal.cuh"

namespace {

constexpr int SCAN_THREADS = 512;
constexpr int SCAN_ITEMS   = 8;
constexpr int SCAN_BLOCK   = SCAN_THREADS * SCAN_ITEMS;

__device__ __forceinline__ uint32_t zero_of(uint32_t) { return 0u; }
__device__ __forceinline__ uint4    zero_of(uint4)    { return make_uint4(0, 0, 0, 0); }
__device__ __forceinline__ uint32_t add(uint32_t a, uint32_t b) { return a + b; }
__device__ __forceinline__ uint4    add(uint4 a, uint4 b) { return make_uint4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w); }

// exclusive scan of one value per thread across the block; returns the block total in *total
template <typename T>
__device__ T block_exclusive(T v, T *smem, T *total)
{
    const int tid = threadIdx.x;
    smem[tid] = v;
    __syncthreads();
    for (int o = 1; o < SCAN_THREADS; o <<= 1) {
        T t = zero_of(T());
        if (tid >= o) t = smem[tid - o];
        __syncthreads();
        if (tid >= o) smem[tid] = add(smem[tid], t);
        __syncthreads();
    }
    T incl = smem[tid];
    *total = smem[SCAN_THREADS - 1];
    T excl = tid ? smem[tid - 1] : zero_of(T());
    __syncthreads();
    (void)incl;
    return excl;
}

template <typename T>
__global__ void __launch_bounds__(SCAN_THREADS) k_block_sums(const T *__restrict__ in, T *__restrict__ sums, size_t n)
{
    __shared__ T smem[SCAN_THREADS];
    size_t base = (size_t)blockIdx.x * SCAN_BLOCK + (size_t)threadIdx.x * SCAN_ITEMS;
    T s = zero_of(T());
#pragma unroll
    for (int i = 0; i < SCAN_ITEMS; ++i)
        if (base + i < n) s = add(s, in[base + i]);
    T total;
    block_exclusive(s, smem, &total);
    if (threadIdx.x == 0) sums[blockIdx.x] = total;
}

// out[i] = offset[block] + exclusive prefix inside the block.  offsets == NULL means 0.
template <typename T>
__global__ void __launch_bounds__(SCAN_THREADS) k_block_scan(const T *in, T *out, size_t n,   // in == out allowed
                                                             const T *offsets, T *total_out)
{
    __shared__ T smem[SCAN_THREADS];
    size_t base = (size_t)blockIdx.x * SCAN_BLOCK + (size_t)threadIdx.x * SCAN_ITEMS;
    T v[SCAN_ITEMS];
    T s = zero_of(T());
#pragma unroll
    for (int i = 0; i < SCAN_ITEMS; ++i) {
        v[i] = (base + i < n) ? in[base + i] : zero_of(T());
        s = add(s, v[i]);
    }
    T total;
    T excl = block_exclusive(s, smem, &total);
    T off = offsets ? offsets[blockIdx.x] : zero_of(T());
    T run = add(off, excl);
#pragma unroll
    for (int i = 0; i < SCAN_ITEMS; ++i) {
        if (base + i < n) out[base + i] = run;
        run = add(run, v[i]);
    }
    if (total_out && blockIdx.x == gridDim.x - 1 && threadIdx.x == 0) *total_out = add(off, total);
}

template <typename T>
__global__ void k_zero_total(T *t) { *t = zero_of(T()); }

// Single-launch scan for tables of up to ~1M entries (the per-tile counts of a 4 GiB batch are 131 072):
// one block of 32 warps; warp w owns a contiguous slice and walks it 32 entries at a time (coalesced),
// first to get its total, then -- after the 32 totals have been scanned -- to write the prefixes.
// The table is read twice from L2; what is saved is two launches and their drain/fill latency.
template <typename T>
__device__ __forceinline__ T warp_inclusive(T v, int lane);
template <>
__device__ __forceinline__ uint32_t warp_inclusive<uint32_t>(uint32_t v, int lane)
{
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const uint32_t t = __shfl_up_sync(0xffffffffu, v, o); if (lane >= o) v += t; }
    return v;
}
template <>
__device__ __forceinline__ uint4 warp_inclusive<uint4>(uint4 v, int lane)
{
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t a = __shfl_up_sync(0xffffffffu, v.x, o), b = __shfl_up_sync(0xffffffffu, v.y, o);
        const uint32_t c = __shfl_up_sync(0xffffffffu, v.z, o), d = __shfl_up_sync(0xffffffffu, v.w, o);
        if (lane >= o) { v.x += a; v.y += b; v.z += c; v.w += d; }
    }
    return v;
}
__device__ __forceinline__ uint32_t bcast31(uint32_t v) { return __shfl_sync(0xffffffffu, v, 31); }
__device__ __forceinline__ uint4 bcast31(uint4 v)
{
    return make_uint4(__shfl_sync(0xffffffffu, v.x, 31), __shfl_sync(0xffffffffu, v.y, 31), __shfl_sync(0xffffffffu, v.z, 31), __shfl_sync(0xffffffffu, v.w, 31));
}
__device__ __forceinline__ uint32_t sub(uint32_t a, uint32_t b) { return a - b; }
__device__ __forceinline__ uint4 sub(uint4 a, uint4 b) { return make_uint4(a.x - b.x, a.y - b.y, a.z - b.z, a.w - b.w); }

template <typename T>
__global__ void __launch_bounds__(1024) k_scan_single(const T *in, T *out, size_t n, T *total_out)
{
    __shared__ T wtot[32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const size_t per = ((n + 31) / 32 + 31) / 32 * 32;           // slice length, multiple of 32
    const size_t lo = (size_t)warp * per, hi = lo + per < n ? lo + per : n;
    T acc = zero_of(T());
    for (size_t i = lo + lane; i < hi; i += 32) acc = add(acc, in[i]);
    acc = bcast31(warp_inclusive<T>(acc, lane));
    if (lane == 0) wtot[warp] = acc;
    __syncthreads();
    if (warp == 0) {
        const T v = wtot[lane];
        const T inc = warp_inclusive<T>(v, lane);
        wtot[lane] = sub(inc, v);                                  // exclusive base of every warp's slice
        if (lane == 31 && total_out) *total_out = inc;
    }
    __syncthreads();
    T base = wtot[warp];
    for (size_t i0 = lo; i0 < hi; i0 += 32) {
        const size_t i = i0 + lane;
        const T v = i < hi ? in[i] : zero_of(T());
        const T inc = warp_inclusive<T>(v, lane);
        if (i < hi) out[i] = add(base, sub(inc, v));
        base = add(base, bcast31(inc));
    }
}

template <typename T>
int scan_impl(corn_ctx *ctx, const T *d_in, T *d_out, size_t n, T *d_total)
{
    if (n == 0) {
        if (d_total) { k_zero_total<T><<<1, 1, 0, ctx->stream>>>(d_total); corn_count_launch(ctx); CORN_LAUNCH_CHECK(ctx); }
        return CORN_OK;
    }
    size_t nb = (n + SCAN_BLOCK - 1) / SCAN_BLOCK;
    if (nb > 1 && n * sizeof(T) <= (128u << 10)) {   /* one SM streams ~50 GB/s: beyond this the 3-launch scan wins */
        k_scan_single<T><<<1, 1024, 0, ctx->stream>>>(d_in, d_out, n, d_total);
        corn_count_launch(ctx);
        CORN_LAUNCH_CHECK(ctx);
        return CORN_OK;
    }
    if (nb == 1) {
        k_block_scan<T><<<1, SCAN_THREADS, 0, ctx->stream>>>(d_in, d_out, n, (const T *)NULL, d_total);
        corn_count_launch(ctx);
        CORN_LAUNCH_CHECK(ctx);
        return CORN_OK;
    }
    // two temporaries per level; levels shrink by 4096x so two levels cover 2^36 entries
    size_t nb2 = (nb + SCAN_BLOCK - 1) / SCAN_BLOCK;
    CORN_TRY(corn_dbuf_reserve(ctx, &ctx->scan_tmp, sizeof(T) * (nb + nb2 + 2)));
    T *sums = (T *)ctx->scan_tmp.p, *sums2 = sums + nb;
    k_block_sums<T><<<(unsigned)nb, SCAN_THREADS, 0, ctx->stream>>>(d_in, sums, n);
    corn_count_launch(ctx);
    CORN_LAUNCH_CHECK(ctx);
    if (nb2 == 1) {
        k_block_scan<T><<<1, SCAN_THREADS, 0, ctx->stream>>>(sums, sums, nb, (const T *)NULL, (T *)NULL);
        corn_count_launch(ctx);
    } else {
        k_block_sums<T><<<(unsigned)nb2, SCAN_THREADS, 0, ctx->stream>>>(sums, sums2, nb);
        k_block_scan<T><<<1, SCAN_THREADS, 0, ctx->stream>>>(sums2, sums2, nb2, (const T *)NULL, (T *)NULL);
        k_block_scan<T><<<(unsigned)nb2, SCAN_THREADS, 0, ctx->stream>>>(sums, sums, nb, sums2, (T *)NULL);
        corn_count_launch(ctx, 3);
        if (nb2 > SCAN_BLOCK) return corn_set_err(ctx, CORN_E_TOOBIG, "scan of %zu entries", n);
    }
    CORN_LAUNCH_CHECK(ctx);
    k_block_scan<T><<<(unsigned)nb, SCAN_THREADS, 0, ctx->stream>>>(d_in, d_out, n, sums, d_total);
    corn_count_launch(ctx);
    CORN_LAUNCH_CHECK(ctx);
    return CORN_OK;
}

}  // namespace

int corn_scan_u32(corn_ctx *ctx, const uint32_t *d_in, uint32_t *d_out, size_t n, uint32_t *d_total)
{
    return scan_impl<uint32_t>(ctx, d_in, d_out, n, d_total);
}

int corn_scan_u32x4(corn_ctx *ctx, const uint4 *d_in, uint4 *d_out, size_t n, uint4 *d_total)
{
    return scan_impl<uint4>(ctx, d_in, d_out, n, d_total);
}
