// funk-SVD / RSVD SGD for sm_100a: daisy_mf_fit, daisy_mf_predict.
//
// Replaces (reference, file:line): the per-rating loops of util/matrix_factorization.pyx
//   SVD.fit  :132-151   RSVD.fit  :41-61   SVD.predict :157-167   RSVD.predict :68-78
//
// The reference loop is STRICTLY SEQUENTIAL: the update for rating t+1 reads the rows rating t wrote.  Two
// ratings only interact when they share the user row or the item row, so the loop is a dataflow graph: rating t
// depends on the previous rating of the same user and the previous rating of the same item.  The kernel runs
// that graph directly:
//   * per row a version counter = number of updates applied so far;
//   * for rating t of epoch e the expected versions are   e * (#ratings of the row) + (rank of t among the
//     ratings of the row), computed once per fit with two radix sorts + a max-scan;
//   * persistent warps claim ratings in sequence order from a ticket counter, wait (acquire loads, L2-coherent)
//     until both rows reached the expected version, apply the update in float64, and publish version + 1
//     (release stores).  A claimed ticket only ever waits on lower tickets, which are held by warps that are
//     already running, so the schedule cannot deadlock; a spin cap aborts with an error instead of hanging.
// Independent ratings run in parallel, dependent ones in the reference's order: the result is the sequential
// result up to float64 rounding of the warp-tree dot product (the row updates themselves are element-wise and
// therefore identical).
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>

#include "ctx.cuh"

namespace {

__device__ __forceinline__ unsigned ld_acquire_u32(const unsigned *p) {
    unsigned v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_u32(unsigned *p, unsigned v) {
    asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

struct MfArgs {
    double *pu, *qi, *bu, *bi;
    const int32_t *users, *items;
    const double *ratings;
    const int *need_u, *need_i, *cnt_u, *cnt_i;
    unsigned *ver_u, *ver_i;
    unsigned long long *ticket;
    double *err2;  // [epochs * n] squared errors (may be null)
    int *abort_flag;
    long long n;
    int epochs, D;
    daisy_mf_params prm;
    unsigned spin_cap;
};

__global__ void __launch_bounds__(128) k_mf_dataflow(MfArgs a) {
    const unsigned FULL = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    const unsigned long long total = (unsigned long long)a.n * (unsigned long long)a.epochs;
    const int D = a.D;
    const daisy_mf_params prm = a.prm;
    while (true) {
        unsigned long long tk = 0;
        if (lane == 0) tk = atomicAdd(a.ticket, 1ULL);
        tk = __shfl_sync(FULL, tk, 0);
        if (tk >= total) break;
        const int e = (int)(tk / (unsigned long long)a.n);
        const long long t = (long long)(tk % (unsigned long long)a.n);
        const int u = a.users[t], i = a.items[t];
        const double r = a.ratings[t];
        const unsigned want_u = (unsigned)e * (unsigned)a.cnt_u[u] + (unsigned)a.need_u[t];
        const unsigned want_i = (unsigned)e * (unsigned)a.cnt_i[i] + (unsigned)a.need_i[t];
        int bail = 0;
        if (lane == 0) {
            unsigned spins = 0;
            while (ld_acquire_u32(&a.ver_u[u]) != want_u || ld_acquire_u32(&a.ver_i[i]) != want_i) {
                __nanosleep(32);
                if (++spins > a.spin_cap || *(volatile int *)a.abort_flag) {
                    atomicExch(a.abort_flag, 1);
                    bail = 1;
                    break;
                }
            }
        }
        bail = __shfl_sync(FULL, bail, 0);
        if (bail) break;
        // Every lane acquires the two versions itself (satisfied at once: lane 0 has just seen them).  A shuffle is
        // not a memory barrier -- without this the compiler may hoist the bias / row loads above the wait.
        if (ld_acquire_u32(&a.ver_u[u]) != want_u || ld_acquire_u32(&a.ver_i[i]) != want_i) {
            atomicExch(a.abort_flag, 2);  // cannot happen: versions only advance through this warp
        }
        __syncwarp();
        double *p = a.pu + (size_t)u * D;
        double *q = a.qi + (size_t)i * D;
        // rows live in L2 (ld.cg): other SMs wrote them
        double dot = 0.0;
        for (int f = lane; f < D; f += 32) dot += __ldcg(q + f) * __ldcg(p + f);
        dot = warp_sum_d(dot);
        double err;
        // the two biases are read by lane 0 alone and broadcast: one load each, taken strictly before lane 0's
        // stores below whatever the intra-warp schedule is
        double b0 = 0.0, b1 = 0.0;
        if (lane == 0) {
            b0 = __ldcg(a.bu + u);
            b1 = __ldcg(a.bi + i);
        }
        b0 = __shfl_sync(FULL, b0, 0);
        b1 = __shfl_sync(FULL, b1, 0);
        if (prm.variant == 0) {  // SVD, util/matrix_factorization.pyx:140-151
            const double b_u = b0, b_i = b1;
            err = r - (prm.global_mean + b_u + b_i + dot);
            if (prm.biased && lane == 0) {
                __stcg(a.bu + u, b_u + prm.lr_bu * (err - prm.reg_bu * b_u));
                __stcg(a.bi + i, b_i + prm.lr_bi * (err - prm.reg_bi * b_i));
            }
        } else {  // RSVD, :49-61
            const double c_i = b0, d_j = b1;
            err = r - (c_i + d_j + dot);
            if (prm.variant == 2 && lane == 0) {
                const double inc = prm.lr_bu * (err - prm.reg2 * (c_i + d_j - prm.global_mean));
                __stcg(a.bu + u, c_i + inc);
                __stcg(a.bi + i, d_j + inc);
            }
        }
        for (int f = lane; f < D; f += 32) {
            const double puf = __ldcg(p + f), qif = __ldcg(q + f);
            __stcg(p + f, puf + prm.lr_pu * (err * qif - prm.reg_pu * puf));
            __stcg(q + f, qif + prm.lr_qi * (err * puf - prm.reg_qi * qif));
        }
        if (a.err2 && lane == 0) a.err2[(size_t)e * a.n + t] = err * err;
        __threadfence();
        __syncwarp();
        if (lane == 0) {
            st_release_u32(&a.ver_u[u], want_u + 1);
            st_release_u32(&a.ver_i[i], want_i + 1);
        }
    }
}

// ---- rank of each rating among the ratings of its row, and per-row counts ------------------------
__global__ void k_mf_keys(const int32_t *__restrict__ ids, long long n, unsigned rows, uint32_t *key, uint32_t *val,
                          int *err) {
    long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (t >= n) return;
    uint32_t r = (uint32_t)ids[t];
    if (r >= rows) {
        atomicOr(&err[0], 1);
        atomicMin(&err[1], (int)t);
        r = 0;
    }
    key[t] = r;
    val[t] = (uint32_t)t;
}
__global__ void k_mf_starts(const uint32_t *__restrict__ key_s, long long n, int *start) {
    long long p = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (p >= n) return;
    start[p] = (p == 0 || key_s[p - 1] != key_s[p]) ? (int)p : 0;
}
__global__ void k_mf_need(const uint32_t *__restrict__ key_s, const uint32_t *__restrict__ val_s,
                          const int *__restrict__ seg_start, long long n, int *need, int *cnt) {
    long long p = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (p >= n) return;
    const int rank = (int)p - seg_start[p];
    need[val_s[p]] = rank;
    if (p == n - 1 || key_s[p + 1] != key_s[p]) cnt[key_s[p]] = rank + 1;
}
__global__ void k_mf_sse(const double *__restrict__ err2, long long n, int epochs, double *sse_out) {
    __shared__ double sh[32];
    const int e = blockIdx.x;
    double s = 0.0;
    for (long long t = threadIdx.x; t < n; t += blockDim.x) s += err2[(size_t)e * n + t];
    s = warp_sum_d(s);
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x < 32) {
        double v = (threadIdx.x < (blockDim.x >> 5)) ? sh[threadIdx.x] : 0.0;
        v = warp_sum_d(v);
        if (threadIdx.x == 0) sse_out[e] = v;
    }
}
__global__ void k_mf_predict(const double *__restrict__ pu, const double *__restrict__ qi, const double *__restrict__ bu,
                             const double *__restrict__ bi, const int32_t *__restrict__ users,
                             const int32_t *__restrict__ items, long long n, unsigned U, unsigned I, int D, int with_bias,
                             double mu, double *__restrict__ est, int *err) {
    const int lane = threadIdx.x & 31;
    const long long nw = ((long long)gridDim.x * blockDim.x) >> 5;
    for (long long t = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5; t < n; t += nw) {
        unsigned u = (unsigned)users[t], i = (unsigned)items[t];
        if (u >= U || i >= I) {
            if (lane == 0) {
                atomicOr(&err[0], u >= U ? 2 : 4);  // 2: invalid user code, 4: invalid item code
                atomicMin(&err[1], (int)t);
            }
            u = u < U ? u : 0;
            i = i < I ? i : 0;
        }
        double dot = 0.0;
        for (int f = lane; f < D; f += 32) dot += qi[(size_t)i * D + f] * pu[(size_t)u * D + f];
        dot = warp_sum_d(dot);
        if (lane == 0) est[t] = with_bias ? mu + bu[u] + bi[i] + dot : dot;
    }
}

struct Scratch {  // frees everything on scope exit
    void *p[16];
    int n = 0;
    template <class T>
    int get(T **out, size_t count) {
        *out = nullptr;
        cudaError_t e = cudaMalloc((void **)out, (count ? count : 1) * sizeof(T));
        if (e != cudaSuccess) {
            daisy_set_error("cudaMalloc of %zu bytes failed: %s", count * sizeof(T), cudaGetErrorString(e));
            return DAISY_ENOMEM;
        }
        p[n++] = *out;
        return DAISY_OK;
    }
    ~Scratch() {
        for (int i = 0; i < n; ++i) cudaFree(p[i]);
    }
};

int rank_in_row(daisy_ctx *h, const int32_t *ids, long long n, unsigned rows, uint32_t *key, uint32_t *key_s, uint32_t *val,
                uint32_t *val_s, int *start, int *start_scan, void *tmp, size_t tmp_bytes, int *need, int *cnt,
                cudaStream_t s) {
    const int T = 256;
    const int G = daisy_ceil_div(n, T);
    k_mf_keys<<<G, T, 0, s>>>(ids, n, rows, key, val, h->err);
    DAISY_LAUNCH_CHECK(h);
    int bits = 1;
    while (bits < 32 && ((uint64_t)(rows - 1) >> bits)) ++bits;
    size_t tb = tmp_bytes;
    DAISY_CUDA(cub::DeviceRadixSort::SortPairs(tmp, tb, key, key_s, val, val_s, (int)n, 0, bits, s));
    k_mf_starts<<<G, T, 0, s>>>(key_s, n, start);
    DAISY_LAUNCH_CHECK(h);
    tb = tmp_bytes;
    DAISY_CUDA(cub::DeviceScan::InclusiveScan(tmp, tb, start, start_scan, cub::Max(), (int)n, s));
    k_mf_need<<<G, T, 0, s>>>(key_s, val_s, start_scan, n, need, cnt);
    DAISY_LAUNCH_CHECK(h);
    h->launches += 8;
    return DAISY_OK;
}

}  // namespace

extern "C" int daisy_mf_fit(daisy_handle_t h, double *pu, double *qi, double *bu, double *bi, const int32_t *users,
                            const int32_t *items, const double *ratings, int64_t n, int n_epochs,
                            const daisy_mf_params *prm, double *sse_out, daisy_stream_t stream) {
    DAISY_REQUIRE(h && pu && qi && bu && bi && prm, DAISY_EINVAL, "null argument");
    DAISY_REQUIRE(n >= 0 && n < (1LL << 30), DAISY_EINVAL, "bad rating count");
    DAISY_REQUIRE(n_epochs >= 0, DAISY_EINVAL, "bad epoch count");
    DAISY_REQUIRE(prm->variant >= 0 && prm->variant <= 2, DAISY_EINVAL, "variant must be 0 (SVD), 1 (RSVD1) or 2 (RSVD2)");
    DAISY_REQUIRE((double)n * (double)n_epochs < 4.0e9, DAISY_EUNSUPPORTED, "n * n_epochs must stay below 2^32 row versions");
    if (n == 0 || n_epochs == 0) return DAISY_OK;
    DAISY_REQUIRE(users && items && ratings, DAISY_EINVAL, "null rating arrays");
    DeviceGuard g(h->device);
    DAISY_REQUIRE(g.ok, DAISY_ECUDA, "cannot select device %d", h->device);
    cudaStream_t s = (cudaStream_t)stream;
    const size_t N = (size_t)n;
    Scratch ws;
    uint32_t *key, *key_s, *val, *val_s;
    int *start, *start_scan, *need_u, *need_i, *cnt_u, *cnt_i, *abort_flag;
    unsigned *ver_u, *ver_i;
    unsigned long long *ticket;
    double *err2 = nullptr;
    void *tmp;
    int rc = DAISY_OK;
#define G_(x, c) if (!rc) rc = ws.get(&x, (c))
    G_(key, N); G_(key_s, N); G_(val, N); G_(val_s, N); G_(start, N); G_(start_scan, N);
    G_(need_u, N); G_(need_i, N); G_(cnt_u, (size_t)h->U); G_(cnt_i, (size_t)h->I);
    G_(ver_u, (size_t)h->U); G_(ver_i, (size_t)h->I); G_(ticket, 1); G_(abort_flag, 1);
    if (sse_out) G_(err2, N * (size_t)n_epochs);
#undef G_
    if (rc) return rc;
    size_t tb_sort = 0, tb_scan = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, tb_sort, key, key_s, val, val_s, (int)n, 0, 32, s);
    cub::DeviceScan::InclusiveScan(nullptr, tb_scan, start, start_scan, cub::Max(), (int)n, s);
    const size_t tmp_bytes = (tb_sort > tb_scan ? tb_sort : tb_scan) + 256;
    char *tmpc;
    rc = ws.get(&tmpc, tmp_bytes);
    if (rc) return rc;
    tmp = tmpc;
    DAISY_CUDA(cudaMemsetAsync(cnt_u, 0, (size_t)h->U * sizeof(int), s));
    DAISY_CUDA(cudaMemsetAsync(cnt_i, 0, (size_t)h->I * sizeof(int), s));
    DAISY_CUDA(cudaMemsetAsync(ver_u, 0, (size_t)h->U * sizeof(unsigned), s));
    DAISY_CUDA(cudaMemsetAsync(ver_i, 0, (size_t)h->I * sizeof(unsigned), s));
    DAISY_CUDA(cudaMemsetAsync(ticket, 0, sizeof(unsigned long long), s));
    DAISY_CUDA(cudaMemsetAsync(abort_flag, 0, sizeof(int), s));
    rc = rank_in_row(h, users, n, (unsigned)h->U, key, key_s, val, val_s, start, start_scan, tmp, tmp_bytes, need_u, cnt_u, s);
    if (rc) return rc;
    rc = rank_in_row(h, items, n, (unsigned)h->I, key, key_s, val, val_s, start, start_scan, tmp, tmp_bytes, need_i, cnt_i, s);
    if (rc) return rc;
    // ids out of range were parked on row 0 by k_mf_keys but the dataflow kernel indexes with the raw ids:
    // stop here and let daisy_check report it
    DAISY_CUDA(cudaMemcpyAsync(h->err_host, h->err, 2 * sizeof(int), cudaMemcpyDeviceToHost, s));
    DAISY_CUDA(cudaStreamSynchronize(s));
    if (h->err_host[0]) {
        daisy_set_error("Invalid user or item code at rating %d", h->err_host[1]);
        return DAISY_EINDEX;
    }
    MfArgs a;
    a.pu = pu; a.qi = qi; a.bu = bu; a.bi = bi;
    a.users = users; a.items = items; a.ratings = ratings;
    a.need_u = need_u; a.need_i = need_i; a.cnt_u = cnt_u; a.cnt_i = cnt_i;
    a.ver_u = ver_u; a.ver_i = ver_i; a.ticket = ticket; a.err2 = err2; a.abort_flag = abort_flag;
    a.n = n; a.epochs = n_epochs; a.D = h->D; a.prm = *prm;
    a.spin_cap = 1u << 24;  // ~1 s of polling: a correct schedule never gets near it
    int occ = 0;
    DAISY_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_mf_dataflow, 128, 0));
    if (occ < 1) occ = 1;
    int per_sm = occ < 8 ? occ : 8;
    const char *env = getenv("DAISY_MF_BLOCKS_PER_SM");
    if (env && atoi(env) > 0) per_sm = atoi(env) < occ ? atoi(env) : occ;
    k_mf_dataflow<<<h->num_sms * per_sm, 128, 0, s>>>(a);
    DAISY_LAUNCH_CHECK(h);
    if (sse_out) {
        k_mf_sse<<<n_epochs, 256, 0, s>>>(err2, n, n_epochs, sse_out);
        DAISY_LAUNCH_CHECK(h);
    }
    int aborted = 0;
    DAISY_CUDA(cudaMemcpyAsync(&aborted, abort_flag, sizeof(int), cudaMemcpyDeviceToHost, s));
    DAISY_CUDA(cudaStreamSynchronize(s));
    DAISY_REQUIRE(!aborted, DAISY_ECUDA, "daisy_mf_fit: dataflow schedule stalled (spin cap reached) -- tables are partial");
    return DAISY_OK;
}

extern "C" int daisy_mf_predict(daisy_handle_t h, const double *pu, const double *qi, const double *bu, const double *bi,
                                const int32_t *users, const int32_t *items, int64_t n, int with_bias, double mu,
                                double *est, daisy_stream_t stream) {
    DAISY_REQUIRE(h && pu && qi, DAISY_EINVAL, "null argument");
    DAISY_REQUIRE(!with_bias || (bu && bi), DAISY_EINVAL, "bias arrays required when with_bias");
    DAISY_REQUIRE(n >= 0, DAISY_EINVAL, "bad count");
    if (n == 0) return DAISY_OK;
    DAISY_REQUIRE(users && items && est, DAISY_EINVAL, "null argument");
    DeviceGuard g(h->device);
    long long blocks = (n + 7) / 8;
    const long long cap = (long long)h->num_sms * 16;
    if (blocks > cap) blocks = cap;
    k_mf_predict<<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(pu, qi, bu, bi, users, items, n, (unsigned)h->U,
                                                              (unsigned)h->I, h->D, with_bias, mu, est, h->err);
    DAISY_LAUNCH_CHECK(h);
    return DAISY_OK;
}
