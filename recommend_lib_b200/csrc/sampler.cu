// Device-side negative sampler for sm_100a: daisy_sample_triples.
//
// Replaces (reference, file:line): BPRData.ng_sample (util/data_loader.py:680-690: for every training positive (u, i),
// num_ng times, j = np.random.randint(item_num) re-drawn while (u, j) is a training positive) and the
// DataLoader(shuffle=True) permutation of the resulting triples (BPRMFRecommender.py:141-142).  SURVEY.md section 8f,
// row N1: at > 1 G triples/s of training the host sampler + H2D copy is the bottleneck.
//
// The reference draws from unseeded global RNG state, so there is nothing to match bit for bit; this sampler is
// deterministic by construction -- every random number is a pure function of (seed, epoch, slot, attempt) through the
// counter-based Philox4x32-10 generator -- and oracle/sampler_oracle.py restates the same rule in numpy, so device
// and oracle agree bit for bit (tests/test_sampler_gpu.py).
//
//   slot  s = p * num_ng + g          (positive p, g-th negative: the reference's features_fill order)
//   draw  j = mulhi32(philox(key = seed, ctr = (s_lo, s_hi, epoch, attempt))[0], item_num),  attempt = 0, 1, ...
//           until u * item_num + j is not in the sorted positive-key list (binary search)
//   shuffle: slots ordered by (philox(key = seed ^ (0, 0x9E3779B9), ctr = (s_lo, s_hi, epoch, 0xFFFFFFFF))[0], s)
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>

#include "ctx.cuh"

namespace {

struct Philox {
    uint32_t v[4];
};

__device__ __forceinline__ Philox philox4x32_10(uint32_t k0, uint32_t k1, uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3) {
    const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        if (r > 0) {
            k0 += W0;
            k1 += W1;
        }
        const uint32_t hi0 = __umulhi(M0, c0), lo0 = M0 * c0;
        const uint32_t hi1 = __umulhi(M1, c2), lo1 = M1 * c2;
        const uint32_t n0 = hi1 ^ c1 ^ k0, n1 = lo1, n2 = hi0 ^ c3 ^ k1, n3 = lo0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    }
    Philox p;
    p.v[0] = c0; p.v[1] = c1; p.v[2] = c2; p.v[3] = c3;
    return p;
}

__device__ __forceinline__ bool is_positive(const int64_t *__restrict__ keys, int64_t m, int64_t k) {
    int64_t lo = 0, hi = m;
    while (lo < hi) {
        const int64_t mid = (lo + hi) >> 1;
        if (keys[mid] < k) lo = mid + 1; else hi = mid;
    }
    return lo < m && keys[lo] == k;
}

__global__ void k_sample(const int32_t *__restrict__ pairs, long long n_slots, int num_ng, uint32_t item_num,
                         uint32_t user_num, const int64_t *__restrict__ pos_keys, int64_t m, uint32_t k0, uint32_t k1,
                         uint32_t epoch, int32_t *__restrict__ out, uint32_t *__restrict__ skey, uint32_t *__restrict__ sval,
                         int *err) {
    const long long s = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (s >= n_slots) return;
    const long long p = s / num_ng;
    uint32_t u = (uint32_t)pairs[2 * p], i = (uint32_t)pairs[2 * p + 1];
    if (u >= user_num || i >= item_num) {
        atomicOr(&err[0], 1);
        atomicMin(&err[1], (int)(p < 0x7fffffff ? p : 0x7fffffff));
        u = u < user_num ? u : 0;
        i = i < item_num ? i : 0;
    }
    const uint32_t s_lo = (uint32_t)(s & 0xffffffffLL), s_hi = (uint32_t)(s >> 32);
    uint32_t j = 0;
    uint32_t attempt = 0;
    while (true) {
        j = __umulhi(philox4x32_10(k0, k1, s_lo, s_hi, epoch, attempt).v[0], item_num);
        if (m == 0 || !is_positive(pos_keys, m, (int64_t)u * item_num + j)) break;
        if (++attempt == 4096u) {  // a user with (nearly) every item positive: give up, flag it
            atomicOr(&err[0], 64);
            break;
        }
    }
    out[3 * s] = (int32_t)u;
    out[3 * s + 1] = (int32_t)i;
    out[3 * s + 2] = (int32_t)j;
    if (skey) {
        skey[s] = philox4x32_10(k0, k1 ^ 0x9E3779B9u, s_lo, s_hi, epoch, 0xFFFFFFFFu).v[0];
        sval[s] = (uint32_t)s;
    }
}

__global__ void k_permute_triples(const int32_t *__restrict__ in, const uint32_t *__restrict__ order, long long n,
                                  int32_t *__restrict__ out) {
    const long long r = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (r >= n) return;
    const size_t s = order[r];
    out[3 * r] = in[3 * s];
    out[3 * r + 1] = in[3 * s + 1];
    out[3 * r + 2] = in[3 * s + 2];
}

// ---- owner routing of a global epoch (row-sharded training) -------------------------------------
__global__ void k_route_flag(const int32_t *__restrict__ tri, long long n, uint32_t u0, uint32_t u1, uint32_t *__restrict__ flag) {
    const long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (t >= n) return;
    const uint32_t u = (uint32_t)tri[3 * t];
    flag[t] = (u >= u0 && u < u1) ? 1u : 0u;
}
// pos = exclusive prefix sum of flag: a kept triple lands at pos[t] (stable: the order inside a batch is the epoch's),
// and the first triple of global batch b records where the rank's share of that batch starts
__global__ void k_route_scatter(const int32_t *__restrict__ tri, long long n, long long batch, uint32_t u0,
                                const uint32_t *__restrict__ flag, const uint32_t *__restrict__ pos,
                                int32_t *__restrict__ out, int64_t *__restrict__ batch_off, long long n_batches) {
    const long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (t >= n) return;
    const uint32_t p = pos[t];
    if (t % batch == 0) batch_off[t / batch] = (int64_t)p;
    if (t == n - 1) batch_off[n_batches] = (int64_t)p + flag[t];
    if (flag[t]) {
        out[3 * (size_t)p] = (int32_t)((uint32_t)tri[3 * t] - u0);
        out[3 * (size_t)p + 1] = tri[3 * t + 1];
        out[3 * (size_t)p + 2] = tri[3 * t + 2];
    }
}

}  // namespace

extern "C" int daisy_sample_triples(daisy_handle_t h, const int32_t *pairs, int64_t n_pairs, int num_ng,
                                    const int64_t *pos_keys, int64_t n_keys, uint64_t seed, uint32_t epoch, int shuffle,
                                    int32_t *triples_out, daisy_stream_t stream) {
    DAISY_REQUIRE(h != nullptr, DAISY_EINVAL, "null handle");
    DAISY_REQUIRE(n_pairs >= 0 && num_ng >= 1 && n_keys >= 0, DAISY_EINVAL, "bad sizes");
    const long long n = (long long)n_pairs * num_ng;
    // one CUB radix sort shuffles the epoch and its item count is an int
    DAISY_REQUIRE(n < (1LL << 31), DAISY_EUNSUPPORTED, "at most 2^31 - 1 triples per call (%lld asked for)", n);
    if (n == 0) return DAISY_OK;
    DAISY_REQUIRE(pairs && triples_out && (pos_keys || n_keys == 0), DAISY_EINVAL, "null argument");
    DeviceGuard g(h->device);
    DAISY_REQUIRE(g.ok, DAISY_ECUDA, "cannot select device %d", h->device);
    cudaStream_t s = (cudaStream_t)stream;
    const uint32_t k0 = (uint32_t)(seed & 0xffffffffu), k1 = (uint32_t)(seed >> 32);
    const int T = 256;
    const int grid = daisy_ceil_div(n, T);
    if (!shuffle) {
        k_sample<<<grid, T, 0, s>>>(pairs, n, num_ng, (uint32_t)h->I, (uint32_t)h->U, pos_keys, n_keys, k0, k1, epoch,
                                    triples_out, nullptr, nullptr, h->err);
        DAISY_LAUNCH_CHECK(h);
        return DAISY_OK;
    }
    // stream-ordered scratch (cached in the handle's private pool between epochs)
    int32_t *tmp_tri = nullptr;
    uint32_t *key = nullptr, *key_s = nullptr, *val = nullptr, *val_s = nullptr;
    void *cub_tmp = nullptr;
    size_t cub_bytes = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, cub_bytes, key, key_s, val, val_s, (int)n, 0, 32, s);
    bool ok = daisy_scratch_alloc(h, (void **)&tmp_tri, (size_t)n * 3 * sizeof(int32_t), s) == cudaSuccess;
    ok = ok && daisy_scratch_alloc(h, (void **)&key, (size_t)n * 4, s) == cudaSuccess;
    ok = ok && daisy_scratch_alloc(h, (void **)&key_s, (size_t)n * 4, s) == cudaSuccess;
    ok = ok && daisy_scratch_alloc(h, (void **)&val, (size_t)n * 4, s) == cudaSuccess;
    ok = ok && daisy_scratch_alloc(h, (void **)&val_s, (size_t)n * 4, s) == cudaSuccess;
    ok = ok && daisy_scratch_alloc(h, &cub_tmp, cub_bytes + 256, s) == cudaSuccess;
    int rc = DAISY_OK;
    if (!ok) {
        cudaGetLastError();
        daisy_set_error("sampler scratch allocation failed (%lld triples)", n);
        rc = DAISY_ENOMEM;
    } else {
        k_sample<<<grid, T, 0, s>>>(pairs, n, num_ng, (uint32_t)h->I, (uint32_t)h->U, pos_keys, n_keys, k0, k1, epoch,
                                    tmp_tri, key, val, h->err);
        h->launches++;
        size_t tb = cub_bytes + 256;
        cudaError_t e = cub::DeviceRadixSort::SortPairs(cub_tmp, tb, key, key_s, val, val_s, (int)n, 0, 32, s);
        h->launches += 4;
        k_permute_triples<<<grid, T, 0, s>>>(tmp_tri, val_s, n, triples_out);
        h->launches++;
        if (e != cudaSuccess || cudaGetLastError() != cudaSuccess) {
            daisy_set_error("sampler launch failed: %s", cudaGetErrorString(e));
            rc = DAISY_ECUDA;
        }
    }
    for (void *p : {(void *)tmp_tri, (void *)key, (void *)key_s, (void *)val, (void *)val_s, cub_tmp})
        if (p) cudaFreeAsync(p, s);
    return rc;
}

// Row-sharded training (one process per GPU, users block-sharded): the rank's share of a GLOBAL epoch.  triples is the
// epoch every rank holds identically (int32 [n, 3], global ids, already shuffled; device memory); the triples whose
// user lies in [u0, u1) are compacted into out -- order kept, user column made local (u - u0) -- and batch_off
// [ceil(n / batch) + 1] (device, int64) receives where the rank's share of every global batch of `batch` triples
// starts, so that step k of every rank processes exactly its part of global batch k: the sharded step then equals
// daisy_bpr_step on the global batch (tests/test_sharded_gpu.py).  No counterpart in the reference (single device).
extern "C" int daisy_route_triples(daisy_handle_t h, const int32_t *triples, int64_t n, int64_t batch, int64_t u0, int64_t u1,
                                   int32_t *out, int64_t *batch_off, daisy_stream_t stream) {
    DAISY_REQUIRE(h != nullptr, DAISY_EINVAL, "null handle");
    DAISY_REQUIRE(n >= 0 && n < (1LL << 31) && batch >= 1 && u0 >= 0 && u1 >= u0 && u1 < (1LL << 32), DAISY_EINVAL, "bad sizes");
    DAISY_REQUIRE(batch_off != nullptr, DAISY_EINVAL, "null argument");
    DeviceGuard g(h->device);
    DAISY_REQUIRE(g.ok, DAISY_ECUDA, "cannot select device %d", h->device);
    cudaStream_t s = (cudaStream_t)stream;
    if (n == 0) {
        DAISY_CUDA(cudaMemsetAsync(batch_off, 0, sizeof(int64_t), s));
        return DAISY_OK;
    }
    DAISY_REQUIRE(triples && out, DAISY_EINVAL, "null argument");
    const long long n_batches = (n + batch - 1) / batch;
    uint32_t *flag = nullptr, *pos = nullptr;
    void *tmp = nullptr;
    size_t tb = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, tb, flag, pos, (int)n, s);
    bool ok = daisy_scratch_alloc(h, (void **)&flag, (size_t)n * 4, s) == cudaSuccess;
    ok = ok && daisy_scratch_alloc(h, (void **)&pos, (size_t)n * 4, s) == cudaSuccess;
    ok = ok && daisy_scratch_alloc(h, &tmp, tb + 256, s) == cudaSuccess;
    int rc = DAISY_OK;
    if (!ok) {
        cudaGetLastError();
        daisy_set_error("routing scratch allocation failed (%lld triples)", (long long)n);
        rc = DAISY_ENOMEM;
    } else {
        const int T = 256, grid = daisy_ceil_div(n, T);
        k_route_flag<<<grid, T, 0, s>>>(triples, n, (uint32_t)u0, (uint32_t)u1, flag);
        size_t tb2 = tb + 256;
        cudaError_t e = cub::DeviceScan::ExclusiveSum(tmp, tb2, flag, pos, (int)n, s);
        k_route_scatter<<<grid, T, 0, s>>>(triples, n, batch, (uint32_t)u0, flag, pos, out, batch_off, n_batches);
        h->launches += 4;
        if (e != cudaSuccess || cudaGetLastError() != cudaSuccess) {
            daisy_set_error("routing launch failed: %s", cudaGetErrorString(e));
            rc = DAISY_ECUDA;
        }
    }
    for (void *p : {(void *)flag, (void *)pos, tmp})
        if (p) cudaFreeAsync(p, s);
    return rc;
}
