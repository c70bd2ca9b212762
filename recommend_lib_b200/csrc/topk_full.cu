// Full-catalogue scoring + exact top-K for sm_100a: daisy_topk_full.
//
// Replaces (reference, file:line): the final ranking loop BPRMFRecommender.py:196-207 -- one scalar
// model(tensor(u), tensor(i), tensor(i)) forward per (user, candidate) followed by np.argsort -- by
//   k_score_tile   fp32 register-tiled scores S[m, i] = c^2 <P[users[m]], Q[i]> for a tile of users x ALL items
//                  (CUDA cores on purpose: the scores must be fp32 sums of fp32 products for ranking parity),
//   k_mask         optional -inf on each user's training positives (CSR lists),
//   k_select_topk  one block per user: exact radix select (11 + 11 + 10 bits of the order-preserving key) of the
//                  K-th largest score, then collection and a (score desc, item asc) ordering of the K winners.
// The score tile lives in the handle's workspace (<= ~1 GiB), users are processed tile by tile.
#include <float.h>

#include "ctx.cuh"

namespace {

constexpr int BM = 64, BN = 128, BK = 16;

__global__ void __launch_bounds__(256) k_score_tile(const float *__restrict__ P, const float *__restrict__ Q,
                                                     const int32_t *__restrict__ users, int n_users, uint32_t U,
                                                     int64_t I, int D, float c2, float *__restrict__ S, int *err) {
    __shared__ __align__(16) float As[2][BK][BM + 4];
    __shared__ __align__(16) float Bs[2][BK][BN + 4];
    const int tid = threadIdx.x;
    const int m0 = blockIdx.y * BM;
    const int64_t i0 = (int64_t)blockIdx.x * BN;
    const int lr = tid >> 2, lk = (tid & 3) * 4;  // loader coordinates: row 0..63, k offset 0,4,8,12
    // resolve the user row this thread loads (once)
    int64_t urow = -1;
    if (m0 + lr < n_users) {
        uint32_t u = (uint32_t)users[m0 + lr];
        if (u >= U) {
            atomicOr(&err[0], 1);
            atomicMin(&err[1], m0 + lr);
            u = 0;
        }
        urow = (int64_t)u;
    }
    const int64_t it0 = i0 + lr, it1 = i0 + lr + 64;
    const int ty = tid >> 4, tx = tid & 15;
    float acc[4][8];
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 8; ++b) acc[a][b] = 0.f;

    auto gload = [&](int k0, float4 &a, float4 &b0, float4 &b1) {
        const int k = k0 + lk;
        a = (urow >= 0 && k < D) ? *reinterpret_cast<const float4 *>(P + urow * D + k) : f4_zero();
        b0 = (it0 < I && k < D) ? *reinterpret_cast<const float4 *>(Q + it0 * D + k) : f4_zero();
        b1 = (it1 < I && k < D) ? *reinterpret_cast<const float4 *>(Q + it1 * D + k) : f4_zero();
    };
    auto sstore = [&](int buf, const float4 &a, const float4 &b0, const float4 &b1) {
        As[buf][lk + 0][lr] = a.x; As[buf][lk + 1][lr] = a.y; As[buf][lk + 2][lr] = a.z; As[buf][lk + 3][lr] = a.w;
        Bs[buf][lk + 0][lr] = b0.x; Bs[buf][lk + 1][lr] = b0.y; Bs[buf][lk + 2][lr] = b0.z; Bs[buf][lk + 3][lr] = b0.w;
        Bs[buf][lk + 0][lr + 64] = b1.x; Bs[buf][lk + 1][lr + 64] = b1.y; Bs[buf][lk + 2][lr + 64] = b1.z;
        Bs[buf][lk + 3][lr + 64] = b1.w;
    };
    float4 ra, rb0, rb1;
    gload(0, ra, rb0, rb1);
    sstore(0, ra, rb0, rb1);
    __syncthreads();
    const int nk = (D + BK - 1) / BK;
    for (int kt = 0; kt < nk; ++kt) {
        const int buf = kt & 1;
        if (kt + 1 < nk) gload((kt + 1) * BK, ra, rb0, rb1);
#pragma unroll
        for (int kk = 0; kk < BK; ++kk) {
            const float4 a = *reinterpret_cast<const float4 *>(&As[buf][kk][ty * 4]);
            const float4 b0 = *reinterpret_cast<const float4 *>(&Bs[buf][kk][tx * 4]);
            const float4 b1 = *reinterpret_cast<const float4 *>(&Bs[buf][kk][64 + tx * 4]);
            const float av[4] = {a.x, a.y, a.z, a.w};
            const float bv[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
            for (int x = 0; x < 4; ++x)
#pragma unroll
                for (int y = 0; y < 8; ++y) acc[x][y] = fmaf(av[x], bv[y], acc[x][y]);
        }
        if (kt + 1 < nk) {
            sstore(buf ^ 1, ra, rb0, rb1);
            __syncthreads();
        }
    }
#pragma unroll
    for (int x = 0; x < 4; ++x) {
        const int m = m0 + ty * 4 + x;
        if (m >= n_users) continue;
        float *row = S + (size_t)m * (size_t)I;
#pragma unroll
        for (int half = 0; half < 2; ++half)
#pragma unroll
            for (int y = 0; y < 4; ++y) {
                const int64_t i = i0 + half * 64 + tx * 4 + y;
                if (i < I) row[i] = acc[x][half * 4 + y] * c2;
            }
    }
}

__global__ void k_mask(float *__restrict__ S, int64_t I, const int64_t *__restrict__ excl_ptr,
                       const int32_t *__restrict__ excl_idx, int64_t first_user, int n_users, int *err) {
    const int m = blockIdx.x;
    if (m >= n_users) return;
    const int64_t a = excl_ptr[first_user + m], b = excl_ptr[first_user + m + 1];
    for (int64_t e = a + threadIdx.x; e < b; e += blockDim.x) {
        const uint32_t i = (uint32_t)excl_idx[e];
        if ((int64_t)i < I)
            S[(size_t)m * (size_t)I + i] = -INFINITY;
        else {
            atomicOr(&err[0], 1);
            atomicMin(&err[1], (int)(first_user + m));
        }
    }
}

__device__ __forceinline__ uint32_t okey(float f) {  // order-preserving float -> uint32
    const uint32_t u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float okey_inv(uint32_t k) {
    return __uint_as_float((k & 0x80000000u) ? (k & 0x7fffffffu) : ~k);
}

// From a histogram (bins ascending), find the bin holding the `need`-th largest element; returns the bin and
// updates need to the rank inside that bin (1-based count still required from it).
__device__ int pick_bin(const unsigned *hist, int nbins, int &need, unsigned *scratch) {
    // single-thread scan from the top: nbins <= 2048, called by thread 0 only
    int b = nbins - 1;
    int acc = 0;
    for (; b > 0; --b) {
        if (acc + (int)hist[b] >= need) break;
        acc += (int)hist[b];
    }
    need -= acc;
    (void)scratch;
    return b;
}

constexpr int SEL_THREADS = 1024;

__global__ void __launch_bounds__(SEL_THREADS) k_select_topk(const float *__restrict__ S, int64_t I, int K,
                                                             int64_t first_user, int32_t *__restrict__ out_item,
                                                             float *__restrict__ out_score) {
    __shared__ unsigned hist[2048];
    __shared__ uint32_t sel_key[128];
    __shared__ int32_t sel_idx[128];
    __shared__ int s_bin, s_need, s_cnt, s_eqtaken;
    __shared__ int warp_cnt[SEL_THREADS / 32];
    const int tid = threadIdx.x;
    const float *row = S + (size_t)blockIdx.x * (size_t)I;
    const int64_t n = I;

    uint32_t prefix = 0;     // key bits decided so far
    uint32_t pmask = 0;      // which bits of the key are decided
    int need = K;
    const int shifts[3] = {21, 10, 0};
    const int widths[3] = {11, 11, 10};
    for (int pass = 0; pass < 3; ++pass) {
        const int nb = 1 << widths[pass];
        for (int b = tid; b < nb; b += SEL_THREADS) hist[b] = 0;
        __syncthreads();
        for (int64_t i = tid; i < n; i += SEL_THREADS) {
            const uint32_t k = okey(row[i]);
            if ((k & pmask) == prefix) atomicAdd(&hist[(k >> shifts[pass]) & (nb - 1)], 1u);
        }
        __syncthreads();
        if (tid == 0) {
            int nd = need;
            s_bin = pick_bin(hist, nb, nd, nullptr);
            s_need = nd;
        }
        __syncthreads();
        prefix |= (uint32_t)s_bin << shifts[pass];
        pmask |= (uint32_t)(nb - 1) << shifts[pass];
        need = s_need;
        __syncthreads();
    }
    const uint32_t T = prefix;        // key of the K-th largest score
    const int take_eq = need;         // how many elements equal to T belong to the top K (lowest item ids first)
    const int eq_total = (int)hist[T & 1023u];
    if (tid == 0) {
        s_cnt = 0;
        s_eqtaken = 0;
    }
    __syncthreads();
    if (eq_total == take_eq) {  // common case: no tie straddles the K boundary -> one unordered collection pass
        for (int64_t i = tid; i < n; i += SEL_THREADS) {
            const uint32_t k = okey(row[i]);
            if (k >= T) {
                const int slot = atomicAdd(&s_cnt, 1);
                if (slot < 128) {
                    sel_key[slot] = k;
                    sel_idx[slot] = (int32_t)i;
                }
            }
        }
    } else {  // ties across the boundary: equal elements are taken in ascending item order
        for (int64_t base = 0; base < n; base += SEL_THREADS) {
            const int64_t i = base + tid;
            const uint32_t k = (i < n) ? okey(row[i]) : 0u;
            const bool gt = (i < n) && (k > T);
            const bool eq = (i < n) && (k == T);
            if (gt) {
                const int slot = atomicAdd(&s_cnt, 1);
                if (slot < 128) {
                    sel_key[slot] = k;
                    sel_idx[slot] = (int32_t)i;
                }
            }
            // ordered rank of the equal elements of this stripe
            const unsigned bal = __ballot_sync(0xffffffffu, eq);
            const int lane = tid & 31, w = tid >> 5;
            if (lane == 0) warp_cnt[w] = __popc(bal);
            __syncthreads();
            int before = s_eqtaken;
            for (int ww = 0; ww < w; ++ww) before += warp_cnt[ww];
            const int rank = before + __popc(bal & ((1u << lane) - 1u));
            if (eq && rank < take_eq) {
                const int slot = atomicAdd(&s_cnt, 1);
                if (slot < 128) {
                    sel_key[slot] = k;
                    sel_idx[slot] = (int32_t)i;
                }
            }
            __syncthreads();
            if (tid == 0) {
                int tot = 0;
                for (int ww = 0; ww < SEL_THREADS / 32; ++ww) tot += warp_cnt[ww];
                s_eqtaken += tot;
            }
            __syncthreads();
        }
    }
    __syncthreads();
    // order the K winners: (score desc, item asc) by rank counting
    const int cnt = min(s_cnt, 128);
    if (tid < cnt) {
        const uint32_t k = sel_key[tid];
        const int32_t idx = sel_idx[tid];
        int rank = 0;
        for (int j = 0; j < cnt; ++j) {
            const uint32_t kj = sel_key[j];
            const int32_t ij = sel_idx[j];
            rank += (kj > k) || (kj == k && ij < idx);
        }
        if (rank < K) {
            out_item[(size_t)(first_user + blockIdx.x) * K + rank] = idx;
            out_score[(size_t)(first_user + blockIdx.x) * K + rank] = okey_inv(k);
        }
    }
}

}  // namespace

extern "C" int daisy_topk_full(daisy_handle_t h, const float *P, const float *Q, const int32_t *users, int64_t N, int K,
                               const int64_t *excl_ptr, const int32_t *excl_idx, int32_t *out_item, float *out_score,
                               daisy_stream_t stream) {
    DAISY_REQUIRE(h && P && Q, DAISY_EINVAL, "null argument");
    DAISY_REQUIRE(h->D % 4 == 0, DAISY_EUNSUPPORTED, "daisy_topk_full needs dim %% 4 == 0 (dim is %d)", h->D);
    DAISY_REQUIRE(K >= 1 && K <= 128 && (int64_t)K <= h->I, DAISY_EUNSUPPORTED, "top_k %d unsupported (1..min(128, item_num))", K);
    DAISY_REQUIRE(N >= 0 && N < (1LL << 31), DAISY_EINVAL, "bad user count");
    if (N == 0) return DAISY_OK;
    DAISY_REQUIRE(users && out_item && out_score, DAISY_EINVAL, "null argument");
    DAISY_REQUIRE((excl_ptr == nullptr) == (excl_idx == nullptr) || excl_ptr != nullptr, DAISY_EINVAL,
                  "exclusion lists need both excl_ptr and excl_idx");
    DeviceGuard g(h->device);
    DAISY_REQUIRE(g.ok, DAISY_ECUDA, "cannot select device %d", h->device);
    cudaStream_t s = (cudaStream_t)stream;
    const int64_t I = h->I;
    // user tile: as many rows of scores as fit ~1 GiB, at least BM so the score kernel's tiles are full
    int64_t tile = (int64_t)((1ull << 30) / ((uint64_t)I * sizeof(float)));
    if (tile < 1) tile = 1;
    if (tile > 4096) tile = 4096;
    if (tile > N) tile = N;
    const size_t need = (size_t)tile * (size_t)I;
    if (h->scores_cap < need) {
        DAISY_CUDA(cudaStreamSynchronize(s));
        if (h->scores) cudaFree(h->scores);
        h->scores = nullptr;
        h->scores_cap = 0;
        cudaError_t e = cudaMalloc((void **)&h->scores, need * sizeof(float));
        DAISY_REQUIRE(e == cudaSuccess, DAISY_ENOMEM, "cudaMalloc of the %zu-byte score tile failed: %s", need * sizeof(float),
                      cudaGetErrorString(e));
        h->scores_cap = need;
    }
    const float c2 = (float)(h->scale * h->scale);
    for (int64_t first = 0; first < N; first += tile) {
        const int nu = (int)((N - first < tile) ? (N - first) : tile);
        dim3 grid((unsigned)((I + BN - 1) / BN), (unsigned)((nu + BM - 1) / BM));
        k_score_tile<<<grid, 256, 0, s>>>(P, Q, users + first, nu, (uint32_t)h->U, I, h->D, c2, h->scores, h->err);
        DAISY_LAUNCH_CHECK(h);
        if (excl_ptr && excl_idx) {
            k_mask<<<nu, 128, 0, s>>>(h->scores, I, excl_ptr, excl_idx, first, nu, h->err);
            DAISY_LAUNCH_CHECK(h);
        }
        k_select_topk<<<nu, SEL_THREADS, 0, s>>>(h->scores, I, K, first, out_item, out_score);
        DAISY_LAUNCH_CHECK(h);
    }
    return DAISY_OK;
}
