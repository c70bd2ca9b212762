// Item2Vec / skip-gram with negative sampling: the training step of
// Item2VecRecommender.py:60-97 (Item2Vec.forward_i / forward_o, SGNS.forward) + :266, 274-277 (dense torch Adam).
//
// STATUS: GPU-verified in round 2 (tests/test_sgns_gpu.py passes on a B200 and runs in the default -m gpu suite; bench line under
// profiles/r02a_bench_sgns.json).  Before that it had run under the host emulation of tests/emu
// (tests/test_kernel_emulation.py: golden run + oracle green), which still covers it on CPU.
// The checker exists and is pinned to the unmodified reference: oracle/sgns_oracle.py.
//
// Per example b: one centre row i_b = ivectors[iword_b], C context rows and C * n_negs negative rows of ovectors
// (R = C (1 + n_negs) = 210 at the script's defaults).  The negatives are GIVEN (the reference draws them inside forward):
//     loss = (1 / (B C)) sum_b [ sum_c softplus(-o_bc . i_b) + sum_cn softplus(n_bcn . i_b) ]
// A ref's gradient contribution to its output row is coef * i_b and to the centre row coef * o_row, so only the
// coefficient of every ref is stored (4 bytes instead of a D-float row):
//   k_sgns_main     one warp per example: i_b in registers, R rows streamed, coefficients, centre-row gradient, loss
//   cub sort x 2    (output row, ref) and (centre row, example), stable
//   k_sgns_seg / k_sgns_scan / k_sgns_slices / k_sgns_combine   every row's segment of the sorted list is cut into slices
//                   of SG_SLICE refs -- as many slices as the row is long -- one warp per slice, partial sums to scratch,
//                   then one warp per row adds its slices in slice order -> dense gradient rows, fixed order =>
//                   bit-reproducible.  (The first version cut EVERY row into 8 slices inside one block: under a
//                   unigram^0.75 sampler the hottest rows hold 20-40 k of the 860 k refs, and their blocks made the
//                   kernel 1.93 ms of a 2.36 ms step, profiles/r02u_launches_sgns_summary.md.)  The padding row
//                   (nn.Embedding(padding_idx=0), :40-41) gets a zero gradient and is not summed at all
//   k_sgns_adam     torch.optim.Adam is dense: every element of both tables is stepped
#include <cub/cub.cuh>

#include "ctx.cuh"

namespace {

constexpr int SG_MAX_D = 512;            // 16 lane-strided values per lane
constexpr int SG_K = SG_MAX_D / 32;

struct SgScratch {
    float *coef, *ci, *lossp, *gI, *gO;
    uint32_t *kin, *kout, *vin, *vout, *ikin, *ikout, *ivin, *ivout;
    uint32_t *seg_lo, *seg_len, *soff;   // [2V], [2V], [2V + 1]: per (table, row) first ref, refs, first slice
    float *partial;                      // [max slices][D] slice sums
    void *cub;
    size_t cub_bytes, total;
};

constexpr int SG_SLICE = 128;            // refs per slice of a row's segment (one warp per slice)
// a row of len refs has ceil(len / SG_SLICE) slices: at most (refs / SG_SLICE + 1) per row over both tables
static size_t sg_max_slices(int64_t n, int64_t B, int64_t V) {
    return (size_t)(n / SG_SLICE + B / SG_SLICE + 2 * V + 2);
}

// bound on cub::DeviceRadixSort::SortPairs temporary storage for n (uint32, uint32) pairs (see csrc/fmbn.cu)
static size_t sg_sort_bytes(int64_t n) { return (size_t)n * 16 + ((size_t)1 << 20); }

static void sg_carve(char *base, int64_t B, int64_t R, int64_t V, int D, SgScratch &w) {
    size_t off = 0;
    auto take = [&](size_t bytes) {
        char *p = base ? base + off : nullptr;
        off += (bytes + 255) / 256 * 256;
        return p;
    };
    const size_t n = (size_t)B * (size_t)R, b = (size_t)B, vd = (size_t)V * (size_t)D;
    w.coef = (float *)take(n * 4);
    w.ci = (float *)take(b * (size_t)D * 4);
    w.lossp = (float *)take(b * 4);
    w.gI = (float *)take(vd * 4);
    w.gO = (float *)take(vd * 4);
    w.kin = (uint32_t *)take(n * 4);
    w.kout = (uint32_t *)take(n * 4);
    w.vin = (uint32_t *)take(n * 4);
    w.vout = (uint32_t *)take(n * 4);
    w.ikin = (uint32_t *)take(b * 4);
    w.ikout = (uint32_t *)take(b * 4);
    w.ivin = (uint32_t *)take(b * 4);
    w.ivout = (uint32_t *)take(b * 4);
    w.seg_lo = (uint32_t *)take(2 * (size_t)V * 4);
    w.seg_len = (uint32_t *)take(2 * (size_t)V * 4);
    w.soff = (uint32_t *)take((2 * (size_t)V + 1) * 4);
    w.partial = (float *)take(sg_max_slices((int64_t)n, B, V) * (size_t)D * 4);
    w.cub_bytes = sg_sort_bytes((int64_t)n);
    w.cub = take(w.cub_bytes);
    w.total = off;
}

__device__ __forceinline__ uint32_t sg_row(int32_t id, uint32_t V, int pos, int *err) {
    const uint32_t r = (uint32_t)id;
    if (r >= V) {  // never fault: park on row 0, the error flag tells the caller
        atomicOr(&err[0], 1);
        atomicMin(&err[1], pos);
        return 0u;
    }
    return r;
}

__device__ __forceinline__ float sg_softplus(float x) { return fmaxf(x, 0.f) + log1pf(expf(-fabsf(x))); }

// one warp per example; K = lane-strided values per lane (D <= 32 K), so that a short row does not pay for 16 registers
// per row held
template <int K>
__global__ void __launch_bounds__(256) k_sgns_main(const float *__restrict__ iv, const float *__restrict__ ov,
                                                    const int32_t *__restrict__ iword, const int32_t *__restrict__ owords,
                                                    const int32_t *__restrict__ nwords, int B, int C, int R, uint32_t V, int D,
                                                    float inv_bc, float *__restrict__ coef, float *__restrict__ ci,
                                                    float *__restrict__ lossp, uint32_t *__restrict__ kin,
                                                    uint32_t *__restrict__ vin, uint32_t *__restrict__ ikin,
                                                    uint32_t *__restrict__ ivin, int *err) {
    const int lane = threadIdx.x & 31;
    const int b = (int)((blockIdx.x * (size_t)blockDim.x + threadIdx.x) >> 5);
    if (b >= B) return;
    uint32_t w = 0;
    if (lane == 0) w = sg_row(iword[b], V, b, err);
    w = __shfl_sync(0xffffffffu, w, 0);
    float ir[K], gi[K];
#pragma unroll
    for (int k = 0; k < K; ++k) {
        const int f = lane + 32 * k;
        ir[k] = f < D ? iv[(size_t)w * D + f] : 0.f;
        gi[k] = 0.f;
    }
    float loss = 0.f;
    const int NR = R - C;  // negatives per example
    // Ids: lane l holds the (validated) row of ref r0 + l of the current group of 32 refs (one coalesced load per group,
    // the next group's ids fetched a group ahead).  Rows: the row of ref r + 1 is loaded before ref r is consumed, so two
    // rows of a warp are in flight (the first version loaded and consumed one row at a time: ~1 us per ref, 226 us for
    // the kernel at the script's defaults, profiles/r02u_launches_sgns_summary.md).  Same arithmetic in the same order.
    auto load_ids = [&](int r0) -> uint32_t {
        const int rr = r0 + lane;
        return rr < R ? sg_row(rr < C ? owords[(size_t)b * C + rr] : nwords[(size_t)b * NR + (rr - C)], V, b, err) : 0u;
    };
    uint32_t ids = load_ids(0), ids_next = load_ids(32);
    float nrow[K];
    uint32_t row_next = __shfl_sync(0xffffffffu, ids, 0);
    {
        const float *o = ov + (size_t)row_next * D;
#pragma unroll
        for (int k = 0; k < K; ++k) {
            const int f = lane + 32 * k;
            nrow[k] = f < D ? o[f] : 0.f;
        }
    }
    for (int r = 0; r < R; ++r) {
        const uint32_t row = row_next;
        float orow[K];
#pragma unroll
        for (int k = 0; k < K; ++k) orow[k] = nrow[k];
        if (r + 1 < R) {  // warp-uniform
            if (((r + 1) & 31) == 0) {
                ids = ids_next;
                ids_next = load_ids(r + 1 + 32);
            }
            row_next = __shfl_sync(0xffffffffu, ids, (r + 1) & 31);
            const float *o = ov + (size_t)row_next * D;
#pragma unroll
            for (int k = 0; k < K; ++k) {
                const int f = lane + 32 * k;
                nrow[k] = f < D ? o[f] : 0.f;
            }
        }
        float dot = 0.f;
#pragma unroll
        for (int k = 0; k < K; ++k) dot += ir[k] * orow[k];
        dot = warp_sum(dot);
        float c;
        if (r < C) {  // context row: -log sigmoid(x)
            c = -inv_bc / (1.f + expf(dot));
            loss += sg_softplus(-dot);
        } else {      // negative row: -log sigmoid(-y)
            c = inv_bc / (1.f + expf(-dot));
            loss += sg_softplus(dot);
        }
#pragma unroll
        for (int k = 0; k < K; ++k) gi[k] += c * orow[k];
        if (lane == 0) {
            const size_t p = (size_t)b * R + r;
            coef[p] = c;
            kin[p] = row;
            vin[p] = (uint32_t)p;
        }
    }
#pragma unroll
    for (int k = 0; k < K; ++k) {
        const int f = lane + 32 * k;
        if (f < D) ci[(size_t)b * D + f] = gi[k];
    }
    if (lane == 0) {
        lossp[b] = loss;
        ikin[b] = w;
        ivin[b] = (uint32_t)b;
    }
}

__device__ __forceinline__ int sg_lower_bound(const uint32_t *__restrict__ a, int n, uint32_t key) {
    int lo = 0, hi = n;
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (a[mid] < key) lo = mid + 1; else hi = mid;
    }
    return lo;
}

// idx = table * V + row, table 0: gradient of ovectors[row] = sum over its refs of coef * ivectors[iword_b] (pre-step);
//                         table 1: gradient of ivectors[row] = sum over the examples centred on it of their ci rows.
__global__ void k_sgns_seg(const uint32_t *__restrict__ kout, int n, const uint32_t *__restrict__ ikout, int B, uint32_t V,
                           int padding_idx, uint32_t *__restrict__ seg_lo, uint32_t *__restrict__ seg_len,
                           uint32_t *__restrict__ nsl) {
    const uint32_t idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= 2u * V) return;
    const bool otab = idx < V;
    const uint32_t row = otab ? idx : idx - V;
    const uint32_t *keys = otab ? kout : ikout;
    const int len_all = otab ? n : B;
    const int lo = sg_lower_bound(keys, len_all, row), hi = sg_lower_bound(keys, len_all, row + 1u);
    seg_lo[idx] = (uint32_t)lo;
    seg_len[idx] = (uint32_t)(hi - lo);
    nsl[idx] = ((int)row == padding_idx) ? 0u : (uint32_t)((hi - lo + SG_SLICE - 1) / SG_SLICE);
}

// exclusive scan of the 2V slice counts, in place, plus the total behind them (one block; 2V is a few thousand)
__global__ void __launch_bounds__(1024) k_sgns_scan(uint32_t *__restrict__ soff, uint32_t count) {
    __shared__ uint32_t part[1024];
    const uint32_t per = (count + 1023u) / 1024u, t = threadIdx.x;
    const uint32_t b = t * per, e = min(count, b + per);
    uint32_t sum = 0;
    for (uint32_t i = b; i < e; ++i) sum += soff[i];
    part[t] = sum;
    __syncthreads();
    for (uint32_t o = 1; o < 1024u; o <<= 1) {
        const uint32_t v = t >= o ? part[t - o] : 0u;
        __syncthreads();
        part[t] += v;
        __syncthreads();
    }
    uint32_t run = part[t] - sum;  // exclusive prefix of this thread's chunk
    for (uint32_t i = b; i < e; ++i) {
        const uint32_t c = soff[i];
        soff[i] = run;
        run += c;
    }
    if (t == 1023u) soff[count] = part[1023];
}

// one warp per slice (grid-stride): the slice's refs summed in list order -> partial[slice]
template <int K>
__global__ void __launch_bounds__(256) k_sgns_slices(const uint32_t *__restrict__ vout, const uint32_t *__restrict__ ivout,
                                                      int R, const float *__restrict__ coef, const float *__restrict__ ci,
                                                      const float *__restrict__ iv, const uint32_t *__restrict__ ikin, int D,
                                                      uint32_t V, const uint32_t *__restrict__ seg_lo,
                                                      const uint32_t *__restrict__ seg_len, const uint32_t *__restrict__ soff,
                                                      float *__restrict__ partial) {
    const int lane = threadIdx.x & 31;
    const uint32_t warp = (uint32_t)((blockIdx.x * (size_t)blockDim.x + threadIdx.x) >> 5);
    const uint32_t nwarps = (uint32_t)((gridDim.x * (size_t)blockDim.x) >> 5);
    const uint32_t total = soff[2u * V];
    for (uint32_t t = warp; t < total; t += nwarps) {
        uint32_t lo = 0, hi = 2u * V;  // last idx with soff[idx] <= t (rows without slices share their successor's offset)
        while (hi - lo > 1u) {
            const uint32_t mid = (lo + hi) >> 1;
            if (soff[mid] <= t) lo = mid; else hi = mid;
        }
        const uint32_t idx = lo;
        const bool otab = idx < V;
        const uint32_t j = t - soff[idx];
        const uint32_t q0 = seg_lo[idx] + j * SG_SLICE;
        const uint32_t q1 = min(seg_lo[idx] + seg_len[idx], q0 + SG_SLICE);
        float acc[K];
#pragma unroll
        for (int k = 0; k < K; ++k) acc[k] = 0.f;
#pragma unroll 4  // independent index -> row chains: let the scheduler put several rows in flight
        for (uint32_t q = q0; q < q1; ++q) {
            if (otab) {
                const uint32_t p = vout[q];
                const float c = coef[p];
                const float *src = iv + (size_t)ikin[p / (uint32_t)R] * D;  // ikin[b] = validated centre row of example b
#pragma unroll
                for (int k = 0; k < K; ++k) {
                    const int f = lane + 32 * k;
                    if (f < D) acc[k] += c * src[f];
                }
            } else {
                const float *src = ci + (size_t)ivout[q] * D;
#pragma unroll
                for (int k = 0; k < K; ++k) {
                    const int f = lane + 32 * k;
                    if (f < D) acc[k] += src[f];
                }
            }
        }
        float *dst = partial + (size_t)t * D;
#pragma unroll
        for (int k = 0; k < K; ++k) {
            const int f = lane + 32 * k;
            if (f < D) dst[f] = acc[k];
        }
    }
}

// one warp per (table, row): its slices added in slice order -> dense gradient row (zero for the padding row and for
// rows without refs)
template <int K>
__global__ void __launch_bounds__(256) k_sgns_combine(const uint32_t *__restrict__ soff, uint32_t V, int D,
                                                       const float *__restrict__ partial, float *__restrict__ gO,
                                                       float *__restrict__ gI) {
    const int lane = threadIdx.x & 31;
    const uint32_t idx = (uint32_t)((blockIdx.x * (size_t)blockDim.x + threadIdx.x) >> 5);
    if (idx >= 2u * V) return;
    const uint32_t s0 = soff[idx], s1 = soff[idx + 1];
    float acc[K];
#pragma unroll
    for (int k = 0; k < K; ++k) acc[k] = 0.f;
    for (uint32_t sl = s0; sl < s1; ++sl) {
        const float *src = partial + (size_t)sl * D;
#pragma unroll
        for (int k = 0; k < K; ++k) {
            const int f = lane + 32 * k;
            if (f < D) acc[k] += src[f];
        }
    }
    float *g = (idx < V ? gO + (size_t)idx * D : gI + (size_t)(idx - V) * D);
#pragma unroll
    for (int k = 0; k < K; ++k) {
        const int f = lane + 32 * k;
        if (f < D) g[f] = acc[k];
    }
}

struct SgAdam {
    float b1, b2, step_size, inv_sqrt_bc2, eps;
};

// torch.optim.Adam, single-tensor form: exp_avg, exp_avg_sq, denom = sqrt(exp_avg_sq) / sqrt(bc2) + eps
__global__ void k_sgns_adam(float *__restrict__ iv, float *__restrict__ m_iv, float *__restrict__ v_iv,
                            const float *__restrict__ gI, float *__restrict__ ov, float *__restrict__ m_ov,
                            float *__restrict__ v_ov, const float *__restrict__ gO, size_t nvd, SgAdam a) {
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t k = (size_t)blockIdx.x * blockDim.x + threadIdx.x; k < 2 * nvd; k += stride) {
        const bool o = k >= nvd;
        const size_t e = o ? k - nvd : k;
        float *T = o ? ov : iv, *M = o ? m_ov : m_iv, *Vv = o ? v_ov : v_iv;
        const float g = (o ? gO : gI)[e];
        const float m = a.b1 * M[e] + (1.f - a.b1) * g;
        const float v = a.b2 * Vv[e] + (1.f - a.b2) * g * g;
        M[e] = m;
        Vv[e] = v;
        T[e] -= a.step_size * m / (sqrtf(v) * a.inv_sqrt_bc2 + a.eps);
    }
}

__global__ void __launch_bounds__(256) k_sgns_loss(const float *__restrict__ lossp, int B, double scale, double *loss_accum) {
    __shared__ double sh[8];
    double t = 0.0;
    for (int b = threadIdx.x; b < B; b += 256) t += (double)lossp[b];
    t = warp_sum_d(t);
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = t;
    __syncthreads();
    if (threadIdx.x == 0) {
        double tot = 0.0;
        for (int k = 0; k < 8; ++k) tot += sh[k];
        *loss_accum += tot * scale;
    }
}

static int sg_bits(int64_t v) {
    int bits = 1;
    while (bits < 32 && (1ll << bits) < v) ++bits;
    return bits;
}

}  // namespace

extern "C" int daisy_sgns_scratch_bytes(int64_t B, int C, int n_negs, int64_t vocab, int D, int64_t *bytes) {
    DAISY_REQUIRE(bytes != nullptr && B >= 0 && C >= 1 && n_negs >= 0 && vocab >= 1 && D >= 1 && D <= SG_MAX_D, DAISY_EINVAL,
                  "bad scratch query");
    SgScratch w;
    sg_carve(nullptr, B > 0 ? B : 1, (int64_t)C * (1 + n_negs), vocab, D, w);
    *bytes = (int64_t)w.total;
    return DAISY_OK;
}

extern "C" int daisy_sgns_step(daisy_handle_t h, const daisy_sgns_params *p, const int32_t *iword, const int32_t *owords,
                               const int32_t *nwords, int64_t B, int C, int n_negs, int64_t step_no, void *scratch,
                               int64_t scratch_bytes, double *loss_accum, daisy_stream_t stream) {
    DAISY_REQUIRE(h != nullptr && p != nullptr, DAISY_EINVAL, "null handle or parameter block");
    DAISY_REQUIRE(p->iv && p->ov && p->m_iv && p->v_iv && p->m_ov && p->v_ov, DAISY_EINVAL, "null table / Adam moment pointer");
    DAISY_REQUIRE(p->D >= 1 && p->D <= SG_MAX_D, DAISY_EUNSUPPORTED, "embedding size %d out of range (1..%d)", p->D, SG_MAX_D);
    DAISY_REQUIRE(p->vocab >= 1 && p->vocab < 0x7fffffffLL, DAISY_EINVAL, "bad vocabulary size %lld", (long long)p->vocab);
    DAISY_REQUIRE(C >= 1 && n_negs >= 0 && step_no >= 1, DAISY_EINVAL, "bad context size / negatives / step number");
    const int64_t R = (int64_t)C * (1 + n_negs);
    DAISY_REQUIRE(B >= 0 && B * R < 0x7fffffffLL, DAISY_EINVAL, "batch %lld x %lld refs out of range", (long long)B, (long long)R);
    DAISY_REQUIRE((uintptr_t)scratch % 256 == 0 && scratch != nullptr, DAISY_EINVAL, "scratch must be 256-byte aligned");
    SgScratch w;
    sg_carve((char *)scratch, B > 0 ? B : 1, R, p->vocab, p->D, w);
    DAISY_REQUIRE((int64_t)w.total <= scratch_bytes, DAISY_EINVAL, "scratch of %lld bytes, %zu needed (daisy_sgns_scratch_bytes)",
                  (long long)scratch_bytes, w.total);
    DAISY_REQUIRE(B == 0 || (iword && owords && (n_negs == 0 || nwords)), DAISY_EINVAL, "null id arrays");
    DeviceGuard g(h->device);
    DAISY_REQUIRE(g.ok, DAISY_ECUDA, "cannot select device %d", h->device);
    cudaStream_t s = (cudaStream_t)stream;
    const int Bi = (int)B, D = p->D, n = (int)(B * R);
    const uint32_t V = (uint32_t)p->vocab;
    const size_t nvd = (size_t)p->vocab * D;
    if (B > 0) {
        const float inv_bc = 1.0f / ((float)B * (float)C);
        const int mblocks = daisy_ceil_div(B, 8);
        if (D <= 128)
            k_sgns_main<4><<<mblocks, 256, 0, s>>>(p->iv, p->ov, iword, owords, nwords, Bi, C, (int)R, V, D, inv_bc, w.coef, w.ci,
                                                   w.lossp, w.kin, w.vin, w.ikin, w.ivin, h->err);
        else if (D <= 320)
            k_sgns_main<10><<<mblocks, 256, 0, s>>>(p->iv, p->ov, iword, owords, nwords, Bi, C, (int)R, V, D, inv_bc, w.coef, w.ci,
                                                    w.lossp, w.kin, w.vin, w.ikin, w.ivin, h->err);
        else
            k_sgns_main<SG_K><<<mblocks, 256, 0, s>>>(p->iv, p->ov, iword, owords, nwords, Bi, C, (int)R, V, D, inv_bc, w.coef,
                                                      w.ci, w.lossp, w.kin, w.vin, w.ikin, w.ivin, h->err);
        DAISY_LAUNCH_CHECK(h);
        const int bits = sg_bits(p->vocab);
        size_t cub_bytes = w.cub_bytes;
        DAISY_CUDA(cub::DeviceRadixSort::SortPairs(w.cub, cub_bytes, w.kin, w.kout, w.vin, w.vout, n, 0, bits, s));
        cub_bytes = w.cub_bytes;
        DAISY_CUDA(cub::DeviceRadixSort::SortPairs(w.cub, cub_bytes, w.ikin, w.ikout, w.ivin, w.ivout, Bi, 0, bits, s));
        h->launches += 6;
        const int T = 256;
        k_sgns_seg<<<daisy_ceil_div(2 * (int64_t)V, T), T, 0, s>>>(w.kout, n, w.ikout, Bi, V, p->padding_idx, w.seg_lo, w.seg_len,
                                                                   w.soff);
        DAISY_LAUNCH_CHECK(h);
        k_sgns_scan<<<1, 1024, 0, s>>>(w.soff, 2u * V);
        DAISY_LAUNCH_CHECK(h);
        const int sgrid = h->num_sms > 0 ? h->num_sms * 8 : 8, cgrid = daisy_ceil_div(2 * (int64_t)V, 8);
        if (D <= 128) {
            k_sgns_slices<4><<<sgrid, 256, 0, s>>>(w.vout, w.ivout, (int)R, w.coef, w.ci, p->iv, w.ikin, D, V, w.seg_lo, w.seg_len,
                                                   w.soff, w.partial);
            k_sgns_combine<4><<<cgrid, 256, 0, s>>>(w.soff, V, D, w.partial, w.gO, w.gI);
        } else if (D <= 320) {
            k_sgns_slices<10><<<sgrid, 256, 0, s>>>(w.vout, w.ivout, (int)R, w.coef, w.ci, p->iv, w.ikin, D, V, w.seg_lo, w.seg_len,
                                                    w.soff, w.partial);
            k_sgns_combine<10><<<cgrid, 256, 0, s>>>(w.soff, V, D, w.partial, w.gO, w.gI);
        } else {
            k_sgns_slices<SG_K><<<sgrid, 256, 0, s>>>(w.vout, w.ivout, (int)R, w.coef, w.ci, p->iv, w.ikin, D, V, w.seg_lo,
                                                      w.seg_len, w.soff, w.partial);
            k_sgns_combine<SG_K><<<cgrid, 256, 0, s>>>(w.soff, V, D, w.partial, w.gO, w.gI);
        }
        DAISY_LAUNCH_CHECK(h);
        h->launches += 3;
    } else {  // optimizer.step() on an empty batch: zero gradients, the moments still decay
        DAISY_CUDA(cudaMemsetAsync(w.gI, 0, nvd * sizeof(float), s));
        DAISY_CUDA(cudaMemsetAsync(w.gO, 0, nvd * sizeof(float), s));
    }
    SgAdam a;
    a.b1 = p->beta1;
    a.b2 = p->beta2;
    a.eps = p->eps;
    a.step_size = (float)((double)p->lr / (1.0 - pow((double)p->beta1, (double)step_no)));
    a.inv_sqrt_bc2 = (float)(1.0 / sqrt(1.0 - pow((double)p->beta2, (double)step_no)));
    const size_t want = (2 * nvd + 255) / 256;
    const int grid = (int)(want < (size_t)h->num_sms * 16 ? want : (size_t)h->num_sms * 16);
    k_sgns_adam<<<grid > 0 ? grid : 1, 256, 0, s>>>(p->iv, p->m_iv, p->v_iv, w.gI, p->ov, p->m_ov, p->v_ov, w.gO, nvd, a);
    DAISY_LAUNCH_CHECK(h);
    if (loss_accum && B > 0) {
        k_sgns_loss<<<1, 256, 0, s>>>(w.lossp, Bi, 1.0 / ((double)B * (double)C), loss_accum);
        DAISY_LAUNCH_CHECK(h);
    }
    return DAISY_OK;
}
