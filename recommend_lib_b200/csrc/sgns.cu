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
//   k_sgns_rows     one block per vocabulary row and table: its segment of the sorted list is cut into 8 fixed slices
//                   (one per warp), the slice sums are added in slice order -> dense gradient rows, fixed order
//                   => bit-reproducible; the padding row (nn.Embedding(padding_idx=0), :40-41) gets a zero gradient
//   k_sgns_adam     torch.optim.Adam is dense: every element of both tables is stepped
#include <cub/cub.cuh>

#include "ctx.cuh"

namespace {

constexpr int SG_MAX_D = 512;            // 16 lane-strided values per lane
constexpr int SG_K = SG_MAX_D / 32;

struct SgScratch {
    float *coef, *ci, *lossp, *gI, *gO;
    uint32_t *kin, *kout, *vin, *vout, *ikin, *ikout, *ivin, *ivout;
    void *cub;
    size_t cub_bytes, total;
};

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
    uint32_t ids = 0;  // lane l holds the (validated) row of ref r0 + l of the current group of 32 refs
    for (int r = 0; r < R; ++r) {
        if ((r & 31) == 0) {  // one coalesced id load per 32 refs instead of a dependent scalar load per ref
            const int rr = r + lane;
            ids = rr < R ? sg_row(rr < C ? owords[(size_t)b * C + rr] : nwords[(size_t)b * NR + (rr - C)], V, b, err) : 0u;
        }
        const uint32_t row = __shfl_sync(0xffffffffu, ids, r & 31);
        const float *o = ov + (size_t)row * D;
        float orow[K];
        float dot = 0.f;
#pragma unroll
        for (int k = 0; k < K; ++k) {
            const int f = lane + 32 * k;
            orow[k] = f < D ? o[f] : 0.f;
            dot += ir[k] * orow[k];
        }
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

// grid (V, 2): blockIdx.y = 0: gradient of ovectors[row] = sum over its refs of coef * ivectors[iword_b] (pre-step);
//              blockIdx.y = 1: gradient of ivectors[row] = sum over the examples centred on it of their ci rows.
template <int K>
__global__ void __launch_bounds__(256) k_sgns_rows(const uint32_t *__restrict__ kout, const uint32_t *__restrict__ vout, int n,
                                                    const uint32_t *__restrict__ ikout, const uint32_t *__restrict__ ivout, int B,
                                                    int R, const float *__restrict__ coef, const float *__restrict__ ci,
                                                    const float *__restrict__ iv, const uint32_t *__restrict__ ikin, int D,
                                                    int padding_idx, float *__restrict__ gO, float *__restrict__ gI) {
    __shared__ float part[8][32 * K];
    __shared__ int seg[2];
    const int lane = threadIdx.x & 31, wv = threadIdx.x >> 5;
    const uint32_t row = blockIdx.x;
    const bool otab = blockIdx.y == 0;
    const uint32_t *keys = otab ? kout : ikout;
    const int len_all = otab ? n : B;
    if (threadIdx.x == 0) seg[0] = sg_lower_bound(keys, len_all, row);
    if (threadIdx.x == 32) seg[1] = sg_lower_bound(keys, len_all, row + 1u);
    __syncthreads();
    const int lo = seg[0], hi = seg[1], len = hi - lo;
    const int per = (len + 7) / 8;
    const int q0 = lo + wv * per, q1 = min(hi, q0 + per);
    float acc[K];
#pragma unroll
    for (int k = 0; k < K; ++k) acc[k] = 0.f;
#pragma unroll 4  // independent index -> row chains: let the scheduler put several rows in flight
    for (int q = q0; q < q1; ++q) {
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
#pragma unroll
    for (int k = 0; k < K; ++k) {
        const int f = lane + 32 * k;
        if (f < D) part[wv][f] = acc[k];
    }
    __syncthreads();
    float *g = (otab ? gO : gI) + (size_t)row * D;
    const bool pad = (int)row == padding_idx;
    for (int f = threadIdx.x; f < D; f += 256) {
        float t = 0.f;
#pragma unroll
        for (int s = 0; s < 8; ++s) t += part[s][f];  // slice order
        g[f] = pad ? 0.f : t;
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
        if (D <= 128)
            k_sgns_rows<4><<<dim3(V, 2), 256, 0, s>>>(w.kout, w.vout, n, w.ikout, w.ivout, Bi, (int)R, w.coef, w.ci, p->iv, w.ikin,
                                                      D, p->padding_idx, w.gO, w.gI);
        else if (D <= 320)
            k_sgns_rows<10><<<dim3(V, 2), 256, 0, s>>>(w.kout, w.vout, n, w.ikout, w.ivout, Bi, (int)R, w.coef, w.ci, p->iv, w.ikin,
                                                       D, p->padding_idx, w.gO, w.gI);
        else
            k_sgns_rows<SG_K><<<dim3(V, 2), 256, 0, s>>>(w.kout, w.vout, n, w.ikout, w.ivout, Bi, (int)R, w.coef, w.ci, p->iv,
                                                         w.ikin, D, p->padding_idx, w.gO, w.gI);
        DAISY_LAUNCH_CHECK(h);
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
