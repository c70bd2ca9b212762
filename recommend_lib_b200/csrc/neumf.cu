// NCF with an MLP tower: model 'MLP' and the script's default 'NeuMF-end' (NCFRecommender.py:28-125, 175-178) with
// dropout 0 (the script's default), BCEWithLogitsLoss (mean) and dense torch Adam (:255-260, 283-287).
// SURVEY.md section 8f, row N3; the GMF variant is csrc/gmf.cu.
//
// STATUS: GPU-verified in round 2 (tests/test_neumf_gpu.py passes on a B200 and runs in the default -m gpu suite; bench line under
// profiles/r02a_bench_neumf.json).  Before that it had run under the host emulation of tests/emu
// (tests/test_kernel_emulation.py: golden run + oracle green), which still covers it on CPU.
// The checker is pinned to the unmodified reference: the NeuMF checker under oracle/.
//
// First version, plain kernels, every reduction in a fixed order (bit-reproducible):
//   k_nm_sample   one block (4 warps) per sample: gathers, the MLP tower forward in shared memory (a warp per output
//                 neuron, coalesced weight rows), predict layer, BCE, and the backward pass down to the embedding rows;
//                 layer inputs and pre-activation gradients go to scratch for the weight gradients
//   k_nm_wgrad    dW_l[r, c] = sum_b dz_l[b, r] * in_l[b, c] (one thread per entry, samples in order), db_l, dwp, dbp
//   cub sort x 2  (user, sample) and (item, sample), stable
//   k_nm_rows     one block per table row: its samples' contributions summed in sample order -> dense gradient rows
//   k_nm_adam     torch.optim.Adam over every parameter that has a gradient (one launch, a job per tensor)
#include <cub/cub.cuh>

#include "ctx.cuh"

namespace {

constexpr int NM_MAX_IN = 1024;  // widest layer input (= 2 * MLP embedding size) the shared-memory tower holds
constexpr int NM_T = 128;

struct NmDims {
    int L, F, Dm, P, neumf;
    int in[DAISY_NEUMF_MAX_LAYERS], out[DAISY_NEUMF_MAX_LAYERS];
    int act_off[DAISY_NEUMF_MAX_LAYERS], dz_off[DAISY_NEUMF_MAX_LAYERS];  // offsets inside a sample's scratch rows
    int act_w, dz_w;                                                      // floats per sample
};

static void nm_dims(const daisy_neumf_params *p, NmDims &d) {
    d.L = p->num_layers;
    d.F = p->factor;
    d.neumf = p->neumf ? 1 : 0;
    d.Dm = p->factor << (p->num_layers - 1);
    d.P = d.neumf ? 2 * d.F : d.F;
    int in = 2 * d.Dm, ao = 0, zo = 0;
    for (int l = 0; l < d.L; ++l) {
        d.in[l] = in;
        d.out[l] = in / 2;
        d.act_off[l] = ao;
        d.dz_off[l] = zo;
        ao += in;
        zo += in / 2;
        in /= 2;
    }
    d.act_w = ao;
    d.dz_w = zo;
}

struct NmScratch {
    float *acts, *dz, *concat, *dx, *lossp, *cg_u, *cg_i, *cm_u, *cm_i;
    float *gPg, *gQg, *gPm, *gQm, *gW[DAISY_NEUMF_MAX_LAYERS], *gb[DAISY_NEUMF_MAX_LAYERS], *gwp, *gbp;
    uint32_t *ukin, *ukout, *uvin, *uvout, *ikin, *ikout, *ivin, *ivout;
    void *cub;
    size_t cub_bytes, total;
};

static size_t nm_sort_bytes(int64_t n) { return (size_t)n * 16 + ((size_t)1 << 20); }  // bound, see csrc/fmbn.cu

static void nm_carve(char *base, int64_t B, const daisy_neumf_params *p, const NmDims &d, NmScratch &w) {
    size_t off = 0;
    auto take = [&](size_t bytes) {
        char *q = base ? base + off : nullptr;
        off += (bytes + 255) / 256 * 256;
        return q;
    };
    const size_t b = (size_t)B, U = (size_t)p->user_num, I = (size_t)p->item_num;
    w.acts = (float *)take(b * d.act_w * 4);
    w.dz = (float *)take(b * d.dz_w * 4);
    w.concat = (float *)take(b * d.P * 4);
    w.dx = (float *)take(b * 4);
    w.lossp = (float *)take(b * 4);
    w.cg_u = (float *)take(b * d.F * 4);
    w.cg_i = (float *)take(b * d.F * 4);
    w.cm_u = (float *)take(b * d.Dm * 4);
    w.cm_i = (float *)take(b * d.Dm * 4);
    w.gPg = (float *)take(U * d.F * 4);
    w.gQg = (float *)take(I * d.F * 4);
    w.gPm = (float *)take(U * d.Dm * 4);
    w.gQm = (float *)take(I * d.Dm * 4);
    for (int l = 0; l < d.L; ++l) {
        w.gW[l] = (float *)take((size_t)d.out[l] * d.in[l] * 4);
        w.gb[l] = (float *)take((size_t)d.out[l] * 4);
    }
    w.gwp = (float *)take((size_t)d.P * 4);
    w.gbp = (float *)take(4);
    w.ukin = (uint32_t *)take(b * 4);
    w.ukout = (uint32_t *)take(b * 4);
    w.uvin = (uint32_t *)take(b * 4);
    w.uvout = (uint32_t *)take(b * 4);
    w.ikin = (uint32_t *)take(b * 4);
    w.ikout = (uint32_t *)take(b * 4);
    w.ivin = (uint32_t *)take(b * 4);
    w.ivout = (uint32_t *)take(b * 4);
    w.cub_bytes = nm_sort_bytes(B);
    w.cub = take(w.cub_bytes);
    w.total = off;
}

struct NmTower {  // what k_nm_sample needs, by value
    NmDims d;
    const float *Pg, *Qg, *Pm, *Qm, *W[DAISY_NEUMF_MAX_LAYERS], *b[DAISY_NEUMF_MAX_LAYERS], *wp, *bp;
};

// one block of 4 warps per sample.  train = 0: forward only (logits out).
__global__ void __launch_bounds__(NM_T) k_nm_sample(NmTower t, const int32_t *__restrict__ samples, int B, uint32_t U,
                                                     uint32_t I, int train, float inv_b, float *__restrict__ logits,
                                                     float *__restrict__ acts, float *__restrict__ dzs,
                                                     float *__restrict__ concat_out, float *__restrict__ dx_out,
                                                     float *__restrict__ lossp, float *__restrict__ cg_u,
                                                     float *__restrict__ cg_i, float *__restrict__ cm_u,
                                                     float *__restrict__ cm_i, uint32_t *__restrict__ ukin,
                                                     uint32_t *__restrict__ uvin, uint32_t *__restrict__ ikin,
                                                     uint32_t *__restrict__ ivin, int *err) {
    __shared__ float act[2 * NM_MAX_IN];   // inputs of every layer, then the tower output: in_0 + in_0/2 + ... + out_L-1
    __shared__ float grad[2][NM_MAX_IN];   // ping-pong: gradient w.r.t. a layer's input / pre-activation gradient
    __shared__ float cc[2 * NM_MAX_IN / 2];  // concat (P <= 2 F <= in_0 / 2 ... bounded by NM_MAX_IN)
    __shared__ float sdx;
    const NmDims &d = t.d;
    const int tid = threadIdx.x, lane = tid & 31, wv = tid >> 5;
    const int b = blockIdx.x;
    if (b >= B) return;
    uint32_t u = (uint32_t)samples[3 * (size_t)b], i = (uint32_t)samples[3 * (size_t)b + 1];
    const float y = (float)samples[3 * (size_t)b + 2];
    if (u >= U || i >= I) {  // never fault: park on row 0, the error flag tells the caller
        if (tid == 0) {
            atomicOr(&err[0], 1);
            atomicMin(&err[1], b);
        }
        u = u < U ? u : 0u;
        i = i < I ? i : 0u;
    }
    const int Dm = d.Dm, F = d.F, L = d.L;
    for (int c = tid; c < 2 * Dm; c += NM_T) act[c] = c < Dm ? t.Pm[(size_t)u * Dm + c] : t.Qm[(size_t)i * Dm + (c - Dm)];
    __syncthreads();
    int aoff = 0;
    for (int l = 0; l < L; ++l) {  // Linear + ReLU; a warp per output neuron, lanes over the (coalesced) weight row
        const int in = d.in[l], out = d.out[l];
        const float *W = t.W[l], *bias = t.b[l], *x = act + aoff;
        float *h = act + aoff + in;
        for (int r = wv; r < out; r += NM_T / 32) {
            float s = 0.f;
            for (int c = lane; c < in; c += 32) s += W[(size_t)r * in + c] * x[c];
            s = warp_sum(s);
            if (lane == 0) h[r] = fmaxf(s + bias[r], 0.f);
        }
        aoff += in;
        __syncthreads();
    }
    const float *hL = act + aoff;  // tower output, F values
    const int P = d.P;
    for (int f = tid; f < P; f += NM_T) {
        float v;
        if (d.neumf) v = f < F ? t.Pg[(size_t)u * F + f] * t.Qg[(size_t)i * F + f] : hL[f - F];
        else v = hL[f];
        cc[f] = v;
    }
    __syncthreads();
    if (wv == 0) {
        float s = 0.f;
        for (int f = lane; f < P; f += 32) s += t.wp[f] * cc[f];
        s = warp_sum(s);
        if (lane == 0) {
            const float x = s + t.bp[0];
            if (logits) logits[b] = x;
            if (train) {
                lossp[b] = fmaxf(x, 0.f) - x * y + log1pf(expf(-fabsf(x)));
                const float dx = (1.f / (1.f + expf(-x)) - y) * inv_b;
                dx_out[b] = dx;
                sdx = dx;
                ukin[b] = u;
                uvin[b] = (uint32_t)b;
                ikin[b] = i;
                ivin[b] = (uint32_t)b;
            }
        }
    }
    if (!train) return;
    __syncthreads();
    const float dx = sdx;
    // what the weight gradients need: the layer inputs and the concat vector
    for (int c = tid; c < d.act_w; c += NM_T) acts[(size_t)b * d.act_w + c] = act[c];
    for (int f = tid; f < P; f += NM_T) concat_out[(size_t)b * P + f] = cc[f];
    // d concat = dx * wp; GMF branch contributions; pre-activation gradient of the last layer
    const int moff = d.neumf ? F : 0;
    if (d.neumf)
        for (int f = tid; f < F; f += NM_T) {
            const float g = dx * t.wp[f];
            cg_u[(size_t)b * F + f] = g * t.Qg[(size_t)i * F + f];
            cg_i[(size_t)b * F + f] = g * t.Pg[(size_t)u * F + f];
        }
    int cur = 0;
    for (int r = tid; r < F; r += NM_T) grad[cur][r] = hL[r] > 0.f ? dx * t.wp[moff + r] : 0.f;  // dz of layer L-1
    __syncthreads();
    for (int l = L - 1; l >= 0; --l) {
        const int in = d.in[l], out = d.out[l];
        const float *W = t.W[l], *dz = grad[cur];
        for (int r = tid; r < out; r += NM_T) dzs[(size_t)b * d.dz_w + d.dz_off[l] + r] = dz[r];
        const float *x = act + d.act_off[l];  // this layer's input (= the previous layer's ReLU output for l > 0)
        float *dn = grad[cur ^ 1];
        for (int c = tid; c < in; c += NM_T) {  // d input[c] = sum_r dz[r] W[r, c]: coalesced over c for every r
            float s = 0.f;
            for (int r = 0; r < out; ++r) s += dz[r] * W[(size_t)r * in + c];
            dn[c] = (l > 0) ? (x[c] > 0.f ? s : 0.f) : s;  // through the previous ReLU, or the raw embedding gradient
        }
        cur ^= 1;
        __syncthreads();
    }
    const float *d0 = grad[cur];
    for (int c = tid; c < Dm; c += NM_T) {
        cm_u[(size_t)b * Dm + c] = d0[c];
        cm_i[(size_t)b * Dm + c] = d0[Dm + c];
    }
}

struct NmWJob {
    const float *dz, *in;  // dz[b * dz_w + off_r + r], in[b * in_w + off_c + c]
    float *gW, *gb;
    int out, inw, dz_w, in_w, off_r, off_c;
};
struct NmWJobs {
    NmWJob j[DAISY_NEUMF_MAX_LAYERS + 1];
    int count;
};

// grid (ceil(max entries / 256), jobs): one thread per weight entry, samples in order; thread (r, 0) also sums the bias
__global__ void __launch_bounds__(256) k_nm_wgrad(NmWJobs jobs, int B) {
    const NmWJob &j = jobs.j[blockIdx.y];
    const size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= (size_t)j.out * j.inw) return;
    const int r = (int)(e / j.inw), c = (int)(e % j.inw);
    float s = 0.f, sb = 0.f;
    for (int b = 0; b < B; ++b) {
        const float g = j.dz[(size_t)b * j.dz_w + j.off_r + r];
        s += g * j.in[(size_t)b * j.in_w + j.off_c + c];
        sb += g;
    }
    j.gW[e] = s;
    if (c == 0 && j.gb) j.gb[r] = sb;
}

__device__ __forceinline__ int nm_lower_bound(const uint32_t *__restrict__ a, int n, uint32_t key) {
    int lo = 0, hi = n;
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (a[mid] < key) lo = mid + 1; else hi = mid;
    }
    return lo;
}

// grid (max(U, I), 2): side 0 = user rows, side 1 = item rows: dense gradient rows of the MLP (and GMF) tables
__global__ void __launch_bounds__(NM_T) k_nm_rows(const uint32_t *__restrict__ ukout, const uint32_t *__restrict__ uvout,
                                                   const uint32_t *__restrict__ ikout, const uint32_t *__restrict__ ivout, int B,
                                                   uint32_t U, uint32_t I, int Dm, int F, int neumf,
                                                   const float *__restrict__ cm_u, const float *__restrict__ cm_i,
                                                   const float *__restrict__ cg_u, const float *__restrict__ cg_i,
                                                   float *__restrict__ gPm, float *__restrict__ gQm, float *__restrict__ gPg,
                                                   float *__restrict__ gQg) {
    __shared__ int seg[2];
    const uint32_t row = blockIdx.x;
    const int side = blockIdx.y;
    if (row >= (side ? I : U)) return;
    const uint32_t *keys = side ? ikout : ukout, *vals = side ? ivout : uvout;
    if (threadIdx.x == 0) seg[0] = nm_lower_bound(keys, B, row);
    if (threadIdx.x == 32) seg[1] = nm_lower_bound(keys, B, row + 1u);
    __syncthreads();
    const int lo = seg[0], hi = seg[1];
    const float *cm = side ? cm_i : cm_u, *cg = side ? cg_i : cg_u;
    float *gm = (side ? gQm : gPm) + (size_t)row * Dm;
    for (int c = threadIdx.x; c < Dm; c += NM_T) {
        float s = 0.f;
        for (int q = lo; q < hi; ++q) s += cm[(size_t)vals[q] * Dm + c];  // sorted = sample order inside a row
        gm[c] = s;
    }
    if (neumf) {
        float *gg = (side ? gQg : gPg) + (size_t)row * F;
        for (int c = threadIdx.x; c < F; c += NM_T) {
            float s = 0.f;
            for (int q = lo; q < hi; ++q) s += cg[(size_t)vals[q] * F + c];
            gg[c] = s;
        }
    }
}

struct NmAdam {
    float b1, b2, step_size, inv_sqrt_bc2, eps;
};
struct NmAJob {
    float *th, *m, *v;
    const float *g;
    size_t n;
};
struct NmAJobs {
    NmAJob j[2 * DAISY_NEUMF_MAX_LAYERS + 6];
    int count;
};

__global__ void k_nm_adam(NmAJobs jobs, NmAdam a) {
    const NmAJob &j = jobs.j[blockIdx.y];
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < j.n; e += stride) {
        const float g = j.g[e];
        const float m = a.b1 * j.m[e] + (1.f - a.b1) * g;
        const float v = a.b2 * j.v[e] + (1.f - a.b2) * g * g;
        j.m[e] = m;
        j.v[e] = v;
        j.th[e] -= a.step_size * m / (sqrtf(v) * a.inv_sqrt_bc2 + a.eps);
    }
}

__global__ void __launch_bounds__(256) k_nm_loss(const float *__restrict__ lossp, int B, double scale, double *loss_accum) {
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

static int nm_check(daisy_ctx *h, const daisy_neumf_params *p, const void *samples, int64_t B) {
    DAISY_REQUIRE(h != nullptr && p != nullptr, DAISY_EINVAL, "null handle or parameter block");
    DAISY_REQUIRE(p->num_layers >= 1 && p->num_layers <= DAISY_NEUMF_MAX_LAYERS && p->factor >= 1, DAISY_EINVAL,
                  "num_layers %d / factor_num %d out of range", p->num_layers, p->factor);
    DAISY_REQUIRE(((int64_t)p->factor << p->num_layers) <= NM_MAX_IN, DAISY_EUNSUPPORTED,
                  "MLP input of %lld values (factor_num * 2^num_layers) exceeds %d", (long long)p->factor << p->num_layers, NM_MAX_IN);
    DAISY_REQUIRE(p->user_num > 0 && p->item_num > 0 && p->user_num < 0x7fffffffLL && p->item_num < 0x7fffffffLL, DAISY_EINVAL,
                  "bad table sizes");
    DAISY_REQUIRE(p->Pm && p->Qm && p->wp && p->bp && (!p->neumf || (p->Pg && p->Qg)), DAISY_EINVAL, "null table / predict pointer");
    for (int l = 0; l < p->num_layers; ++l) DAISY_REQUIRE(p->W[l] && p->b[l], DAISY_EINVAL, "null MLP layer %d", l);
    DAISY_REQUIRE(B >= 0 && B < 0x7fffffffLL / 2048, DAISY_EINVAL, "batch %lld out of range", (long long)B);
    DAISY_REQUIRE(B == 0 || samples != nullptr, DAISY_EINVAL, "null samples");
    return DAISY_OK;
}

static void nm_tower(const daisy_neumf_params *p, const NmDims &d, NmTower &t) {
    t.d = d;
    t.Pg = p->Pg; t.Qg = p->Qg; t.Pm = p->Pm; t.Qm = p->Qm; t.wp = p->wp; t.bp = p->bp;
    for (int l = 0; l < DAISY_NEUMF_MAX_LAYERS; ++l) {
        t.W[l] = l < d.L ? p->W[l] : nullptr;
        t.b[l] = l < d.L ? p->b[l] : nullptr;
    }
}

}  // namespace

extern "C" int daisy_neumf_scratch_bytes(const daisy_neumf_params *p, int64_t B, int64_t *bytes) {
    DAISY_REQUIRE(p != nullptr && bytes != nullptr && B >= 0 && p->num_layers >= 1 && p->num_layers <= DAISY_NEUMF_MAX_LAYERS &&
                      p->factor >= 1 && p->user_num > 0 && p->item_num > 0,
                  DAISY_EINVAL, "bad scratch query");
    NmDims d;
    nm_dims(p, d);
    NmScratch w;
    nm_carve(nullptr, B > 0 ? B : 1, p, d, w);
    *bytes = (int64_t)w.total;
    return DAISY_OK;
}

extern "C" int daisy_neumf_forward(daisy_handle_t h, const daisy_neumf_params *p, const int32_t *samples, int64_t B,
                                   float *logits, daisy_stream_t stream) {
    int rc = nm_check(h, p, samples, B);
    if (rc) return rc;
    if (B == 0) return DAISY_OK;
    DAISY_REQUIRE(logits != nullptr, DAISY_EINVAL, "null output");
    DeviceGuard g(h->device);
    DAISY_REQUIRE(g.ok, DAISY_ECUDA, "cannot select device %d", h->device);
    NmDims d;
    nm_dims(p, d);
    NmTower t;
    nm_tower(p, d, t);
    k_nm_sample<<<(int)B, NM_T, 0, (cudaStream_t)stream>>>(t, samples, (int)B, (uint32_t)p->user_num, (uint32_t)p->item_num, 0, 0.f,
                                                           logits, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr,
                                                           nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, h->err);
    DAISY_LAUNCH_CHECK(h);
    return DAISY_OK;
}

extern "C" int daisy_neumf_step(daisy_handle_t h, const daisy_neumf_params *p, const int32_t *samples, int64_t B,
                                int64_t step_no, void *scratch, int64_t scratch_bytes, double *loss_accum,
                                daisy_stream_t stream) {
    int rc = nm_check(h, p, samples, B);
    if (rc) return rc;
    DAISY_REQUIRE(step_no >= 1, DAISY_EINVAL, "step_no is 1-based");
    DAISY_REQUIRE(p->m_Pm && p->v_Pm && p->m_Qm && p->v_Qm && p->m_wp && p->v_wp && p->m_bp && p->v_bp &&
                      (!p->neumf || (p->m_Pg && p->v_Pg && p->m_Qg && p->v_Qg)),
                  DAISY_EINVAL, "null Adam moment pointer");
    for (int l = 0; l < p->num_layers; ++l)
        DAISY_REQUIRE(p->m_W[l] && p->v_W[l] && p->m_b[l] && p->v_b[l], DAISY_EINVAL, "null Adam moment pointer of MLP layer %d", l);
    if (B == 0) return DAISY_OK;  // optimizer.step() without gradients: Adam skips parameters whose .grad is None
    DAISY_REQUIRE((uintptr_t)scratch % 256 == 0 && scratch != nullptr, DAISY_EINVAL, "scratch must be 256-byte aligned");
    NmDims d;
    nm_dims(p, d);
    NmScratch w;
    nm_carve((char *)scratch, B, p, d, w);
    DAISY_REQUIRE((int64_t)w.total <= scratch_bytes, DAISY_EINVAL, "scratch of %lld bytes, %zu needed (daisy_neumf_scratch_bytes)",
                  (long long)scratch_bytes, w.total);
    DeviceGuard g(h->device);
    DAISY_REQUIRE(g.ok, DAISY_ECUDA, "cannot select device %d", h->device);
    cudaStream_t s = (cudaStream_t)stream;
    const int Bi = (int)B;
    const uint32_t U = (uint32_t)p->user_num, I = (uint32_t)p->item_num;
    NmTower t;
    nm_tower(p, d, t);
    k_nm_sample<<<Bi, NM_T, 0, s>>>(t, samples, Bi, U, I, 1, 1.0f / (float)B, nullptr, w.acts, w.dz, w.concat, w.dx, w.lossp,
                                    w.cg_u, w.cg_i, w.cm_u, w.cm_i, w.ukin, w.uvin, w.ikin, w.ivin, h->err);
    DAISY_LAUNCH_CHECK(h);
    NmWJobs wj;
    size_t max_entries = 0;
    for (int l = 0; l < d.L; ++l) {
        wj.j[l] = NmWJob{w.dz, w.acts, w.gW[l], w.gb[l], d.out[l], d.in[l], d.dz_w, d.act_w, d.dz_off[l], d.act_off[l]};
        if ((size_t)d.out[l] * d.in[l] > max_entries) max_entries = (size_t)d.out[l] * d.in[l];
    }
    // predict layer as a 1 x P "layer": dz = dx (one value per sample), input = concat
    wj.j[d.L] = NmWJob{w.dx, w.concat, w.gwp, w.gbp, 1, d.P, 1, d.P, 0, 0};
    if ((size_t)d.P > max_entries) max_entries = (size_t)d.P;
    wj.count = d.L + 1;
    k_nm_wgrad<<<dim3(daisy_ceil_div((int64_t)max_entries, 256), wj.count), 256, 0, s>>>(wj, Bi);
    DAISY_LAUNCH_CHECK(h);
    int ubits = 1, ibits = 1;
    while (ubits < 32 && (1ll << ubits) < p->user_num) ++ubits;
    while (ibits < 32 && (1ll << ibits) < p->item_num) ++ibits;
    size_t cub_bytes = w.cub_bytes;
    DAISY_CUDA(cub::DeviceRadixSort::SortPairs(w.cub, cub_bytes, w.ukin, w.ukout, w.uvin, w.uvout, Bi, 0, ubits, s));
    cub_bytes = w.cub_bytes;
    DAISY_CUDA(cub::DeviceRadixSort::SortPairs(w.cub, cub_bytes, w.ikin, w.ikout, w.ivin, w.ivout, Bi, 0, ibits, s));
    h->launches += 2;
    const uint32_t rows = U > I ? U : I;
    k_nm_rows<<<dim3(rows, 2), NM_T, 0, s>>>(w.ukout, w.uvout, w.ikout, w.ivout, Bi, U, I, d.Dm, d.F, d.neumf, w.cm_u, w.cm_i,
                                             w.cg_u, w.cg_i, w.gPm, w.gQm, w.gPg, w.gQg);
    DAISY_LAUNCH_CHECK(h);
    NmAJobs aj;
    int n = 0;
    size_t max_n = 0;
    auto add = [&](float *th, float *m, float *v, const float *gr, size_t cnt) {
        aj.j[n++] = NmAJob{th, m, v, gr, cnt};
        if (cnt > max_n) max_n = cnt;
    };
    if (d.neumf) {
        add(p->Pg, p->m_Pg, p->v_Pg, w.gPg, (size_t)U * d.F);
        add(p->Qg, p->m_Qg, p->v_Qg, w.gQg, (size_t)I * d.F);
    }
    add(p->Pm, p->m_Pm, p->v_Pm, w.gPm, (size_t)U * d.Dm);
    add(p->Qm, p->m_Qm, p->v_Qm, w.gQm, (size_t)I * d.Dm);
    for (int l = 0; l < d.L; ++l) {
        add(p->W[l], p->m_W[l], p->v_W[l], w.gW[l], (size_t)d.out[l] * d.in[l]);
        add(p->b[l], p->m_b[l], p->v_b[l], w.gb[l], (size_t)d.out[l]);
    }
    add(p->wp, p->m_wp, p->v_wp, w.gwp, (size_t)d.P);
    add(p->bp, p->m_bp, p->v_bp, w.gbp, 1);
    aj.count = n;
    NmAdam a;
    a.b1 = p->beta1;
    a.b2 = p->beta2;
    a.eps = p->eps;
    a.step_size = (float)((double)p->lr / (1.0 - pow((double)p->beta1, (double)step_no)));
    a.inv_sqrt_bc2 = (float)(1.0 / sqrt(1.0 - pow((double)p->beta2, (double)step_no)));
    const size_t want = (max_n + 255) / 256;
    const int gx = (int)(want < (size_t)h->num_sms * 8 ? want : (size_t)h->num_sms * 8);
    k_nm_adam<<<dim3(gx > 0 ? gx : 1, n), 256, 0, s>>>(aj, a);
    DAISY_LAUNCH_CHECK(h);
    if (loss_accum) {
        k_nm_loss<<<1, 256, 0, s>>>(w.lossp, Bi, 1.0 / (double)B, loss_accum);
        DAISY_LAUNCH_CHECK(h);
    }
    return DAISY_OK;
}
