// BPR-MF training step for sm_100a: daisy_bpr_step / daisy_bpr_step_host / daisy_bpr_adam_step.
//
// Replaces (reference, file:line): zero_grad + BPR.forward + loss + backward + optim.SGD.step,
// BPRMFRecommender.py:42-50,154,172-176 -- i.e. ATen embedding gathers, embedding_dense_backward and the
// dense _foreach SGD pass.
//
// Pipeline of one step (B triples, tables P [U,D], Q [I,D], all gradients at the PRE-step tables):
//
//   prep     validate ids; key = positive item, value = triple id
//   sort_i   radix sort (CUB) -> triples grouped by positive item ("sorted triple" k)
//   refs     gather triples into sorted order; emit user refs (u, k), negative-item refs (j, k) and one
//            positive-item ref per RUN HEAD (a run = consecutive sorted triples with the same positive item
//            inside one warp chunk of C triples); non-heads get a sentinel key that sorts last
//   sort_u / sort_q   radix sort refs by row -> every row's contributions are contiguous, in a fixed order
//   slots    a row referenced once in the whole batch is DIRECT (updated in place by the main kernel);
//            every other contribution gets a staging slot = its position in the sorted ref array
//   main     one warp per chunk of C sorted triples: 128-bit row gathers (next triple prefetched), warp
//            shuffle dot, sigmoid coefficient in registers, positive-item gradient accumulated in registers
//            over the run; DIRECT rows are written back in place (read once, written once), the others
//            write their contribution to the staging slot
//   seg_u / seg_q   one warp per window of 32 sorted refs: for every multi-contribution row, sum the staged
//            contributions in sorted order (contiguous, streaming reads) and update the row once
//   heavy    rows with more contributions than `heavy_len` are reduced in two levels: one warp per slice of 64
//            contributions, then one block per row over the slice partials
//   loss     fixed-order reduction of the per-warp loss partials into *loss_accum
//
// No float atomics anywhere: the result is bit-reproducible run to run.
// The L2 weight decay of optim.SGD (every row shrinks by 1 - lr*wd per step) is carried by the scalar c
// (ctx.cuh: scale): stored tables are W/c, scores use c^2, the row update is
//   W_hat += [lr / (1 - lr*wd)] * s * (...)          (derivation in DESIGN.md)
// so untouched rows cost no traffic.
#include <cub/device/device_radix_sort.cuh>

#include "ctx.cuh"

namespace {

// ------------------------------------------------------------------------------------------------
// optimiser functors: apply(table id, f4 index, old row slice, descent direction d = -gradient)
// ------------------------------------------------------------------------------------------------
struct SgdOpt {
    float *P, *Q;
    float alpha;  // lr / (1 - lr*wd)
    __device__ __forceinline__ void apply(int tbl, size_t idx, float4 old, float4 d) const {
        float4 r = make_float4(fmaf(alpha, d.x, old.x), fmaf(alpha, d.y, old.y), fmaf(alpha, d.z, old.z),
                               fmaf(alpha, d.w, old.w));
        st_row(tbl ? Q : P, idx, r);
    }
};

// torch.optim.SparseAdam semantics on the rows present in the batch (lazy: other rows and their moments
// are untouched).  g = -d.
struct AdamOpt {
    float *P, *Q, *mP, *vP, *mQ, *vQ;
    float b1, b2, step_size, inv_sqrt_bc2, eps;
    __device__ __forceinline__ float one(float w, float g, float &m, float &v) const {
        m = fmaf(b1, m, (1.f - b1) * g);
        v = fmaf(b2, v, (1.f - b2) * g * g);
        return w - step_size * m / (sqrtf(v) * inv_sqrt_bc2 + eps);
    }
    __device__ __forceinline__ void apply(int tbl, size_t idx, float4 old, float4 d) const {
        float *M = tbl ? mQ : mP, *Vv = tbl ? vQ : vP;
        float4 m = ld_row(M, idx), v = ld_row(Vv, idx), r;
        r.x = one(old.x, -d.x, m.x, v.x);
        r.y = one(old.y, -d.y, m.y, v.y);
        r.z = one(old.z, -d.z, m.z, v.z);
        r.w = one(old.w, -d.w, m.w, v.w);
        st_row(M, idx, m);
        st_row(Vv, idx, v);
        st_row(tbl ? Q : P, idx, r);
    }
};

// Row-sharded step: users are local (SGD + lazy L2 as above); the "item table" is a cache of rows fetched
// from their owners, and instead of updating it the step emits the complete descent sum of every cache row -- the
// owners apply them (k_owner_apply).  Every cache row is referenced by at least one triple, so every row of G is
// written exactly once.
struct ShardOpt {
    float *P, *G;
    float alpha;
    __device__ __forceinline__ void apply(int tbl, size_t idx, float4 old, float4 d) const {
        if (tbl) {
            st_row(G, idx, d);
        } else {
            st_row(P, idx, make_float4(fmaf(alpha, d.x, old.x), fmaf(alpha, d.y, old.y), fmaf(alpha, d.z, old.z),
                                       fmaf(alpha, d.w, old.w)));
        }
    }
};

// ------------------------------------------------------------------------------------------------
// prep: validate, emit sort keys
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void load_triple(const int32_t *__restrict__ triples, int t, uint32_t U, uint32_t I,
                                            uint32_t &u, uint32_t &i, uint32_t &j, bool &bad) {
    u = (uint32_t)triples[3 * (size_t)t];
    i = (uint32_t)triples[3 * (size_t)t + 1];
    j = (uint32_t)triples[3 * (size_t)t + 2];
    bad = (u >= U) | (i >= I) | (j >= I);
    if (bad) {  // never fault: park the triple on row 0, the error flag tells the caller
        u = u < U ? u : 0u;
        i = i < I ? i : 0u;
        j = j < I ? j : 0u;
    }
}

__global__ void k_prep(const int32_t *__restrict__ triples, int B, uint32_t U, uint32_t I, uint32_t *__restrict__ key,
                       uint32_t *__restrict__ val, int *err) {
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= B) return;
    uint32_t u, i, j;
    bool bad;
    load_triple(triples, t, U, I, u, i, j, bad);
    if (bad) {
        atomicOr(&err[0], 1);
        atomicMin(&err[1], t);
    }
    key[t] = i;
    val[t] = (uint32_t)t;
}

// ------------------------------------------------------------------------------------------------
// refs: sorted triples + reference lists
// ------------------------------------------------------------------------------------------------
__global__ void k_refs(const int32_t *__restrict__ triples, const uint32_t *__restrict__ order,
                       const uint32_t *__restrict__ sorted_i, int B, uint32_t U, uint32_t I, int C,
                       int32_t *__restrict__ st, uint32_t *__restrict__ ukey, uint32_t *__restrict__ uval,
                       uint32_t *__restrict__ qkey, uint32_t *__restrict__ qval) {
    int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= B) return;
    uint32_t u, i, j;
    bool bad;
    load_triple(triples, (int)order[k], U, I, u, i, j, bad);
    st[3 * (size_t)k] = (int32_t)u;
    st[3 * (size_t)k + 1] = (int32_t)i;
    st[3 * (size_t)k + 2] = (int32_t)j;
    ukey[k] = u;
    uval[k] = (uint32_t)k;
    qkey[k] = j;  // negative-item ref of sorted triple k
    qval[k] = (uint32_t)k;
    const bool head = (k % C == 0) || (sorted_i[k - 1] != i);
    qkey[B + k] = head ? i : I;  // I = sentinel, sorts after every real row
    qval[B + k] = (uint32_t)(B + k);
}

// ------------------------------------------------------------------------------------------------
// slots: DIRECT (single contribution in the batch) or staging slot = sorted position
// ------------------------------------------------------------------------------------------------
__global__ void k_slots_user(const uint32_t *__restrict__ key, const uint32_t *__restrict__ val, int n,
                             uint32_t *__restrict__ uslot) {
    int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n) return;
    const uint32_t r = key[p];
    const bool first = (p == 0) || (key[p - 1] != r);
    const bool last = (p == n - 1) || (key[p + 1] != r);
    uslot[val[p]] = (first && last) ? DAISY_DIRECT : (uint32_t)p;
}

__global__ void k_slots_item(const uint32_t *__restrict__ key, const uint32_t *__restrict__ val, int n, int B,
                             uint32_t sentinel, uint32_t *__restrict__ jslot, uint32_t *__restrict__ islot) {
    int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n) return;
    const uint32_t r = key[p];
    if (r == sentinel) return;
    const bool first = (p == 0) || (key[p - 1] != r);
    const bool last = (p == n - 1) || (key[p + 1] != r);
    const uint32_t slot = (first && last) ? DAISY_DIRECT : (uint32_t)p;
    const uint32_t v = val[p];
    if (v < (uint32_t)B)
        jslot[v] = slot;
    else
        islot[v - B] = slot;
}

// ------------------------------------------------------------------------------------------------
// main fused kernel
// ------------------------------------------------------------------------------------------------
struct MainArgs {
    const float *P;
    const float *Q;
    const int32_t *st;
    const uint32_t *uslot, *jslot, *islot;
    float *stageU, *stageQ;
    float *loss_part;
    int B, D4, C;
    float c2;  // score scale = c^2
};

template <int V, class Opt>
__global__ void __launch_bounds__(256) k_bpr_main(MainArgs a, Opt opt) {
    const unsigned FULL = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    const int warp = (int)((blockIdx.x * (size_t)blockDim.x + threadIdx.x) >> 5);
    const int C = a.C;
    const long long k0 = (long long)warp * C;
    if (k0 >= a.B) return;  // warp-uniform
    const int n = (int)min((long long)C, (long long)a.B - k0);
    const int D4 = a.D4;

    // chunk metadata: lane l holds sorted triple k0 + l (C <= 32)
    int mu = 0, mi = 0, mj = 0;
    uint32_t mus = 0, mjs = 0, mis = 0;
    if (lane < n) {
        const size_t k = (size_t)(k0 + lane);
        mu = a.st[3 * k];
        mi = a.st[3 * k + 1];
        mj = a.st[3 * k + 2];
        mus = a.uslot[k];
        mjs = a.jslot[k];
        mis = a.islot[k];  // meaningful at run heads only
    }
    bool act[V];
#pragma unroll
    for (int v = 0; v < V; ++v) act[v] = (lane + 32 * v) < D4;

    float4 pu[V], qj[V], qi[V], acc[V], pu_n[V], qj_n[V], qi_n[V];
#pragma unroll
    for (int v = 0; v < V; ++v) pu[v] = qj[v] = qi[v] = acc[v] = pu_n[v] = qj_n[v] = qi_n[v] = f4_zero();
    {
        const int u = __shfl_sync(FULL, mu, 0), i = __shfl_sync(FULL, mi, 0), j = __shfl_sync(FULL, mj, 0);
#pragma unroll
        for (int v = 0; v < V; ++v)
            if (act[v]) {
                pu_n[v] = ld_row(a.P, (size_t)u * D4 + lane + 32 * v);
                qj_n[v] = ld_row(a.Q, (size_t)j * D4 + lane + 32 * v);
                qi_n[v] = ld_row(a.Q, (size_t)i * D4 + lane + 32 * v);
            }
    }
    int cur_i = -1;
    uint32_t cur_is = 0;
    float loss = 0.f;

    for (int t = 0; t < n; ++t) {
        const int u = __shfl_sync(FULL, mu, t), i = __shfl_sync(FULL, mi, t), j = __shfl_sync(FULL, mj, t);
        const uint32_t us = __shfl_sync(FULL, mus, t), js = __shfl_sync(FULL, mjs, t);
        const uint32_t is_t = __shfl_sync(FULL, mis, t);
        if (t == 0 || i != cur_i) {  // a new positive-item run starts here (warp-uniform)
            if (t > 0) {
#pragma unroll
                for (int v = 0; v < V; ++v)
                    if (act[v]) {
                        const int e = lane + 32 * v;
                        if (cur_is == DAISY_DIRECT)
                            opt.apply(1, (size_t)cur_i * D4 + e, qi[v], acc[v]);
                        else
                            st_stream(a.stageQ, (size_t)cur_is * D4 + e, acc[v]);
                    }
            }
            cur_i = i;
            cur_is = is_t;
#pragma unroll
            for (int v = 0; v < V; ++v) {
                qi[v] = qi_n[v];
                acc[v] = f4_zero();
            }
        }
#pragma unroll
        for (int v = 0; v < V; ++v) {
            pu[v] = pu_n[v];
            qj[v] = qj_n[v];
        }
        // prefetch the rows of sorted triple t+1 before touching triple t.  Safe: a row that is written in
        // place below is referenced exactly once in the whole batch, so no later triple reads it.
        if (t + 1 < n) {
            const int un = __shfl_sync(FULL, mu, t + 1), in = __shfl_sync(FULL, mi, t + 1),
                      jn = __shfl_sync(FULL, mj, t + 1);
#pragma unroll
            for (int v = 0; v < V; ++v)
                if (act[v]) {
                    pu_n[v] = ld_row(a.P, (size_t)un * D4 + lane + 32 * v);
                    qj_n[v] = ld_row(a.Q, (size_t)jn * D4 + lane + 32 * v);
                    if (in != i) qi_n[v] = ld_row(a.Q, (size_t)in * D4 + lane + 32 * v);
                }
        }
        // x = c^2 <P[u], Q[i] - Q[j]>;  s = sigmoid(-x) = -d(loss)/dx
        float d = 0.f;
#pragma unroll
        for (int v = 0; v < V; ++v)
            if (act[v]) d += f4_dot(pu[v], f4_sub(qi[v], qj[v]));
        d = warp_sum(d);
        const float x = d * a.c2;
        const float s = 1.f / (1.f + expf(x));
        loss += fmaxf(-x, 0.f) + log1pf(expf(-fabsf(x)));  // -log sigmoid(x), overflow-safe
#pragma unroll
        for (int v = 0; v < V; ++v)
            if (act[v]) {
                const int e = lane + 32 * v;
                // user row: descent direction +s (Q[i] - Q[j])
                const float4 gu = f4_scale(f4_sub(qi[v], qj[v]), s);
                if (us == DAISY_DIRECT)
                    opt.apply(0, (size_t)u * D4 + e, pu[v], gu);
                else
                    st_stream(a.stageU, (size_t)us * D4 + e, gu);
                // negative item row: descent direction -s P[u]
                const float4 gj = f4_scale(pu[v], -s);
                if (js == DAISY_DIRECT)
                    opt.apply(1, (size_t)j * D4 + e, qj[v], gj);
                else
                    st_stream(a.stageQ, (size_t)js * D4 + e, gj);
                // positive item row: +s P[u], accumulated over the run in sorted order
                acc[v].x = fmaf(s, pu[v].x, acc[v].x);
                acc[v].y = fmaf(s, pu[v].y, acc[v].y);
                acc[v].z = fmaf(s, pu[v].z, acc[v].z);
                acc[v].w = fmaf(s, pu[v].w, acc[v].w);
            }
    }
#pragma unroll
    for (int v = 0; v < V; ++v)
        if (act[v]) {
            const int e = lane + 32 * v;
            if (cur_is == DAISY_DIRECT)
                opt.apply(1, (size_t)cur_i * D4 + e, qi[v], acc[v]);
            else
                st_stream(a.stageQ, (size_t)cur_is * D4 + e, acc[v]);
        }
    if (lane == 0) a.loss_part[warp] = loss;
}

// ------------------------------------------------------------------------------------------------
// segmented reduce + row update for rows with several contributions
// ------------------------------------------------------------------------------------------------
template <int V>
__device__ __forceinline__ void sum_staged(const float *__restrict__ stage, size_t q0, int len, int D4, int lane,
                                           const bool (&act)[V], float4 (&acc)[V]) {
    constexpr int UN = (V == 1) ? 8 : (V == 2 ? 4 : 2);
    for (int c = 0; c < len; c += UN) {
        float4 r[UN][V];
#pragma unroll
        for (int jj = 0; jj < UN; ++jj)
#pragma unroll
            for (int v = 0; v < V; ++v)
                r[jj][v] = (c + jj < len && act[v]) ? ld_stream(stage, (q0 + c + jj) * D4 + lane + 32 * v) : f4_zero();
#pragma unroll
        for (int jj = 0; jj < UN; ++jj)  // fixed order: sorted position ascending
            if (c + jj < len) {
#pragma unroll
                for (int v = 0; v < V; ++v) acc[v] = f4_add(acc[v], r[jj][v]);
            }
    }
}

template <int V, class Opt>
__global__ void __launch_bounds__(256) k_seg_reduce(int tbl, const float *__restrict__ table,
                                                     const uint32_t *__restrict__ keys, int n, uint32_t sentinel,
                                                     const float *__restrict__ stage, int D4, Opt opt, int heavy_len,
                                                     uint32_t *heavy, int heavy_cap) {
    const unsigned FULL = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    const long long w = (long long)((blockIdx.x * (size_t)blockDim.x + threadIdx.x) >> 5);
    const long long base = w * 32;
    if (base >= n) return;
    const long long p = base + lane;
    const uint32_t key = (p < n) ? keys[p] : sentinel;
    uint32_t prev = __shfl_up_sync(FULL, key, 1);
    if (lane == 0) prev = (p > 0) ? keys[p - 1] : ~key;
    uint32_t next = __shfl_down_sync(FULL, key, 1);
    if (lane == 31) next = (p + 1 < n) ? keys[p + 1] : sentinel;
    const bool valid = (p < n) && (key != sentinel);
    const bool start = valid && (prev != key);
    const bool multi = start && (next == key) && (p + 1 < n);
    const unsigned boundary = __ballot_sync(FULL, start || !valid);
    unsigned todo = __ballot_sync(FULL, multi);
    bool act[V];
#pragma unroll
    for (int v = 0; v < V; ++v) act[v] = (lane + 32 * v) < D4;

    while (todo) {
        const int b = __ffs(todo) - 1;
        todo &= todo - 1;
        const uint32_t row = __shfl_sync(FULL, key, b);
        const unsigned after = (b == 31) ? 0u : (boundary & ~((2u << b) - 1u));
        int len;
        if (after) {
            len = (__ffs(after) - 1) - b;
        } else {  // the segment runs past this window: count matching keys in the following windows
            len = 32 - b;
            long long q = base + 32;
            while (true) {
                const bool ok = (q + lane < n) && (keys[q + lane] == row);
                const unsigned m = __ballot_sync(FULL, ok);
                const int c = (m == FULL) ? 32 : (__ffs(~m) - 1);
                len += c;
                if (c < 32 || len > heavy_len) break;
                q += 32;
            }
        }
        const size_t q0 = (size_t)(base + b);
        if (len > heavy_len) {  // very hot row: reduced in two levels by k_heavy_slices / k_heavy_final
            // exact length by a 32-ary search over the sorted keys: keys[q] == row for q in [q0, q0 + len)
            long long lo = (long long)q0 + len, hi = n;
            while (lo < hi) {
                const long long step = (hi - lo + 31) / 32;
                const long long probe = lo + lane * step;
                const bool ok = (probe < hi) && (keys[probe] == row);
                const unsigned m = __ballot_sync(FULL, ok);
                const int c = (m == FULL) ? 32 : (__ffs(~m) - 1);
                if (c == 0) break;
                const long long nhi = lo + c * step;
                lo = lo + (c - 1) * step + 1;
                hi = nhi < hi ? nhi : hi;
            }
            const uint32_t full = (uint32_t)(lo - (long long)q0);
            if (lane == 0) {
                const uint32_t nsl = (full + DAISY_SLICE - 1) / DAISY_SLICE;
                const uint32_t idx = atomicAdd(&heavy[0], 1u);
                const uint32_t sl0 = atomicAdd(&heavy[1], nsl);
                if ((int)idx < heavy_cap) {
                    uint32_t *rec = heavy + 2 + 5 * (size_t)idx;
                    rec[0] = (uint32_t)tbl;
                    rec[1] = row;
                    rec[2] = (uint32_t)q0;
                    rec[3] = full;
                    rec[4] = sl0;
                }
            }
            continue;
        }
        float4 old[V], acc[V];
#pragma unroll
        for (int v = 0; v < V; ++v) {
            old[v] = act[v] ? ld_row(table, (size_t)row * D4 + lane + 32 * v) : f4_zero();
            acc[v] = f4_zero();
        }
        sum_staged<V>(stage, q0, len, D4, lane, act, acc);
#pragma unroll
        for (int v = 0; v < V; ++v)
            if (act[v]) opt.apply(tbl, (size_t)row * D4 + lane + 32 * v, old[v], acc[v]);
    }
}

// Level 1: every slice of DAISY_SLICE consecutive staged contributions of a hot row is summed by one warp into
// stage2[first_slice + j].  SPLIT blocks share one hot row so that even the hottest row is spread over
// SPLIT * 8 warps.
template <int V>
__global__ void __launch_bounds__(256) k_heavy_slices(const float *__restrict__ stageU, const float *__restrict__ stageQ,
                                                       float *__restrict__ stage2, int D4,
                                                       const uint32_t *__restrict__ heavy, int heavy_cap, int split) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int count = min((int)heavy[0], heavy_cap);
    bool act[V];
#pragma unroll
    for (int v = 0; v < V; ++v) act[v] = (lane + 32 * v) < D4;
    for (int w = blockIdx.x; w < count * split; w += gridDim.x) {
        const uint32_t *rec = heavy + 2 + 5 * (size_t)(w / split);
        const int part = w % split;
        const float *stage = rec[0] ? stageQ : stageU;
        const size_t q0 = rec[2];
        const int len = (int)rec[3];
        const size_t sl0 = rec[4];
        const int nsl = (len + DAISY_SLICE - 1) / DAISY_SLICE;
        for (int j = part + split * wid; j < nsl; j += split * 8) {
            const int c0 = j * DAISY_SLICE;
            const int cn = min(DAISY_SLICE, len - c0);
            float4 acc[V];
#pragma unroll
            for (int v = 0; v < V; ++v) acc[v] = f4_zero();
            sum_staged<V>(stage, q0 + c0, cn, D4, lane, act, acc);
#pragma unroll
            for (int v = 0; v < V; ++v)
                if (act[v]) st_stream(stage2, (sl0 + j) * D4 + lane + 32 * v, acc[v]);
        }
    }
}

// Level 2: one block per hot row sums the row's slice partials (fixed split over 8 warps, fixed combine order)
// and applies the update.
template <int V, class Opt>
__global__ void __launch_bounds__(256) k_heavy_final(const float *__restrict__ P, const float *__restrict__ Q,
                                                      const float *__restrict__ stage2, int D4, Opt opt,
                                                      const uint32_t *__restrict__ heavy, int heavy_cap) {
    __shared__ float4 part[8][32 * V];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int count = min((int)heavy[0], heavy_cap);
    bool act[V];
#pragma unroll
    for (int v = 0; v < V; ++v) act[v] = (lane + 32 * v) < D4;
    for (int idx = blockIdx.x; idx < count; idx += gridDim.x) {
        const uint32_t *rec = heavy + 2 + 5 * (size_t)idx;
        const int tbl = (int)rec[0];
        const uint32_t row = rec[1];
        const int nsl = ((int)rec[3] + DAISY_SLICE - 1) / DAISY_SLICE;
        const size_t sl0 = rec[4];
        const float *table = tbl ? Q : P;
        const int per = (nsl + 7) / 8;
        const int c0 = min(nsl, wid * per), c1 = min(nsl, c0 + per);
        float4 acc[V];
#pragma unroll
        for (int v = 0; v < V; ++v) acc[v] = f4_zero();
        sum_staged<V>(stage2, sl0 + c0, c1 - c0, D4, lane, act, acc);
#pragma unroll
        for (int v = 0; v < V; ++v) part[wid][lane + 32 * v] = acc[v];
        __syncthreads();
        if (wid == 0) {
#pragma unroll
            for (int v = 0; v < V; ++v)
                if (act[v]) {
                    float4 tot = part[0][lane + 32 * v];
                    for (int ww = 1; ww < 8; ++ww) tot = f4_add(tot, part[ww][lane + 32 * v]);  // fixed order
                    const size_t e = (size_t)row * D4 + lane + 32 * v;
                    opt.apply(tbl, e, ld_row(table, e), tot);
                }
        }
        __syncthreads();
    }
}

// ------------------------------------------------------------------------------------------------
// owner side of the sharded step: received (row, gradient-sum) pairs from all ranks, concatenated in rank order.
// A stable sort by row keeps the rank order inside each row; one warp per window of 32 sorted entries sums each
// row's contributions in that order and applies  Q[row] += alpha * sum  once.
// ------------------------------------------------------------------------------------------------
__global__ void k_owner_keys(const int32_t *__restrict__ rows, int n, uint32_t I, uint32_t *__restrict__ key,
                             uint32_t *__restrict__ val, int *err) {
    int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n) return;
    uint32_t r = (uint32_t)rows[p];
    if (r >= I) {
        atomicOr(&err[0], 1);
        atomicMin(&err[1], p);
        r = 0;
    }
    key[p] = r;
    val[p] = (uint32_t)p;
}

template <int V>
__global__ void __launch_bounds__(256) k_owner_apply(float *__restrict__ Q, const uint32_t *__restrict__ keys,
                                                      const uint32_t *__restrict__ perm, int n,
                                                      const float *__restrict__ grads, int D4, float alpha) {
    const unsigned FULL = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    const long long base = (long long)((blockIdx.x * (size_t)blockDim.x + threadIdx.x) >> 5) * 32;
    if (base >= n) return;
    const long long p = base + lane;
    const uint32_t key = (p < n) ? keys[p] : 0xFFFFFFFFu;
    uint32_t prev = __shfl_up_sync(FULL, key, 1);
    if (lane == 0) prev = (p > 0) ? keys[p - 1] : ~key;
    const bool start = (p < n) && (prev != key);
    unsigned todo = __ballot_sync(FULL, start);
    bool act[V];
#pragma unroll
    for (int v = 0; v < V; ++v) act[v] = (lane + 32 * v) < D4;
    while (todo) {
        const int b = __ffs(todo) - 1;
        todo &= todo - 1;
        const uint32_t row = __shfl_sync(FULL, key, b);
        float4 acc[V];
#pragma unroll
        for (int v = 0; v < V; ++v) acc[v] = f4_zero();
        for (long long q = base + b; q < n; ++q) {  // contributions of this row: at most one per rank
            if (keys[q] != row) break;
            const size_t src = perm[q];
#pragma unroll
            for (int v = 0; v < V; ++v)
                if (act[v]) acc[v] = f4_add(acc[v], ld_stream(grads, src * D4 + lane + 32 * v));
        }
#pragma unroll
        for (int v = 0; v < V; ++v)
            if (act[v]) {
                const size_t e = (size_t)row * D4 + lane + 32 * v;
                const float4 old = ld_row(Q, e);
                st_row(Q, e, make_float4(fmaf(alpha, acc[v].x, old.x), fmaf(alpha, acc[v].y, old.y),
                                         fmaf(alpha, acc[v].z, old.z), fmaf(alpha, acc[v].w, old.w)));
            }
    }
}

// ------------------------------------------------------------------------------------------------
// loss: fixed-order reduction of per-warp partials (double accumulation), added to *loss_accum
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(1024) k_loss(const float *__restrict__ part, int n, double *loss_accum) {
    __shared__ double sh[32];
    double s = 0.0;
    for (int i = threadIdx.x; i < n; i += 1024) s += (double)part[i];
    s = warp_sum_d(s);
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x < 32) {
        double t = sh[threadIdx.x];
        t = warp_sum_d(t);
        if (threadIdx.x == 0) *loss_accum += t;
    }
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
static int bits_for(uint64_t max_value) {  // number of low bits needed to represent max_value
    int b = 1;
    while (b < 32 && (max_value >> b)) ++b;
    return b;
}

static int auto_chunk(const daisy_ctx *h, int64_t B) {
    if (h->chunk > 0) return h->chunk;
    // keep >= ~4 waves of 32 resident warps per SM before growing the chunk
    const int64_t want_warps = (int64_t)h->num_sms * 32 * 4;
    int c = 1;
    while (c < 16 && B / (2 * c) >= want_warps) c *= 2;
    return c;
}

static inline void phase_mark(daisy_ctx *h, int ph, cudaStream_t s) {
    if (h->timing == 2) cudaEventRecord(h->ev[ph + 1], s);
}

template <int V, class Opt>
static int run_step_v(daisy_ctx *h, const float *P, const float *Q, const int32_t *triples, int64_t B64, const Opt &opt,
                      float c2, double *loss_accum, cudaStream_t s, const int32_t *host_src, bool inputs_ready) {
    const int B = (int)B64;
    const int D4 = h->D / 4;
    const int C = auto_chunk(h, B);
    // the sharded step runs against a cache of fetched item rows whose row count differs from the local item shard
    const uint32_t U = (uint32_t)h->U, I = (uint32_t)(h->item_rows_override ? h->item_rows_override : h->I);
    const int T = 256;
    // The integer bookkeeping of a step (prep .. slots) depends on the triples only, never on the tables.  It
    // runs on the handle's side stream into one of two bookkeeping sets, so that for step n+1 it overlaps the
    // bandwidth-bound kernels of step n on the caller's stream.  Per-phase timing (mode 2) serialises everything
    // on the caller's stream so that phase times do not overlap.
    const bool piped = h->pipeline && h->timing != 2;
    cudaStream_t bs = piped ? h->side_stream : s;
    BookSet &k = h->book[h->book_idx];
    h->book_idx ^= 1;
    if (h->timing == 2) {
        if (h->ev_pending) {  // fold the previous step's phase times in
            cudaEventSynchronize(h->ev[PH_COUNT]);
            for (int ph = 0; ph < PH_COUNT; ++ph) {
                float ms = 0.f;
                cudaEventElapsedTime(&ms, h->ev[ph], h->ev[ph + 1]);
                h->phase_ms_sum[ph] += ms;
            }
            h->timed_steps++;
            h->ev_pending = 0;
        }
        cudaEventRecord(h->ev[0], s);
    }
    if (piped) {
        if (!inputs_ready) {  // the triples may have been produced by earlier work on the caller's stream
            DAISY_CUDA(cudaEventRecord(h->ev_call, s));
            DAISY_CUDA(cudaStreamWaitEvent(bs, h->ev_call, 0));
        }
        DAISY_CUDA(cudaStreamWaitEvent(bs, k.freed, 0));  // the step that last used this set has finished
    }
    const bool tr = h->trace && h->tr_n < DAISY_TRACE_STEPS;
    if (tr) cudaEventRecord(h->tr_ev[4 * h->tr_n + 0], bs);
    if (host_src) {  // daisy_bpr_step_host: the H2D copy is the first node of the bookkeeping chain
        DAISY_CUDA(cudaMemcpyAsync((void *)triples, host_src, (size_t)B * 3 * sizeof(int32_t), cudaMemcpyHostToDevice, bs));
    }
    // prep
    k_prep<<<daisy_ceil_div(B, T), T, 0, bs>>>(triples, B, U, I, h->ikey_in, h->ival_in, h->err);
    DAISY_LAUNCH_CHECK(h);
    phase_mark(h, PH_PREP, s);
    // sort by positive item
    size_t tmp = h->cub_tmp_bytes;
    DAISY_CUDA(cub::DeviceRadixSort::SortPairs(h->cub_tmp, tmp, h->ikey_in, h->ikey_out, h->ival_in, h->ival_out, B, 0,
                                               bits_for(I - 1), bs));
    h->launches += 4;
    phase_mark(h, PH_SORT_I, s);
    // refs
    k_refs<<<daisy_ceil_div(B, T), T, 0, bs>>>(triples, h->ival_out, h->ikey_out, B, U, I, C, k.st, h->ukey_in,
                                               h->uval_in, h->key_in, h->val_in);
    DAISY_LAUNCH_CHECK(h);
    phase_mark(h, PH_REFS, s);
    tmp = h->cub_tmp_bytes;
    DAISY_CUDA(cub::DeviceRadixSort::SortPairs(h->cub_tmp, tmp, h->ukey_in, k.ukey_s, h->uval_in, h->uval_out, B, 0,
                                               bits_for(U - 1), bs));
    h->launches += 4;
    phase_mark(h, PH_SORT_U, s);
    tmp = h->cub_tmp_bytes;
    DAISY_CUDA(cub::DeviceRadixSort::SortPairs(h->cub_tmp, tmp, h->key_in, k.qkey_s, h->val_in, h->val_out, 2 * B, 0,
                                               bits_for(I), bs));
    h->launches += 4;
    phase_mark(h, PH_SORT_Q, s);
    // slots
    k_slots_user<<<daisy_ceil_div(B, T), T, 0, bs>>>(k.ukey_s, h->uval_out, B, k.uslot);
    DAISY_LAUNCH_CHECK(h);
    k_slots_item<<<daisy_ceil_div(2 * (int64_t)B, T), T, 0, bs>>>(k.qkey_s, h->val_out, 2 * B, B, I, k.jslot, k.islot);
    DAISY_LAUNCH_CHECK(h);
    phase_mark(h, PH_SLOTS, s);
    if (piped) {
        DAISY_CUDA(cudaEventRecord(k.ready, bs));
        DAISY_CUDA(cudaStreamWaitEvent(s, k.ready, 0));
    }
    // main
    if (tr) {
        cudaEventRecord(h->tr_ev[4 * h->tr_n + 1], bs);
        cudaEventRecord(h->tr_ev[4 * h->tr_n + 2], s);
    }
    DAISY_CUDA(cudaMemsetAsync(h->heavy, 0, 2 * sizeof(uint32_t), s));
    MainArgs a;
    a.P = P; a.Q = Q; a.st = k.st; a.uslot = k.uslot; a.jslot = k.jslot; a.islot = k.islot;
    a.stageU = h->stageU; a.stageQ = h->stageQ; a.loss_part = h->loss_part;
    a.B = B; a.D4 = D4; a.C = C; a.c2 = c2;
    const int warps = daisy_ceil_div(B, C);
    const bool pool = (h->timing == 1 && h->pool_used < DAISY_EVPOOL);
    if (pool) cudaEventRecord(h->evpool[2 * h->pool_used], s);
    k_bpr_main<V, Opt><<<daisy_ceil_div(warps, 8), 256, 0, s>>>(a, opt);
    DAISY_LAUNCH_CHECK(h);
    if (pool) {
        cudaEventRecord(h->evpool[2 * h->pool_used + 1], s);
        h->pool_used++;
    }
    phase_mark(h, PH_MAIN, s);
    // segmented reduces
    k_seg_reduce<V, Opt><<<daisy_ceil_div(daisy_ceil_div(B, 32), 8), 256, 0, s>>>(
        0, P, k.ukey_s, B, 0xFFFFFFFFu, h->stageU, D4, opt, h->heavy_len, h->heavy, h->heavy_cap);
    DAISY_LAUNCH_CHECK(h);
    phase_mark(h, PH_SEG_U, s);
    k_seg_reduce<V, Opt><<<daisy_ceil_div(daisy_ceil_div(2 * (int64_t)B, 32), 8), 256, 0, s>>>(
        1, Q, k.qkey_s, 2 * B, I, h->stageQ, D4, opt, h->heavy_len, h->heavy, h->heavy_cap);
    DAISY_LAUNCH_CHECK(h);
    phase_mark(h, PH_SEG_Q, s);
    const int split = 8;
    k_heavy_slices<V><<<h->num_sms * 4, 256, 0, s>>>(h->stageU, h->stageQ, h->stage2, D4, h->heavy, h->heavy_cap, split);
    DAISY_LAUNCH_CHECK(h);
    k_heavy_final<V, Opt><<<h->num_sms, 256, 0, s>>>(P, Q, h->stage2, D4, opt, h->heavy, h->heavy_cap);
    DAISY_LAUNCH_CHECK(h);
    phase_mark(h, PH_HEAVY, s);
    if (loss_accum) {
        k_loss<<<1, 1024, 0, s>>>(h->loss_part, warps, loss_accum);
        DAISY_LAUNCH_CHECK(h);
    }
    phase_mark(h, PH_LOSS, s);
    if (piped) DAISY_CUDA(cudaEventRecord(k.freed, s));
    if (tr) {
        cudaEventRecord(h->tr_ev[4 * h->tr_n + 3], s);
        h->tr_n++;
    }
    if (h->timing == 2) {
        h->ev_pending = 1;
        h->ev_stream = s;
    }
    return DAISY_OK;
}

template <class Opt>
static int run_step(daisy_ctx *h, const float *P, const float *Q, const int32_t *triples, int64_t B, const Opt &opt,
                    float c2, double *loss_accum, cudaStream_t s, const int32_t *host_src, bool inputs_ready) {
    const int D4 = h->D / 4;
    if (D4 <= 32) return run_step_v<1, Opt>(h, P, Q, triples, B, opt, c2, loss_accum, s, host_src, inputs_ready);
    if (D4 <= 64) return run_step_v<2, Opt>(h, P, Q, triples, B, opt, c2, loss_accum, s, host_src, inputs_ready);
    if (D4 <= 96) return run_step_v<3, Opt>(h, P, Q, triples, B, opt, c2, loss_accum, s, host_src, inputs_ready);
    return run_step_v<4, Opt>(h, P, Q, triples, B, opt, c2, loss_accum, s, host_src, inputs_ready);
}

static int check_step_args(daisy_ctx *h, const void *P, const void *Q, const void *triples, int64_t B) {
    DAISY_REQUIRE(h != nullptr, DAISY_EINVAL, "null handle");
    DAISY_REQUIRE(P && Q && (triples || B == 0), DAISY_EINVAL, "null table or triples pointer");
    DAISY_REQUIRE(h->maxB > 0 && h->D % 4 == 0 && h->D <= 512, DAISY_EUNSUPPORTED,
                  "BPR step needs a handle created with max_batch > 0 and dim %% 4 == 0, dim <= 512 (dim is %d)", h->D);
    DAISY_REQUIRE(B >= 0 && B <= h->maxB, DAISY_EINVAL, "batch of %lld triples exceeds max_batch %lld", (long long)B,
                  (long long)h->maxB);
    DAISY_REQUIRE(((uintptr_t)P % 16 == 0) && ((uintptr_t)Q % 16 == 0), DAISY_EINVAL, "tables must be 16-byte aligned");
    return DAISY_OK;
}

// shared body of daisy_bpr_step / daisy_bpr_step_host
static int sgd_step(daisy_ctx *h, float *P, float *Q, const int32_t *triples_dev, const int32_t *host_src, int64_t B,
                    float lr, float wd, double *loss_accum, daisy_stream_t stream) {
    const double shrink = 1.0 - (double)lr * (double)wd;
    DAISY_REQUIRE(shrink > 0.0, DAISY_EINVAL, "lr*wd = %g >= 1: the L2 shrink factor is not positive", (double)lr * wd);
    if (B == 0) {  // an empty batch still decays every row (optim.SGD.step with zero gradients)
        h->scale *= shrink;
        return DAISY_OK;
    }
    DeviceGuard g(h->device);
    DAISY_REQUIRE(g.ok, DAISY_ECUDA, "cannot select device %d", h->device);
    SgdOpt opt;
    opt.P = P;
    opt.Q = Q;
    opt.alpha = (float)((double)lr / shrink);
    const float c2 = (float)(h->scale * h->scale);
    int rc = run_step<SgdOpt>(h, P, Q, triples_dev, B, opt, c2, loss_accum, (cudaStream_t)stream, host_src,
                              host_src != nullptr || h->inputs_ready);
    if (rc) return rc;
    h->scale *= shrink;
    if ((h->flags & DAISY_FLAG_EAGER_DECAY) || h->scale < 1e-4) return daisy_materialize(h, P, Q, stream);
    return DAISY_OK;
}

}  // namespace

// ------------------------------------------------------------------------------------------------
// C ABI
// ------------------------------------------------------------------------------------------------
extern "C" int daisy_bpr_step(daisy_handle_t h, float *P, float *Q, const int32_t *triples, int64_t B, float lr,
                              float wd, double *loss_accum, daisy_stream_t stream) {
    int rc = check_step_args(h, P, Q, triples, B);
    if (rc) return rc;
    return sgd_step(h, P, Q, triples, nullptr, B, lr, wd, loss_accum, stream);
}

extern "C" int daisy_bpr_step_host(daisy_handle_t h, float *P, float *Q, const int32_t *triples_host, int64_t B,
                                   float lr, float wd, double *loss_accum, daisy_stream_t stream) {
    int rc = check_step_args(h, P, Q, triples_host, B);
    if (rc) return rc;
    // landing buffer of the bookkeeping set this step will use; the H2D copy is issued on the side stream as the
    // first node of the step's bookkeeping chain, so it overlaps the previous step's kernels
    int32_t *dst = h->triples + (size_t)h->book_idx * 3 * (size_t)h->maxB;
    return sgd_step(h, P, Q, dst, triples_host, B, lr, wd, loss_accum, stream);
}

extern "C" int daisy_bpr_shard_step(daisy_handle_t h, float *P_local, const float *cache, int64_t cache_rows,
                                    const int32_t *triples, int64_t B, float lr, float wd, float *grad_out,
                                    double *loss_accum, daisy_stream_t stream) {
    if (h && B == 0) {  // a rank without triples this step still decays its rows
        const double sh = 1.0 - (double)lr * (double)wd;
        DAISY_REQUIRE(sh > 0.0, DAISY_EINVAL, "lr*wd = %g >= 1", (double)lr * wd);
        h->scale *= sh;
        return DAISY_OK;
    }
    int rc = check_step_args(h, P_local, cache, triples, B);
    if (rc) return rc;
    DAISY_REQUIRE(grad_out != nullptr, DAISY_EINVAL, "null grad_out");
    DAISY_REQUIRE(cache_rows > 0 && cache_rows <= 2 * h->maxB, DAISY_EINVAL,
                  "cache_rows %lld must be in [1, 2 * max_batch]", (long long)cache_rows);
    DAISY_REQUIRE((uintptr_t)grad_out % 16 == 0, DAISY_EINVAL, "grad_out must be 16-byte aligned");
    const double shrink = 1.0 - (double)lr * (double)wd;
    DAISY_REQUIRE(shrink > 0.0, DAISY_EINVAL, "lr*wd = %g >= 1", (double)lr * wd);
    if (B == 0) {
        h->scale *= shrink;
        return DAISY_OK;
    }
    DeviceGuard g(h->device);
    DAISY_REQUIRE(g.ok, DAISY_ECUDA, "cannot select device %d", h->device);
    ShardOpt opt;
    opt.P = P_local;
    opt.G = grad_out;
    opt.alpha = (float)((double)lr / shrink);
    h->item_rows_override = cache_rows;
    rc = run_step<ShardOpt>(h, P_local, cache, triples, B, opt, (float)(h->scale * h->scale), loss_accum,
                            (cudaStream_t)stream, nullptr, h->inputs_ready != 0);
    h->item_rows_override = 0;
    if (rc) return rc;
    h->scale *= shrink;
    return DAISY_OK;
}

extern "C" int daisy_owner_apply(daisy_handle_t h, float *Q_local, const int32_t *rows, const float *grads, int64_t n,
                                 float lr, float wd, daisy_stream_t stream) {
    DAISY_REQUIRE(h && Q_local, DAISY_EINVAL, "null argument");
    DAISY_REQUIRE(h->D % 4 == 0 && h->D <= 512, DAISY_EUNSUPPORTED, "dim %d unsupported", h->D);
    DAISY_REQUIRE(n >= 0 && n < (1LL << 30), DAISY_EINVAL, "bad row count");
    const double shrink = 1.0 - (double)lr * (double)wd;
    DAISY_REQUIRE(shrink > 0.0, DAISY_EINVAL, "lr*wd = %g >= 1", (double)lr * wd);
    if (n == 0) return DAISY_OK;
    DAISY_REQUIRE(rows && grads, DAISY_EINVAL, "null argument");
    DeviceGuard g(h->device);
    DAISY_REQUIRE(g.ok, DAISY_ECUDA, "cannot select device %d", h->device);
    cudaStream_t s = (cudaStream_t)stream;
    if (h->own_cap < n) {  // grow-only scratch: (key, value) ping/pong + CUB temp
        DAISY_CUDA(cudaStreamSynchronize(s));
        for (void *p : {(void *)h->own_key, (void *)h->own_key_s, (void *)h->own_val, (void *)h->own_val_s, h->own_tmp})
            if (p) cudaFree(p);
        h->own_key = h->own_key_s = h->own_val = h->own_val_s = nullptr;
        h->own_tmp = nullptr;
        const size_t cap = (size_t)n + (size_t)n / 4 + 1024;
        size_t tb = 0;
        cub::DeviceRadixSort::SortPairs(nullptr, tb, (uint32_t *)nullptr, (uint32_t *)nullptr, (uint32_t *)nullptr,
                                        (uint32_t *)nullptr, (int)cap, 0, 32, s);
        bool ok = cudaMalloc((void **)&h->own_key, cap * 4) == cudaSuccess &&
                  cudaMalloc((void **)&h->own_key_s, cap * 4) == cudaSuccess &&
                  cudaMalloc((void **)&h->own_val, cap * 4) == cudaSuccess &&
                  cudaMalloc((void **)&h->own_val_s, cap * 4) == cudaSuccess &&
                  cudaMalloc(&h->own_tmp, tb + 256) == cudaSuccess;
        DAISY_REQUIRE(ok, DAISY_ENOMEM, "owner-side scratch allocation failed");
        h->own_cap = (int64_t)cap;
        h->own_tmp_bytes = tb + 256;
    }
    const int T = 256;
    k_owner_keys<<<daisy_ceil_div(n, T), T, 0, s>>>(rows, (int)n, (uint32_t)h->I, h->own_key, h->own_val, h->err);
    DAISY_LAUNCH_CHECK(h);
    size_t tb = h->own_tmp_bytes;
    DAISY_CUDA(cub::DeviceRadixSort::SortPairs(h->own_tmp, tb, h->own_key, h->own_key_s, h->own_val, h->own_val_s, (int)n,
                                               0, bits_for((uint64_t)h->I - 1), s));
    h->launches += 4;
    const int D4 = h->D / 4;
    const float alpha = (float)((double)lr / shrink);
    const int grid = daisy_ceil_div(daisy_ceil_div(n, 32), 8);
    if (D4 <= 32) k_owner_apply<1><<<grid, 256, 0, s>>>(Q_local, h->own_key_s, h->own_val_s, (int)n, grads, D4, alpha);
    else if (D4 <= 64) k_owner_apply<2><<<grid, 256, 0, s>>>(Q_local, h->own_key_s, h->own_val_s, (int)n, grads, D4, alpha);
    else if (D4 <= 96) k_owner_apply<3><<<grid, 256, 0, s>>>(Q_local, h->own_key_s, h->own_val_s, (int)n, grads, D4, alpha);
    else k_owner_apply<4><<<grid, 256, 0, s>>>(Q_local, h->own_key_s, h->own_val_s, (int)n, grads, D4, alpha);
    DAISY_LAUNCH_CHECK(h);
    return DAISY_OK;
}

extern "C" int daisy_bpr_adam_step(daisy_handle_t h, float *P, float *Q, float *mP, float *vP, float *mQ, float *vQ,
                                   const int32_t *triples, int64_t B, float lr, float beta1, float beta2, float eps,
                                   int64_t step_no, double *loss_accum, daisy_stream_t stream) {
    int rc = check_step_args(h, P, Q, triples, B);
    if (rc) return rc;
    DAISY_REQUIRE(mP && vP && mQ && vQ, DAISY_EINVAL, "null Adam moment pointer");
    DAISY_REQUIRE(step_no >= 1, DAISY_EINVAL, "step_no is 1-based");
    DAISY_REQUIRE(h->scale == 1.0, DAISY_EINVAL, "lazy L2 scale is %g: call daisy_materialize before an Adam step", h->scale);
    if (B == 0) return DAISY_OK;
    DeviceGuard g(h->device);
    DAISY_REQUIRE(g.ok, DAISY_ECUDA, "cannot select device %d", h->device);
    AdamOpt opt;
    opt.P = P; opt.Q = Q; opt.mP = mP; opt.vP = vP; opt.mQ = mQ; opt.vQ = vQ;
    opt.b1 = beta1;
    opt.b2 = beta2;
    const double bc1 = 1.0 - pow((double)beta1, (double)step_no);
    const double bc2 = 1.0 - pow((double)beta2, (double)step_no);
    opt.step_size = (float)((double)lr / bc1);
    opt.inv_sqrt_bc2 = (float)(1.0 / sqrt(bc2));
    opt.eps = eps;
    return run_step<AdamOpt>(h, P, Q, triples, B, opt, 1.0f, loss_accum, (cudaStream_t)stream, nullptr,
                             h->inputs_ready != 0);
}
