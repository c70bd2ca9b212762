// Where the time of k_small_book goes: clock stamps of thread 0 of both blocks (user refs / item refs) at every phase
// of the block-wide radix sort.   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 --expt-relaxed-constexpr
//   -I include -o tools/small_book_probe tools/small_book_probe.cu
#define DAISY_SMALL_PROBE 1
#include "../recommend_lib_b200/csrc/step_kernels.cuh"
#include <vector>
void daisy_set_error(const char *, ...) {}
void daisy_shard_free(daisy_ctx *) {}
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1);} } while (0)

template <int IPT>
static void run(int B, uint32_t U, uint32_t I) {
    std::vector<int32_t> t(3 * (size_t)B);
    uint32_t r = 12345;
    for (int k = 0; k < B; ++k) {
        r = r * 1664525u + 1013904223u; t[3 * k] = (r >> 8) % U;
        r = r * 1664525u + 1013904223u; t[3 * k + 1] = (uint32_t)((double)(r >> 8) / 16777216.0 * (double)(r >> 8) / 16777216.0 * I) % I;
        r = r * 1664525u + 1013904223u; t[3 * k + 2] = (r >> 8) % I;
    }
    int32_t *tri, *st; uint32_t *uk, *qk, *us, *js, *is, *longs; int *err;
    CK(cudaMalloc(&longs, 8 + 20 * 4000));
    CK(cudaMalloc(&tri, 12 * B)); CK(cudaMalloc(&st, 12 * B)); CK(cudaMalloc(&uk, 4 * B)); CK(cudaMalloc(&qk, 8 * B));
    CK(cudaMalloc(&us, 4 * B)); CK(cudaMalloc(&js, 4 * B)); CK(cudaMalloc(&is, 4 * B)); CK(cudaMalloc(&err, 8));
    CK(cudaMemcpy(tri, t.data(), 12 * B, cudaMemcpyHostToDevice)); CK(cudaMemset(err, 0, 8));
    const int vbU = bits_for(B - 1), vbQ = bits_for(2 * B - 1), kbU = bits_for(U), kbQ = bits_for(I);
    const size_t smem = sizeof(uint32_t) * (DAISY_SMALL_HIST_WORDS + 1024 * IPT);
    CK(cudaFuncSetAttribute(k_small_book<IPT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float ms = 0;
    for (int it = 0; it < 5; ++it) {
        cudaEventRecord(e0);
        CK(cudaMemset(longs, 0, 8));
        k_small_book<IPT><<<2 << DAISY_SMALL_CB, 1024, smem>>>(tri, B, U, I, vbU, kbU, vbQ, kbQ, st, uk, qk, us, js, is, longs, 4000, err, 0);
        cudaEventRecord(e1);
        CK(cudaDeviceSynchronize());
        cudaEventElapsedTime(&ms, e0, e1);
    }
    long long pr[256][32];
    CK(cudaMemcpyFromSymbol(pr, g_small_probe, sizeof(pr)));
    printf("B %d IPT %d  U %u (kb %d) I %u (kb %d): %.1f us by events\n", B, IPT, U, kbU, I, kbQ, ms * 1e3);
    for (int b = 0; b < (2 << DAISY_SMALL_CB); b += 3) {
        printf("  block %d cycles: load %lld |", b, pr[b][1] - pr[b][0]);
        int nb_ = (b >= (1 << DAISY_SMALL_CB) ? kbQ : kbU) - DAISY_SMALL_CB; if (nb_ < 1) nb_ = 1;
        const int passes = (nb_ + 8) / 9;
        for (int p = 0; p < passes; ++p)
            printf(" pass%d rank %lld bar %lld scan %lld bar %lld scatter %lld |", p, pr[b][3 + 6 * p] - pr[b][2 + 6 * p],
                   pr[b][4 + 6 * p] - pr[b][3 + 6 * p], pr[b][5 + 6 * p] - pr[b][4 + 6 * p], pr[b][6 + 6 * p] - pr[b][5 + 6 * p],
                   pr[b][7 + 6 * p] - pr[b][6 + 6 * p]);
        printf(" slots %lld | total %lld\n", pr[b][31] - pr[b][30], pr[b][31] - pr[b][0]);
    }
}

struct ProbeOpt {  // the SGD functor of bpr_step.cu
    static constexpr bool kNeedOldItem = true;
    float *P, *Q;
    float alpha;
    int D4;
    __device__ __forceinline__ void apply(int tbl, size_t row, int e, float4 old, float4 d) const {
        const size_t idx = row * D4 + e;
        st_row(tbl ? Q : P, idx, make_float4(fmaf(alpha, d.x, old.x), fmaf(alpha, d.y, old.y), fmaf(alpha, d.z, old.z), fmaf(alpha, d.w, old.w)));
    }
};

// the whole small-batch step (book, main, seg) on Zipf(1.0) positives, with per-block clocks of k_seg_all
static void run_step(int B, uint32_t U, uint32_t I, int D) {
    std::vector<int32_t> t(3 * (size_t)B);
    uint32_t r = 777;
    const double H = log((double)I) + 0.5772;
    for (int k = 0; k < B; ++k) {
        r = r * 1664525u + 1013904223u; t[3 * k] = (r >> 8) % U;
        r = r * 1664525u + 1013904223u; const double x = (double)(r >> 8) / 16777216.0;
        t[3 * k + 1] = (int32_t)((uint32_t)(exp(x * H) - 1.0) % I) * 7919u % I;   // ~ Zipf(1), spread over the table
        r = r * 1664525u + 1013904223u; t[3 * k + 2] = (r >> 8) % I;
    }
    int32_t *tri, *st; uint32_t *uk, *qk, *us, *js, *is, *longs, *ticket; int *err; float *P, *Q, *sU, *sQ, *lp, *st2; double *loss;
    CK(cudaMalloc(&longs, 8 + 20 * 4000)); CK(cudaMalloc(&ticket, 16000)); CK(cudaMemset(ticket, 0, 16000)); CK(cudaMalloc(&st2, (size_t)8000 * D * 4));
    CK(cudaMalloc(&tri, 12 * B)); CK(cudaMalloc(&st, 12 * B)); CK(cudaMalloc(&uk, 4 * B)); CK(cudaMalloc(&qk, 8 * B));
    CK(cudaMalloc(&us, 4 * B)); CK(cudaMalloc(&js, 4 * B)); CK(cudaMalloc(&is, 4 * B)); CK(cudaMalloc(&err, 8));
    CK(cudaMalloc(&P, (size_t)U * D * 4)); CK(cudaMalloc(&Q, (size_t)I * D * 4)); CK(cudaMalloc(&sU, (size_t)B * D * 4));
    CK(cudaMalloc(&sQ, (size_t)2 * B * D * 4)); CK(cudaMalloc(&lp, 4 * B)); CK(cudaMalloc(&loss, 8));
    CK(cudaMemset(P, 0, (size_t)U * D * 4)); CK(cudaMemset(Q, 0, (size_t)I * D * 4)); CK(cudaMemset(loss, 0, 8));
    CK(cudaMemcpy(tri, t.data(), 12 * B, cudaMemcpyHostToDevice)); CK(cudaMemset(err, 0, 8));
    const int vbU = bits_for(B - 1), vbQ = bits_for(2 * B - 1), kbU = bits_for(U), kbQ = bits_for(I);
    const size_t smem = sizeof(uint32_t) * (DAISY_SMALL_HIST_WORDS + 1024 * 16);
    CK(cudaFuncSetAttribute(k_small_book<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    ProbeOpt opt{P, Q, 0.01f, D / 4};
    MainArgs a{};
    a.P = P; a.Q = Q; a.st = st; a.uslot = us; a.jslot = js; a.islot = is; a.stageU = sU; a.stageQ = sQ; a.loss_part = lp;
    a.B = B; a.D4 = D / 4; a.C = 1; a.c2 = 1.f; a.jsrc = nullptr; a.isrc = nullptr;
    const int blocksU = daisy_ceil_div(B, 8 * DAISY_SMALL_WIN), blocksQ = daisy_ceil_div(2 * B, 8 * DAISY_SMALL_WIN);
    cudaEvent_t e[4]; for (auto &x : e) cudaEventCreate(&x);
    float ms[3] = {0, 0, 0};
    for (int it = 0; it < 5; ++it) {
        cudaEventRecord(e[0]);
        cudaMemsetAsync(longs, 0, 8);
        k_small_book<16><<<2 << DAISY_SMALL_CB, 1024, smem>>>(tri, B, U, I, vbU, kbU, vbQ, kbQ, st, uk, qk, us, js, is, longs, 4000, err, 0);
        cudaEventRecord(e[1]);
        k_bpr_main<1, ProbeOpt, false><<<daisy_ceil_div(B, 8), 256>>>(a, opt);
        cudaEventRecord(e[2]);
        k_seg_all<1, ProbeOpt, DAISY_SMALL_WIN, DAISY_SMALL_SLICE><<<64 + blocksU + blocksQ + 1, 256>>>(P, Q, uk, qk, B, 2 * B, 0xFFFFFFFFu, sU, sQ, st2, D / 4, opt, 64, blocksU, blocksQ, DAISY_SMALL_SLICE, longs, 4000, ticket, lp, B, loss);
        cudaEventRecord(e[3]);
        CK(cudaDeviceSynchronize());
        for (int k = 0; k < 3; ++k) cudaEventElapsedTime(&ms[k], e[k], e[k + 1]);
    }
    uint32_t lg[2];
    CK(cudaMemcpy(lg, longs, 8, cudaMemcpyDeviceToHost));
    printf("step B %d U %u I %u D %d: book %.1f us, main %.1f us, seg %.1f us (events); %u long rows, %u slices\n", B, U, I, D,
           ms[0] * 1e3, ms[1] * 1e3, ms[2] * 1e3, lg[0], lg[1]);
}

// micro-probe: one block, 8 warps, each sums 64 contiguous 512-byte rows (UN rows in flight)
__device__ long long g_rd[8][2];
template <int UN, bool PRED>
__global__ void k_read_probe(const float *stage, int len, float *sink) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const long long t0 = clock64();
    float4 acc = f4_zero();
    const size_t q0 = (size_t)wid * 64;
    for (int c = 0; c < len; c += UN) {
        float4 r[UN];
#pragma unroll
        for (int jj = 0; jj < UN; ++jj) r[jj] = (!PRED || c + jj < len) ? ld_stream(stage, (q0 + c + jj) * 32 + lane) : f4_zero();
#pragma unroll
        for (int jj = 0; jj < UN; ++jj) if (!PRED || c + jj < len) acc = f4_add(acc, r[jj]);
    }
    const long long t1 = clock64();
    if (lane == 0) { g_rd[wid][0] = t1 - t0; }
    sink[threadIdx.x] = acc.x + acc.y + acc.z + acc.w;
}
__global__ void k_write_rows(float *stage, int rows) {  // one warp per row, like the main kernel's staging stores
    const int lane = threadIdx.x & 31;
    const int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (w < rows) st_stream(stage, (size_t)w * 32 + lane, make_float4(1.f, 2.f, 3.f, 4.f));
}
template <int UN, bool PRED>
static void read_probe(const char *what, bool write_first) {
    float *stage, *sink;
    CK(cudaMalloc(&stage, 512 * 64 * 8 * 4)); CK(cudaMalloc(&sink, 1024));
    CK(cudaMemset(stage, 0, 512 * 64 * 8 * 4));
    for (int it = 0; it < 3; ++it) {
        if (write_first) k_write_rows<<<64, 256>>>(stage, 512);
        k_read_probe<UN, PRED><<<1, 256>>>(stage, 64, sink);
        CK(cudaDeviceSynchronize());
    }
    long long pr[8][2];
    CK(cudaMemcpyFromSymbol(pr, g_rd, sizeof(pr)));
    printf("read probe %s UN %d pred %d: warp cycles for 64 rows:", what, UN, (int)PRED);
    for (int w = 0; w < 8; ++w) printf(" %lld", pr[w][0]);
    printf("\n");
}

// k_mid_book against a host check: equal rows adjacent, every ref's slot points at its own row, slots used once
#include <map>
static void check_mid(int B, uint32_t U, uint32_t I, int cb) {
    std::vector<int32_t> t(3 * (size_t)B);
    uint32_t r = 4242 + B;
    for (int k = 0; k < B; ++k) {
        r = r * 1664525u + 1013904223u; t[3 * k] = (r >> 8) % U;
        r = r * 1664525u + 1013904223u; t[3 * k + 1] = (k < B / 3) ? 1 % I : (r >> 8) % I;
        r = r * 1664525u + 1013904223u; t[3 * k + 2] = (r >> 8) % I;
    }
    int32_t *tri, *st; uint32_t *uk, *qk, *us, *js, *is, *longs, *buf; int *err;
    CK(cudaMalloc(&longs, 8 + 20 * 30000)); CK(cudaMalloc(&buf, 4 * 3 * (size_t)B * 4));
    CK(cudaMalloc(&tri, 12 * B)); CK(cudaMalloc(&st, 12 * B)); CK(cudaMalloc(&uk, 4 * B)); CK(cudaMalloc(&qk, 8 * B));
    CK(cudaMalloc(&us, 4 * B)); CK(cudaMalloc(&js, 4 * B)); CK(cudaMalloc(&is, 4 * B)); CK(cudaMalloc(&err, 8));
    CK(cudaMemcpy(tri, t.data(), 12 * B, cudaMemcpyHostToDevice)); CK(cudaMemset(err, 0, 8)); CK(cudaMemset(longs, 0, 8));
    CK(cudaMemset(uk, 0xEE, 4 * B)); CK(cudaMemset(qk, 0xEE, 8 * B)); CK(cudaMemset(us, 0xEE, 4 * B)); CK(cudaMemset(js, 0xEE, 4 * B)); CK(cudaMemset(is, 0xEE, 4 * B));
    const size_t smem = sizeof(uint32_t) * DAISY_SMALL_HIST_WORDS;
    CK(cudaFuncSetAttribute(k_mid_book, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    MidScratch ms; ms.ak = buf; ms.av = buf + 3 * (size_t)B; ms.bk = ms.av + 3 * (size_t)B; ms.bv = ms.bk + 3 * (size_t)B;
    k_mid_book<<<2 << cb, 1024, smem>>>(tri, B, U, I, cb, bits_for(U), bits_for(I), ms, st, uk, qk, us, js, is, longs, 30000, err);
    CK(cudaDeviceSynchronize());
    std::vector<uint32_t> huk(B), hqk(2 * B), hus(B), hjs(B), his(B);
    CK(cudaMemcpy(huk.data(), uk, 4 * B, cudaMemcpyDeviceToHost)); CK(cudaMemcpy(hqk.data(), qk, 8 * B, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(hus.data(), us, 4 * B, cudaMemcpyDeviceToHost)); CK(cudaMemcpy(hjs.data(), js, 4 * B, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(his.data(), is, 4 * B, cudaMemcpyDeviceToHost));
    long bad_adj = 0, bad_slot = 0, bad_direct = 0, bad_key = 0;
    auto check = [&](const std::vector<uint32_t> &keys, int n, auto row_of, auto slot_of, const char *name) {
        std::map<uint32_t, int> cnt, seen_end;
        for (int i = 0; i < n; ++i) cnt[row_of(i)]++;
        std::map<uint32_t, int> kc;
        for (int p = 0; p < n; ++p) { kc[keys[p]]++; if (p && keys[p] != keys[p - 1] && seen_end.count(keys[p])) bad_adj++; if (p && keys[p] != keys[p - 1]) seen_end[keys[p - 1]] = 1; }
        if (kc != cnt) bad_key++;
        std::vector<char> used(n, 0);
        for (int i = 0; i < n; ++i) {
            const uint32_t sl = slot_of(i), row = row_of(i);
            if (sl == DAISY_DIRECT) { if (cnt[row] != 1) bad_direct++; continue; }
            if (sl >= (uint32_t)n || keys[sl] != row || used[sl]) bad_slot++; else used[sl] = 1;
        }
        printf("  %s: n %d  key multiset %s, adjacency errors %ld, slot errors %ld, direct errors %ld\n", name, n, bad_key ? "WRONG" : "ok", bad_adj, bad_slot, bad_direct);
        bad_adj = bad_slot = bad_direct = bad_key = 0;
    };
    {
        static long long pr[256][32];
        CK(cudaMemcpyFromSymbol(pr, g_small_probe, sizeof(pr)));
        for (int b : {0, 1 << cb, (1 << cb) + 1, (2 << cb) - 1})
            printf("  block %d cycles: sweep1 %lld scan %lld sweep2 %lld sort %lld slots %lld total %lld\n", b, pr[b][1] - pr[b][0], pr[b][2] - pr[b][1],
                   pr[b][3] - pr[b][2], pr[b][4] - pr[b][3], pr[b][5] - pr[b][4], pr[b][5] - pr[b][0]);
    }
    printf("check_mid B %d U %u I %u cb %d\n", B, U, I, cb);
    check(huk, B, [&](int i) { return (uint32_t)t[3 * i]; }, [&](int i) { return hus[i]; }, "users");
    check(hqk, 2 * B, [&](int i) { return (uint32_t)(i < B ? t[3 * i + 2] : t[3 * (i - B) + 1]); },
          [&](int i) { return i < B ? hjs[i] : his[i - B]; }, "items");
}

// whole step through k_mid_book / k_bpr_main / k_seg_all against a double-precision host step
static void check_step(int B, uint32_t U, uint32_t I, int D, int cb, int NS) {
    std::vector<int32_t> t(3 * (size_t)B);
    uint32_t r = 99 + B;
    for (int k = 0; k < B; ++k) {
        r = r * 1664525u + 1013904223u; t[3 * k] = (k % 8 == 0) ? 2 % U : (r >> 8) % U;
        r = r * 1664525u + 1013904223u; t[3 * k + 1] = (k < B / 3) ? 1 % I : (r >> 8) % I;
        r = r * 1664525u + 1013904223u; t[3 * k + 2] = (r >> 8) % I;
    }
    std::vector<float> hP((size_t)U * D), hQ((size_t)I * D);
    for (auto &x : hP) { r = r * 1664525u + 1013904223u; x = ((int)(r >> 8) % 2001 - 1000) * 3e-4f; }
    for (auto &x : hQ) { r = r * 1664525u + 1013904223u; x = ((int)(r >> 8) % 2001 - 1000) * 3e-4f; }
    const float alpha = 0.02f;
    std::vector<double> eP(hP.begin(), hP.end()), eQ(hQ.begin(), hQ.end());
    for (int k = 0; k < B; ++k) {
        const int u = t[3 * k], i = t[3 * k + 1], j = t[3 * k + 2];
        double x = 0;
        for (int e = 0; e < D; ++e) x += (double)hP[(size_t)u * D + e] * ((double)hQ[(size_t)i * D + e] - (double)hQ[(size_t)j * D + e]);
        const double sg = 1.0 / (1.0 + exp(x));
        for (int e = 0; e < D; ++e) {
            const double pu = hP[(size_t)u * D + e], qi = hQ[(size_t)i * D + e], qj = hQ[(size_t)j * D + e];
            eP[(size_t)u * D + e] += alpha * sg * (qi - qj);
            eQ[(size_t)i * D + e] += alpha * sg * pu;
            eQ[(size_t)j * D + e] -= alpha * sg * pu;
        }
    }
    int32_t *tri, *st; uint32_t *uk, *qk, *us, *js, *is, *longs, *ticket, *buf; int *err; float *P, *Q, *sU, *sQ, *lp, *st2; double *loss;
    CK(cudaMalloc(&longs, 8 + 20 * 30000)); CK(cudaMalloc(&ticket, 120000)); CK(cudaMemset(ticket, 0, 120000));
    CK(cudaMalloc(&st2, (size_t)60000 * D * 4)); CK(cudaMalloc(&buf, 4 * 3 * (size_t)B * 4));
    CK(cudaMalloc(&tri, 12 * B)); CK(cudaMalloc(&st, 12 * B)); CK(cudaMalloc(&uk, 4 * B)); CK(cudaMalloc(&qk, 8 * B));
    CK(cudaMalloc(&us, 4 * B)); CK(cudaMalloc(&js, 4 * B)); CK(cudaMalloc(&is, 4 * B)); CK(cudaMalloc(&err, 8));
    CK(cudaMalloc(&P, (size_t)U * D * 4)); CK(cudaMalloc(&Q, (size_t)I * D * 4)); CK(cudaMalloc(&sU, (size_t)B * D * 4));
    CK(cudaMalloc(&sQ, (size_t)2 * B * D * 4)); CK(cudaMalloc(&lp, 4 * B)); CK(cudaMalloc(&loss, 8));
    CK(cudaMemcpy(P, hP.data(), hP.size() * 4, cudaMemcpyHostToDevice)); CK(cudaMemcpy(Q, hQ.data(), hQ.size() * 4, cudaMemcpyHostToDevice));
    CK(cudaMemset(loss, 0, 8)); CK(cudaMemcpy(tri, t.data(), 12 * B, cudaMemcpyHostToDevice)); CK(cudaMemset(err, 0, 8)); CK(cudaMemset(longs, 0, 8));
    const size_t smem = sizeof(uint32_t) * DAISY_SMALL_HIST_WORDS;
    CK(cudaFuncSetAttribute(k_mid_book, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    MidScratch ms; ms.ak = buf; ms.av = buf + 3 * (size_t)B; ms.bk = ms.av + 3 * (size_t)B; ms.bv = ms.bk + 3 * (size_t)B;
    k_mid_book<<<2 << cb, 1024, smem>>>(tri, B, U, I, cb, bits_for(U), bits_for(I), ms, st, uk, qk, us, js, is, longs, 30000, err);
    ProbeOpt opt{P, Q, alpha, D / 4};
    MainArgs a{};
    a.P = P; a.Q = Q; a.st = st; a.uslot = us; a.jslot = js; a.islot = is; a.stageU = sU; a.stageQ = sQ; a.loss_part = lp;
    a.B = B; a.D4 = D / 4; a.C = 1; a.c2 = 1.f; a.jsrc = nullptr; a.isrc = nullptr;
    const int blocksU = daisy_ceil_div(B, 8 * DAISY_SMALL_WIN), blocksQ = daisy_ceil_div(2 * B, 8 * DAISY_SMALL_WIN);
    k_bpr_main<1, ProbeOpt, false><<<daisy_ceil_div(B, 8), 256>>>(a, opt);
    k_seg_all<1, ProbeOpt, DAISY_SMALL_WIN, DAISY_SMALL_SLICE><<<NS + blocksU + blocksQ + 1, 256>>>(P, Q, uk, qk, B, 2 * B, 0xFFFFFFFFu, sU, sQ, st2, D / 4, opt, NS, blocksU, blocksQ, DAISY_SMALL_SLICE, longs, 30000, ticket, lp, B, loss);
    CK(cudaDeviceSynchronize());
    std::vector<float> gP(hP.size()), gQ(hQ.size());
    CK(cudaMemcpy(gP.data(), P, gP.size() * 4, cudaMemcpyDeviceToHost)); CK(cudaMemcpy(gQ.data(), Q, gQ.size() * 4, cudaMemcpyDeviceToHost));
    std::vector<int32_t> hst(3 * (size_t)B);
    CK(cudaMemcpy(hst.data(), st, 12 * B, cudaMemcpyDeviceToHost));
    long st_bad = 0; for (size_t k = 0; k < hst.size(); ++k) st_bad += hst[k] != t[k];
    uint32_t lg[2]; CK(cudaMemcpy(lg, longs, 8, cudaMemcpyDeviceToHost));
    double eu = 0, ei = 0, mu = 0, mi = 0; long wu = -1, wi = -1;
    for (size_t k = 0; k < gP.size(); ++k) { if (fabs(gP[k] - eP[k]) > eu) { eu = fabs(gP[k] - eP[k]); wu = k / D; } mu = fmax(mu, fabs(eP[k])); }
    for (size_t k = 0; k < gQ.size(); ++k) { if (fabs(gQ[k] - eQ[k]) > ei) { ei = fabs(gQ[k] - eQ[k]); wi = k / D; } mi = fmax(mi, fabs(eQ[k])); }
    printf("check_step B %d U %u I %u D %d cb %d NS %d: rel err P %.2e (row %ld) Q %.2e (row %ld); st mismatches %ld; long rows %u slices %u\n",
           B, U, I, D, cb, NS, eu / mu, wu, ei / mi, wi, st_bad, lg[0], lg[1]);
}

int main(int argc, char **argv) {
    if (argc > 1 && argv[1][0] == 's') {
        check_step(8193, 301, 203, 64, 4, 296); check_step(8193, 301, 203, 64, 5, 296); check_step(8193, 301, 203, 64, 4, 64);
        check_step(20000, 5000, 70000, 32, 5, 296); check_step(9000, 257, 129, 128, 4, 296); check_step(4096, (1u << 21) + 3, 1000, 8, 4, 64);
        return 0;
    }
    if (argc > 1 && argv[1][0] == 'm') {
        check_mid(65536, 138493, 27278, 5); check_mid(16384, 138493, 27278, 4);
        check_mid(8193, 301, 203, 4); check_mid(8193, 301, 203, 5); check_mid(20000, 5000, 70000, 5); check_mid(20000, 5000, 70000, 4);
        check_mid(4096, (1u << 21) + 3, 1000, 4); check_mid(9000, 257, 129, 4);
        return 0;
    }
    read_probe<8, true>("after st_stream kernel", true);
    read_probe<8, false>("after st_stream kernel", true);
    read_probe<2, true>("after st_stream kernel", true);
    read_probe<8, true>("after memset", false);
    read_probe<16, false>("after st_stream kernel", true);
    run_step(4096, 138493, 27278, 128);
    run_step(8192, 138493, 27278, 128);
    if (argc > 1) return 0;
    run<8>(4096, 138493, 27278);
    run<16>(8192, 138493, 27278);
    run<8>(4096, 943, 1682);
    return 0;
}
