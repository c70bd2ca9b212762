// SM partitioning for the pipelined BPR step (CUDA green contexts, driver API through cudaGetDriverEntryPoint: the
// library keeps linking against the runtime only).
//
// Why: the integer bookkeeping of step n+1 (13-19 short, dependent kernels: three radix sorts and their glue) is meant
// to run UNDER the bandwidth-bound table kernels of step n.  On one shared set of SMs it does not: the main kernel
// keeps every SM's register file full (4 blocks x 256 threads x 62 registers), so each bookkeeping kernel -- however
// high its stream's priority -- has to wait for main blocks to retire before its own blocks can be placed, once per
// dependent launch.  Measured at config 4 (profiles/r02f_bench_trace.json): the bookkeeping chain takes 0.33 ms alone
// and 0.66 ms next to the table kernels, which makes IT the period of the pipeline.  Spatial partitioning removes the
// wait: a few SMs belong to the bookkeeping stream, the rest to the table kernels.
//
// daisy_partition_create(h, n): green context A with >= n SMs (rounded up to the architecture's granularity of 8) and
// green context B with the remaining SMs, one stream in each.  The contexts share the primary context's address
// space, modules and events, so the runtime-API launches of step_kernels.cuh work on these streams unchanged.
#include <cuda.h>

#include "ctx.cuh"

namespace {

template <class F>
bool entry(const char *name, F *fn) {
    void *p = nullptr;
    cudaDriverEntryPointQueryResult q = cudaDriverEntryPointSymbolNotFound;
    if (cudaGetDriverEntryPoint(name, &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess || !p) {
        cudaGetLastError();
        return false;
    }
    *fn = reinterpret_cast<F>(p);
    return true;
}

}  // namespace

int daisy_partition_create(daisy_ctx *h, int book_sms) {
    h->part_ok = 0;
    if (book_sms <= 0) return DAISY_OK;
    decltype(&cuDeviceGet) p_device_get = nullptr;
    decltype(&cuDeviceGetDevResource) p_get_res = nullptr;
    decltype(&cuDevSmResourceSplitByCount) p_split = nullptr;
    decltype(&cuDevResourceGenerateDesc) p_desc = nullptr;
    decltype(&cuGreenCtxCreate) p_create = nullptr;
    decltype(&cuGreenCtxStreamCreate) p_stream = nullptr;
    decltype(&cuGreenCtxDestroy) p_destroy = nullptr;
    if (!entry("cuDeviceGet", &p_device_get) || !entry("cuDeviceGetDevResource", &p_get_res) ||
        !entry("cuDevSmResourceSplitByCount", &p_split) || !entry("cuDevResourceGenerateDesc", &p_desc) ||
        !entry("cuGreenCtxCreate", &p_create) || !entry("cuGreenCtxStreamCreate", &p_stream) ||
        !entry("cuGreenCtxDestroy", &p_destroy)) {
        daisy_set_error("SM partitioning: the driver has no green-context entry points");
        return DAISY_EUNSUPPORTED;
    }
    cudaFree(0);  // the primary context exists and is current
    CUdevice dev;
    if (p_device_get(&dev, h->device) != CUDA_SUCCESS) return DAISY_ECUDA;
    CUdevResource all, part[1], rest;
    if (p_get_res(dev, &all, CU_DEV_RESOURCE_TYPE_SM) != CUDA_SUCCESS) {
        daisy_set_error("SM partitioning: cuDeviceGetDevResource failed");
        return DAISY_EUNSUPPORTED;
    }
    unsigned int groups = 1;
    if (p_split(part, &groups, &all, &rest, 0, (unsigned)book_sms) != CUDA_SUCCESS || groups != 1 ||
        rest.sm.smCount == 0) {
        daisy_set_error("SM partitioning: cannot split %u SMs into %d + the rest", all.sm.smCount, book_sms);
        return DAISY_EUNSUPPORTED;
    }
    CUdevResourceDesc d_book, d_main;
    if (p_desc(&d_book, &part[0], 1) != CUDA_SUCCESS || p_desc(&d_main, &rest, 1) != CUDA_SUCCESS) return DAISY_ECUDA;
    CUgreenCtx g_book = nullptr, g_main = nullptr;
    if (p_create(&g_book, d_book, dev, CU_GREEN_CTX_DEFAULT_STREAM) != CUDA_SUCCESS ||
        p_create(&g_main, d_main, dev, CU_GREEN_CTX_DEFAULT_STREAM) != CUDA_SUCCESS) {
        if (g_book) p_destroy(g_book);
        daisy_set_error("SM partitioning: cuGreenCtxCreate failed");
        return DAISY_EUNSUPPORTED;
    }
    CUstream s_book = nullptr, s_main = nullptr;
    int prio_lo = 0, prio_hi = 0;
    cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi);
    if (p_stream(&s_book, g_book, CU_STREAM_NON_BLOCKING, prio_hi) != CUDA_SUCCESS ||
        p_stream(&s_main, g_main, CU_STREAM_NON_BLOCKING, prio_lo) != CUDA_SUCCESS) {
        if (s_book) cudaStreamDestroy((cudaStream_t)s_book);
        p_destroy(g_book);
        p_destroy(g_main);
        daisy_set_error("SM partitioning: cuGreenCtxStreamCreate failed");
        return DAISY_EUNSUPPORTED;
    }
    h->part_green[0] = (void *)g_book;
    h->part_green[1] = (void *)g_main;
    h->part_book_stream = (cudaStream_t)s_book;
    h->part_main_stream = (cudaStream_t)s_main;
    h->part_book_sms = (int)part[0].sm.smCount;
    h->part_main_sms = (int)rest.sm.smCount;
    cudaEventCreateWithFlags(&h->part_ev_in, cudaEventDisableTiming);
    cudaEventCreateWithFlags(&h->part_ev_out, cudaEventDisableTiming);
    h->part_ok = 1;
    return DAISY_OK;
}

void daisy_partition_destroy(daisy_ctx *h) {
    if (!h->part_ok) return;
    decltype(&cuGreenCtxDestroy) p_destroy = nullptr;
    if (h->part_book_stream) cudaStreamDestroy(h->part_book_stream);
    if (h->part_main_stream) cudaStreamDestroy(h->part_main_stream);
    if (h->part_ev_in) cudaEventDestroy(h->part_ev_in);
    if (h->part_ev_out) cudaEventDestroy(h->part_ev_out);
    if (entry("cuGreenCtxDestroy", &p_destroy)) {
        for (int i = 0; i < 2; ++i)
            if (h->part_green[i]) p_destroy((CUgreenCtx)h->part_green[i]);
    }
    h->part_ok = 0;
}
