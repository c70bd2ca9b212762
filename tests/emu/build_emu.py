"""TEST INFRASTRUCTURE ONLY: compile plain-CUDA product translation units (csrc/fmbn.cu, sgns.cu, neumf.cu, svdpp.cu, bpr_eval.cu,
sampler.cu, topk_full.cu) for the HOST against the
emulation shim of tests/emu/emu.h, so that kernels which have not run on a GPU yet can at least be executed and checked
against the oracle.  DAISY_EMU_SANITIZE=1 builds with AddressSanitizer + UBSan (run pytest with LD_PRELOAD=$(gcc -print-file-name=libasan.so)
ASAN_OPTIONS=detect_leaks=0): the memcheck that compute-sanitizer would do on the GPU pool, where it is closed;
DAISY_EMU_SANITIZE=thread builds with ThreadSanitizer (LD_PRELOAD libtsan.so): its racecheck.
The only rewrite of the source is the launch syntax:
    kernel<<<grid, block, smem, stream>>>(args);   ->   emu::launch(emu::Cfg(grid, block, smem, stream), [&] { kernel(args); });
"""
import os
import re
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
CSRC = os.path.join(ROOT, "recommend_lib_b200", "csrc")
OUT = os.path.join(HERE, "_build")

GLUE = r'''
#include <stdarg.h>
static thread_local char g_err[512];
void daisy_set_error(const char *fmt, ...) { va_list ap; va_start(ap, fmt); vsnprintf(g_err, sizeof g_err, fmt, ap); va_end(ap); }
extern "C" const char *emu_last_error() { return g_err; }
extern "C" daisy_ctx *emu_handle_dims(long long U, long long I, int D);
extern "C" daisy_ctx *emu_handle() { return emu_handle_dims(0, 0, 4); }
extern "C" daisy_ctx *emu_handle_dims(long long U, long long I, int D) {
    daisy_ctx *h = (daisy_ctx *)calloc(1, sizeof(daisy_ctx));
    h->num_sms = 4;
    h->U = U; h->I = I; h->D = D;
    h->scale = 1.0;
    h->err = (int *)calloc(2, sizeof(int));
    h->err[1] = 0x7fffffff;
    return h;
}
extern "C" int emu_err_flag(daisy_ctx *h) { return h->err[0]; }
extern "C" int emu_check(daisy_ctx *h, void *) { return h->err[0] ? DAISY_EINDEX : DAISY_OK; }   // api.cu's daisy_check, minus the stream
extern "C" int emu_err_pos(daisy_ctx *h) { return h->err[1]; }
'''


# the tensor-core filter (csrc/topk_tc.cu: tcgen05 / TMA / mbarrier PTX) is outside the emulation's reach: under it
# daisy_topk_full takes its CUDA-core filter, as it does on the device with DAISY_TOPK_TC=0
GLUE_TOPK = r'''
bool daisy_tc_supported(const daisy_ctx *) { return false; }
void daisy_tc_collect(daisy_ctx *) {}
int daisy_tc_prepare_items(daisy_ctx *, const float *, TcItems *, cudaStream_t) { return DAISY_EUNSUPPORTED; }
void daisy_tc_free_items(TcItems *, cudaStream_t) {}
int daisy_tc_filter(daisy_ctx *, const float *, const float *, const TcItems *, const int32_t *, int, float, const float *, int *,
                    unsigned long long *, int, int, cudaStream_t) { return DAISY_EUNSUPPORTED; }
'''


def rewrite(src):
    pat = re.compile(r"(\b\w+(?:<[\w\s,]+>)?)<<<(.+?)>>>\((.*?)\);", re.S)   # kernel or kernel<template args>
    out, n = pat.subn(lambda m: f"emu::launch(emu::Cfg({m.group(2)}), [&] {{ {m.group(1)}({m.group(3)}); }});", src)
    assert n > 0 and '<<<' not in out, 'a launch was not rewritten'
    out = re.sub(r"extern\s+__shared__\s+([\w ]+?)\s+(\w+)\[\];", r"\1 *\2 = (\1 *)emu::dyn_smem();", out)   # dynamic shared memory
    return out


def build(unit):
    """unit: 'fmbn' or 'sgns' -> path of the host shared library executing that unit's kernels."""
    os.makedirs(OUT, exist_ok=True)
    src_path = os.path.join(CSRC, unit + ".cu")
    so = os.path.join(OUT, f"lib{unit}_emu{'_' + os.environ['DAISY_EMU_SANITIZE'] if os.environ.get('DAISY_EMU_SANITIZE') else ''}.so")
    deps = [src_path, os.path.join(CSRC, "ctx.cuh"), os.path.join(CSRC, "topk_tc.cuh"), os.path.join(HERE, "emu.h"), os.path.join(HERE, "cub", "cub.cuh"), __file__]
    if os.path.exists(so) and all(os.path.getmtime(so) >= os.path.getmtime(d) for d in deps):
        return so
    cpp = os.path.join(OUT, unit + "_emu.cpp")
    with open(cpp, "w") as f:
        f.write(rewrite(open(src_path).read()) + GLUE + (GLUE_TOPK if unit == "topk_full" else ""))
    mode = os.environ.get("DAISY_EMU_SANITIZE", "")
    san = {"": [], "1": ["-fsanitize=address,undefined", "-fno-omit-frame-pointer", "-g"],
           "thread": ["-fsanitize=thread", "-fno-omit-frame-pointer", "-g"]}[mode]
    subprocess.check_call(["g++", "-std=c++20", "-O1", "-pthread", "-shared", "-fPIC", "-ffp-contract=off", "-w", *san,
                           "-I", HERE, "-I", CSRC, "-o", so, cpp])
    return so


if __name__ == "__main__":
    for u in ("fmbn", "sgns", "neumf", "svdpp", "bpr_eval", "sampler", "topk_full"):
        print(build(u))
