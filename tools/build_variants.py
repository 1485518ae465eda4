"""Developer A/B builds: compile-time variants of the forward kernel as separate libraries (never the default).

``python tools/build_variants.py`` writes ``ultra_b200/lib/variants/libultra_rspmm_<name>.so`` (git-ignored; travels
with the gpurun snapshot).  Load one with ``ULTRA_RSPMM_LIB=<path>`` (see ``ultra_b200/_capi.py``); ``tools/ab_bench.py``
times them against the default build in subprocesses and checks that their outputs are bit-identical.
"""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ultra_b200 import build as lib_build  # noqa: E402

# name -> {source file: extra nvcc flags}; sources not named reuse the default build's objects.
# Round-1 A/B (profiles/r01_ab_meta_variants.jsonl) made PAD=16 + PAIR=1 the default; "nopadpair" is the old build.
# Queued for the next session:
#  * forward / by-column kernels (ptxas, 64-register cap: U=4 24 B of spills, U=6 104 B with one LDL per loop iteration,
#    U=8 312 B, of which the unit-weight path - the only one ULTRA uses - holds 116 B: 3 LDL + 1 STL per loop iteration
#    next to 8 LDG.128; the non-unit-weight path stays at 4): more row gathers in flight per lane - the kernel is
#    latency-bound (long_scoreboard), so bytes in flight are what count.  U=6 does not divide the 64 staged entries,
#    so it needs the clamped single loads (PAIR=0).
#  * BF16x3 layer epilogue: where the next tile's loads are issued (ULTRA_TC3_PREFETCH, see nbfnet_update.cu): the
#    ncu source page shows the membar of fence.proxy.async waiting for the loads issued just before it.
API, TC3, PLAN = "rspmm_api.cu", "nbfnet_update.cu", "rspmm_plan.cu"
VARIANTS = {
    "nopadpair": {API: ["-DULTRA_META_PAD=0", "-DULTRA_META_PAIR=0"]},
    "unroll6": {API: ["-DULTRA_FWD_UNROLL=6", "-DULTRA_FWD_UNROLL_W=4", "-DULTRA_META_PAIR=0"]},
    "unroll8": {API: ["-DULTRA_FWD_UNROLL=8", "-DULTRA_FWD_UNROLL_W=4"]},
    "cta1024": {API: ["-DULTRA_CTA_THREADS=1024", "-DULTRA_CTAS_PER_SM=1"]},  # one relation tile per SM -> ~3x the L1
    "cta768": {API: ["-DULTRA_CTA_THREADS=768", "-DULTRA_CTAS_PER_SM=1"]},   # 24 warps, 85 registers, no spills
    "cta768u8": {API: ["-DULTRA_CTA_THREADS=768", "-DULTRA_CTAS_PER_SM=1", "-DULTRA_FWD_UNROLL=8", "-DULTRA_FWD_UNROLL_W=4"]},
    "noreuse": {API: ["-DULTRA_REUSE_G1=0"]},                         # the build before the parallel-edge gather reuse
    "reuse1": {API: ["-DULTRA_REUSE_G1=1"]},                          # reuse carried across the groups of U edges
    "relreuse": {API: ["-DULTRA_REUSE_REL=1"]},                       # + relation-row reuse (tunable plan_relation_runs=1)
    "cut256": {PLAN: ["-DULTRA_CUT_BLOCK=256"]},
    "cut1024": {PLAN: ["-DULTRA_CUT_BLOCK=1024"]},
    "pipe768": {API: ["-DULTRA_CTA_THREADS=768", "-DULTRA_CTAS_PER_SM=1", "-DULTRA_FWD_PIPE=1"]},
    "pipe512": {API: ["-DULTRA_FWD_PIPE=1"]},
    "pipe1024": {API: ["-DULTRA_CTA_THREADS=1024", "-DULTRA_CTAS_PER_SM=1", "-DULTRA_FWD_PIPE=1"]},
    "l2pf1": {TC3: ["-DULTRA_TC3_L2PF=1"]},
    "l2pf2": {TC3: ["-DULTRA_TC3_L2PF=2"]},
    "l2pf4": {TC3: ["-DULTRA_TC3_L2PF=4"]},
    "tc3pf1": {TC3: ["-DULTRA_TC3_PREFETCH=1"]},
    "tc3pf2": {TC3: ["-DULTRA_TC3_PREFETCH=2"]},
}
OUT_DIR = os.path.join(lib_build.LIB_DIR, "variants")


def flags_of(name):
    """'a+b' combines the per-source flags of variants a and b (the forward and the epilogue experiments are orthogonal)."""
    merged = {}
    for part in name.split("+"):
        for src, flags in VARIANTS[part].items():
            merged[src] = merged.get(src, []) + flags
    return merged


def build_variant(name):
    os.makedirs(OUT_DIR, exist_ok=True)
    so = os.path.join(OUT_DIR, "libultra_rspmm_%s.so" % name)
    objs, tmp = [], []
    per_source = flags_of(name)
    for src in lib_build.SOURCES:
        flags = per_source.get(src)
        if flags is None:
            objs.append(os.path.join(lib_build.LIB_DIR, src.replace(".cu", ".o")))
            continue
        obj = os.path.join(OUT_DIR, "%s_%s.o" % (src.replace(".cu", ""), name))
        subprocess.run([lib_build.NVCC] + lib_build.NVCC_FLAGS + flags + ["-c", os.path.join(lib_build.CSRC, src), "-o", obj],
                       check=True)
        objs.append(obj)
        tmp.append(obj)
    subprocess.run([lib_build.NVCC, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", so] + objs, check=True)
    for obj in tmp:
        os.remove(obj)
    return so


if __name__ == "__main__":
    lib_build.build()
    names = sys.argv[1:] or list(VARIANTS) + ["unroll8+tc3pf1", "unroll8+tc3pf2"]
    with ThreadPoolExecutor(len(names)) as ex:
        for so in ex.map(build_variant, names):
            print(so)
