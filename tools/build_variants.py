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

VARIANTS = {
    "pad16": ["-DULTRA_META_PAD=16"],
    "pair": ["-DULTRA_META_PAIR=1"],
    "padpair": ["-DULTRA_META_PAD=16", "-DULTRA_META_PAIR=1"],
}
OUT_DIR = os.path.join(lib_build.LIB_DIR, "variants")


def build_variant(name):
    os.makedirs(OUT_DIR, exist_ok=True)
    obj = os.path.join(OUT_DIR, "rspmm_api_%s.o" % name)
    so = os.path.join(OUT_DIR, "libultra_rspmm_%s.so" % name)
    src = os.path.join(lib_build.CSRC, "rspmm_api.cu")
    subprocess.run([lib_build.NVCC] + lib_build.NVCC_FLAGS + VARIANTS[name] + ["-c", src, "-o", obj], check=True)
    others = [os.path.join(lib_build.LIB_DIR, s.replace(".cu", ".o")) for s in lib_build.SOURCES if s != "rspmm_api.cu"]
    subprocess.run([lib_build.NVCC, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", so, obj] + others,
                   check=True)
    os.remove(obj)
    return so


if __name__ == "__main__":
    lib_build.build()
    names = sys.argv[1:] or list(VARIANTS)
    with ThreadPoolExecutor(len(names)) as ex:
        for so in ex.map(build_variant, names):
            print(so)
