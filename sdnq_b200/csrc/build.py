"""Build libsdnq_b200.so in-tree with nvcc for sm_100a.

    python sdnq_b200/csrc/build.py [--force] [--verbose]

One translation unit per kernel family, linked into sdnq_b200/libsdnq_b200.so (git-ignored; it travels to the GPU
box with the gpurun snapshot).  No torch headers: the library is a plain C-ABI CUDA shared object.
"""
import hashlib
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.dirname(HERE)
BUILD = os.path.join(HERE, "_build")
LIB = os.path.join(PKG, "libsdnq_b200.so")
SOURCES = ["capi.cu", "dequant.cu", "dequant_nd.cu", "dequant_svd.cu", "act_quant.cu", "act_quant_conv.cu", "gemv_w8a16.cu", "gemv_packed.cu", "gemm_w8a8.cu", "gemm_w4a16.cu", "svd_low.cu", "gemm_w8a8_packed.cu", "weight_quant.cu", "attention.cu"]
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-lineinfo",
    "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=hidden", "--expt-relaxed-constexpr",
    "-Xptxas", "-v" if os.environ.get("SDNQ_PTXAS_V") else "-O3",
]


def _digest(paths):
    h = hashlib.sha256()
    for p in sorted(paths):
        with open(p, "rb") as f:
            h.update(f.read())
    h.update(" ".join(FLAGS).encode())
    return h.hexdigest()


def build(force=False, verbose=False):
    os.makedirs(BUILD, exist_ok=True)
    deps = [os.path.join(HERE, f) for f in os.listdir(HERE) if f.endswith((".cu", ".cuh"))]
    deps.append(os.path.join(os.path.dirname(PKG), "include", "sdnq_b200.h"))
    stamp = os.path.join(BUILD, "stamp")
    digest = _digest(deps)
    if not force and os.path.exists(LIB) and os.path.exists(stamp) and open(stamp).read() == digest:
        return LIB

    headers = [d for d in deps if not d.endswith(".cu")]

    def compile_one(src):
        obj = os.path.join(BUILD, src.replace(".cu", ".o"))
        # per-object stamp: the source, every header and the flags
        ostamp, odigest = obj + ".stamp", _digest([os.path.join(HERE, src), *headers])
        if not force and os.path.exists(obj) and os.path.exists(ostamp) and open(ostamp).read() == odigest:
            return obj
        cmd = [NVCC, *FLAGS, "-c", os.path.join(HERE, src), "-o", obj]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if verbose or r.returncode != 0:
            sys.stderr.write(" ".join(cmd) + "\n" + r.stdout + r.stderr)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed on {src}")
        with open(ostamp, "w") as f:
            f.write(odigest)
        return obj

    with ThreadPoolExecutor(max_workers=len(SOURCES)) as ex:
        objs = list(ex.map(compile_one, SOURCES))
    cmd = [NVCC, "-shared", "-o", LIB, *objs, "-gencode", "arch=compute_100a,code=sm_100a", "-lcudart_static", "-Xlinker", "--exclude-libs,ALL"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(" ".join(cmd) + "\n" + r.stdout + r.stderr)
        raise RuntimeError("link failed")
    with open(stamp, "w") as f:
        f.write(digest)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
