"""Build the host emulation of libsdnq_b200.so: the CUDA sources of every kernel family that does not need tcgen05 / TMA, compiled
UNCHANGED by g++ (prelude.h supplies host stand-ins for the device intrinsics, warp.h a lock-step warp / CTA runner that
`launch_pdl` dispatches to under SDNQ_HOST_EMU), plus emu.cpp (what capi.cu provides, and a few direct hooks for the tests).
The result exports the same C ABI entry points as the real library for those families.  Test infrastructure only."""
import hashlib
import os
import subprocess
import tempfile
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
CSRC = os.path.join(ROOT, "sdnq_b200", "csrc")
KERNEL_SOURCES = ["act_quant.cu", "act_quant_conv.cu", "dequant.cu", "dequant_nd.cu", "gemv_w8a16.cu", "gemv_packed.cu", "svd_low.cu", "weight_quant.cu"]
# entry points of the real library that live in translation units the emulator cannot run (tcgen05 / TMA GEMM)
NOT_EMULATED = {"sdnq_b200_scaled_mm", "sdnq_b200_scaled_mm_packed", "sdnq_b200_mm", "sdnq_b200_linear_w8a8",
                "sdnq_b200_linear_w8a8_fused", "sdnq_b200_linear_w8a8_workspace_bytes", "sdnq_b200_linear_w4a16", "sdnq_b200_scaled_mm_svd", "sdnq_b200_scaled_mm_grouped", "sdnq_b200_scaled_mm_ws", "sdnq_b200_scaled_mm_workspace_bytes",
                "sdnq_b200_attention", "sdnq_b200_attention_workspace_bytes", "sdnq_b200_smooth_k", "sdnq_b200_attn_colmean", "sdnq_b200_attn_quant"}


def _digest():
    h = hashlib.sha256()
    for d in (CSRC, HERE, os.path.join(ROOT, "include")):
        for f in sorted(os.listdir(d)):
            if f.endswith((".cu", ".cuh", ".h", ".cpp", ".py")):
                with open(os.path.join(d, f), "rb") as fh:
                    h.update(f.encode() + fh.read())
    return h.hexdigest()[:16]


def build() -> str:
    out_dir = os.path.join(tempfile.gettempdir(), f"sdnq_b200_emu_{os.getuid()}_{_digest()}")
    lib = os.path.join(out_dir, "libsdnq_b200_emu.so")
    if os.path.exists(lib):
        return lib
    os.makedirs(out_dir, exist_ok=True)
    cuda_inc = os.path.join(os.environ.get("CUDA_HOME", "/usr/local/cuda"), "include")
    # -fsanitize=alignment: x86 tolerates the misaligned vector loads / stores a GPU faults on; UBSan aborts on them instead
    # (uint4 / float4 / uint2 carry their CUDA alignment on the host too), so the emulated runs also check every access's alignment
    flags = ["-O1", "-std=c++17", "-fPIC", "-w", "-pthread", "-ffp-contract=off", "-fsanitize=alignment", "-fno-sanitize-recover=alignment", "-include", os.path.join(HERE, "prelude.h"), "-I", cuda_inc, "-I", HERE]

    def compile_one(src):
        obj = os.path.join(out_dir, os.path.basename(src) + ".o")
        r = subprocess.run(["g++", *flags, "-x", "c++", "-c", src, "-o", obj], capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"host build of {src} failed:\n{r.stderr[-4000:]}")
        return obj

    sources = [os.path.join(HERE, "emu.cpp")] + [os.path.join(CSRC, s) for s in KERNEL_SOURCES]
    with ThreadPoolExecutor(max_workers=len(sources)) as ex:
        objs = list(ex.map(compile_one, sources))
    # -Bsymbolic: the two CUDA runtime stubs of emu.cpp must win over a libcudart that torch has already loaded globally
    tmp = lib + f".{os.getpid()}.tmp"
    r = subprocess.run(["g++", "-shared", "-pthread", "-fsanitize=alignment", "-Wl,-Bsymbolic", "-o", tmp, *objs], capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("link of the host emulation failed:\n" + r.stderr[-4000:])
    os.replace(tmp, lib)
    return lib


if __name__ == "__main__":
    print(build())
