"""The built library really contains the Blackwell instructions the design claims (no GPU needed: cuobjdump disassembles the
sm_100a cubin).  SASS mnemonics per /opt/skills/guides/B200_PROFILING.md: tcgen05.mma -> UTC*MMA, TMA -> UTMALDG / UTMASTG,
tcgen05.ld -> LDTM, cta_group::2 -> the .2CTA forms, mma.sync -> HMMA, movmatrix -> MOVM."""
import collections
import os
import shutil
import subprocess

import pytest

from sdnq_b200 import _lib

CUOBJDUMP = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"


NVDISASM = shutil.which("nvdisasm") or "/usr/local/cuda/bin/nvdisasm"
PATTERN = (r"UTC[A-Z]+MMA(\.2CTA)?|UTMALDG\.2D(\.2CTA)?|UTMASTG\.2D|LDTM\.x32|UTCBAR(\.2CTA\.MULTICAST)?|"
           r"HMMA\.16816\.F32(\.BF16)?|MOVM\.16\.MT88|UCGABAR_ARV")


def _count_ops(cubin):
    dis = subprocess.Popen([NVDISASM, cubin], stdout=subprocess.PIPE)
    hits = subprocess.run(["grep", "-oE", PATTERN], stdin=dis.stdout, capture_output=True, text=True).stdout.split("\n")
    dis.wait()
    return collections.Counter(h.strip() for h in hits if h.strip())


@pytest.mark.skipif(not os.path.exists(_lib.LIB_PATH) or not os.path.exists(CUOBJDUMP) or not os.path.exists(NVDISASM),
                    reason="library not built or cuobjdump / nvdisasm missing")
def test_library_contains_tcgen05_tma_and_pair_instructions(tmp_path):
    # the disassembly is several hundred MB of text: the embedded cubins (one per translation unit) are extracted and disassembled in
    # parallel, each filtered in a pipe
    from concurrent.futures import ThreadPoolExecutor
    subprocess.run([CUOBJDUMP, "-xelf", "all", _lib.LIB_PATH], cwd=tmp_path, check=True, capture_output=True)
    cubins = sorted(str(p) for p in tmp_path.glob("*.cubin"))
    assert cubins, "no device code in the library"
    ops = collections.Counter({f"arch = {os.path.basename(c).split('.')[-2]}": 1 for c in cubins})
    with ThreadPoolExecutor(max_workers=min(16, len(cubins))) as pool:
        for counted in pool.map(_count_ops, cubins):
            ops.update(counted)
    assert set(k for k in ops if k.startswith("arch")) == {"arch = sm_100a"}, "the library must contain sm_100a code only"
    expect = {
        "UTCIMMA": "tcgen05.mma kind::i8 (K1 int8)", "UTCQMMA": "tcgen05.mma kind::f8f6f4 (K1 fp8)", "UTCHMMA": "tcgen05.mma kind::f16 (K3s SVD update)",
        "UTCIMMA.2CTA": "cta_group::2 int8 MMA (CTA pairs)", "UTCQMMA.2CTA": "cta_group::2 fp8 MMA", "UTCBAR.2CTA.MULTICAST": "tcgen05.commit multicast",
        "UTMALDG.2D": "TMA tile loads", "UTMALDG.2D.2CTA": "TMA loads signalling the pair leader's barrier", "UTMASTG.2D": "TMA tile stores",
        "LDTM.x32": "tcgen05.ld (TMEM -> registers)", "HMMA.16816.F32.BF16": "mma.sync bf16 (tensor-core Hadamard, GEMV)",
        "HMMA.16816.F32": "mma.sync f16", "MOVM.16.MT88": "movmatrix (Hadamard transpose)", "UCGABAR_ARV": "cluster barrier (pair kernels)",
    }
    missing = {k: v for k, v in expect.items() if ops[k] == 0}
    assert not missing, f"expected SASS instructions absent from libsdnq_b200.so: {missing}"
