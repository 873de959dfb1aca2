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


@pytest.mark.skipif(not os.path.exists(_lib.LIB_PATH) or not os.path.exists(CUOBJDUMP), reason="library not built or cuobjdump missing")
def test_library_contains_tcgen05_tma_and_pair_instructions():
    # the disassembly is ~400 MB of text: filter it in a pipe instead of holding it in Python
    pattern = (r"arch = sm_[0-9a-z]+|UTC[A-Z]+MMA(\.2CTA)?|UTMALDG\.2D(\.2CTA)?|UTMASTG\.2D|LDTM\.x32|UTCBAR(\.2CTA\.MULTICAST)?|"
               r"HMMA\.16816\.F32(\.BF16)?|MOVM\.16\.MT88|UCGABAR_ARV")
    dump = subprocess.Popen([CUOBJDUMP, "-sass", _lib.LIB_PATH], stdout=subprocess.PIPE)
    hits = subprocess.run(["grep", "-oE", pattern], stdin=dump.stdout, capture_output=True, text=True).stdout.split("\n")
    dump.wait()
    ops = collections.Counter(h.strip() for h in hits if h.strip())
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
