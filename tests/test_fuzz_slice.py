"""A seeded, time-boxed slice of the randomised parity soak (tests/fuzz_parity.py) under `-m gpu`, so that every driver run
exercises random picture sizes / content kinds / quantisers / bit rates / search modes / chroma types / stream counts against
the CPU oracle + host bit writer and the decoder, not only the fixed cases."""
import os
import subprocess
import sys

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.mark.gpu
@pytest.mark.parametrize("seed", [20261018, 7])
def test_fuzz_slice(seed):
    r = subprocess.run([sys.executable, os.path.join(HERE, "fuzz_parity.py"), "25", str(seed)], capture_output=True, text=True, timeout=600)
    tail = (r.stdout + r.stderr)[-2000:]
    assert r.returncode == 0, tail
    assert "0 mismatches" in r.stdout, tail
    cases = int(r.stdout.rsplit("fuzz:", 1)[1].split("cases")[0])
    assert cases >= 10, tail
