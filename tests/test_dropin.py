"""The drop-in claim, literally: the reference's OWN program (its main(), flag parsing, Y4M reader and stream writer,
built by oracle/build_ref.sh from the sources under /root/reference) with the body of p64EncodeFrame() replaced by one call
into libp64b200.so per frame (examples/p64gpu_dropin.c) must write the same .p64 bytes as the unmodified reference
(tests/golden/streams.json).  oracle/_ref/p64_gpu = stock three-step search, p64_gpu_fs = FastBME."""
import hashlib
import os
import subprocess

import pytest

from helpers import GOLDEN, golden_clip
from oracle import oracle as O
from p64_b200 import y4m

EXE = {False: os.path.join(O.REF_DIR, "p64_gpu"), True: os.path.join(O.REF_DIR, "p64_gpu_fs")}
# the reference-shaped binding of INTEGRATION.md section 3 (examples/p64gpu_glue.c): the reference's own VLC, headers and rate
# control consume the device's per-GOB records and levels
EXE_HV = {False: os.path.join(O.REF_DIR, "p64_gpu_hv"), True: os.path.join(O.REF_DIR, "p64_gpu_hv_fs")}
CASES = [n for n, g in GOLDEN.items() if not g["args"].get("intra_only") and n != "qcif140_q10_tss"]


def _cmd(g, prefix, out, exe=None):
    a = g["args"]
    cmd = [(exe or EXE)[bool(a.get("full_search"))], "-y4m", O.FLAG[g["image_type"]], "-a", str(a.get("start", 0)),
           "-b", str(a.get("last", g["n_frames"] - 1))]
    if a.get("frame_skip"):
        cmd += ["-k", str(a["frame_skip"])]
    if a.get("frame_rate"):
        cmd += ["-f", str(a["frame_rate"])]
    if a.get("q"):
        cmd += ["-q", str(a["q"])]
    if a.get("rate"):
        cmd += ["-r", str(a["rate"])]
    if a.get("search_limit"):
        cmd += ["-i", str(a["search_limit"])]
    return cmd + [prefix, "-s", out]


def test_dropin_binary_fails_loudly_without_a_gpu(tmp_path):
    import torch
    if torch.cuda.is_available() or not os.path.exists(EXE[False]):
        pytest.skip("needs the drop-in binary and no GPU")
    g, clip = golden_clip("qcif12_q8_tss")
    y4m.write_y4m(str(tmp_path / "c.y4m"), g["image_type"], clip)
    r = subprocess.run(_cmd(g, str(tmp_path / "c"), str(tmp_path / "o.p64")), stdout=subprocess.PIPE, stderr=subprocess.STDOUT)
    assert r.returncode != 0 and b"no CPU fallback" in r.stdout


@pytest.mark.gpu
@pytest.mark.parametrize("name", CASES)
def test_reference_program_with_the_library_dropped_in_writes_the_reference_bytes(name, tmp_path):
    if not os.path.exists(EXE[False]):
        pytest.skip("oracle/_ref/p64_gpu not built (needs the reference tree at build time)")
    g, clip = golden_clip(name)
    chroma = g["args"].get("chroma", "420jpeg")
    if g["args"].get("start"):                                        # the file holds the StartFrame frames that -a skips
        clip = y4m.synth_payloads(g["image_type"], g["n_frames"] + g["args"]["start"], g["seed"], chroma)
    y4m.write_y4m(str(tmp_path / "c.y4m"), g["image_type"], clip, chroma=chroma)     # other chroma types: the reference's reader converts
    r = subprocess.run(_cmd(g, str(tmp_path / "c"), str(tmp_path / "o.p64")), stdout=subprocess.PIPE, stderr=subprocess.STDOUT)
    assert r.returncode == 0, r.stdout.decode(errors="replace")[-500:]
    data = open(tmp_path / "o.p64", "rb").read()
    assert len(data) == g["size"] and hashlib.md5(data).hexdigest() == g["md5"]
    assert f"Number of buffer overflows: {g['overflows']}" in r.stdout.decode(errors="replace")


@pytest.mark.gpu
@pytest.mark.parametrize("name", CASES)
def test_reference_program_with_its_own_vlc_fed_from_device_records(name, tmp_path):
    """INTEGRATION.md section 3, executed: WriteMBHeader / EncodeDC / EncodeAC / CBPEncodeAC, ExecuteQuantization and the buffer
    overflow branch are the REFERENCE's own code (p64.c:692-786, 922-961, marker.c, codec.c); the device supplies the motion
    vectors (MeX.. arrays), and per GOB the macroblock records and levels.  Same bytes as the unmodified reference, incl. -r."""
    if not os.path.exists(EXE_HV[False]):
        pytest.skip("oracle/_ref/p64_gpu_hv not built (needs the reference tree at build time)")
    g, clip = golden_clip(name)
    chroma = g["args"].get("chroma", "420jpeg")
    if g["args"].get("start"):
        clip = y4m.synth_payloads(g["image_type"], g["n_frames"] + g["args"]["start"], g["seed"], chroma)
    y4m.write_y4m(str(tmp_path / "c.y4m"), g["image_type"], clip, chroma=chroma)
    r = subprocess.run(_cmd(g, str(tmp_path / "c"), str(tmp_path / "o.p64"), EXE_HV), stdout=subprocess.PIPE, stderr=subprocess.STDOUT)
    assert r.returncode == 0, r.stdout.decode(errors="replace")[-500:]
    data = open(tmp_path / "o.p64", "rb").read()
    assert len(data) == g["size"] and hashlib.md5(data).hexdigest() == g["md5"]
    out = r.stdout.decode(errors="replace")
    assert f"Number of buffer overflows: {g['overflows']}" in out
    assert out.count("Buffer Overflow!") == g["overflows"]          # the reference's own overflow branch fired, macroblock by macroblock
