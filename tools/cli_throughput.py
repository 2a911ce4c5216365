"""Single-stream command-line throughput: `p64b` (this repo) vs the unmodified reference binary, same 300-frame CIF Y4M,
same flags.  One stream is latency-bound on a GPU (396 macroblocks per launch); it is reported for completeness."""
import json
import os
import subprocess
import sys
import tempfile
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from p64_b200 import build, y4m  # noqa: E402

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tmp = tempfile.mkdtemp(dir="/dev/shm" if os.path.isdir("/dev/shm") else None)
nf = 300
clip = y4m.synth_clip(y4m.IT_CIF, nf, seed=4)
y4m.write_y4m(f"{tmp}/c.y4m", y4m.IT_CIF, clip)
cli = build.build_cli()
out = {}
for name, cmd in [("p64b_full31", [cli, "-y4m", "-CIF", "-a", "0", "-b", str(nf - 1), "-q", "8", "--me", "full", "-i", "31"]),
                  ("p64b_tss", [cli, "-y4m", "-CIF", "-a", "0", "-b", str(nf - 1), "-q", "8"]),
                  ("p64b_tss_rate384k", [cli, "-y4m", "-CIF", "-a", "0", "-b", str(nf - 1), "-r", "384000"]),
                  ("reference_full31", [f"{ROOT}/oracle/_ref/p64_ref_fs", "-y4m", "-CIF", "-a", "0", "-b", str(nf - 1), "-q", "8", "-i", "31"]),
                  ("reference_tss", [f"{ROOT}/oracle/_ref/p64_ref", "-y4m", "-CIF", "-a", "0", "-b", str(nf - 1), "-q", "8"])]:
    if not os.path.exists(cmd[0]):
        continue
    t0 = time.perf_counter()
    subprocess.run(cmd + [f"{tmp}/c", "-s", f"{tmp}/{name}.p64"], check=True, stdout=subprocess.DEVNULL)
    dt = time.perf_counter() - t0
    out[name] = {"seconds": round(dt, 3), "frames_per_s_incl_startup": round(nf / dt, 1), "bytes": os.path.getsize(f"{tmp}/{name}.p64")}
same = {a: open(f"{tmp}/p64b_{a}.p64", "rb").read() == open(f"{tmp}/reference_{a}.p64", "rb").read()
        for a in ("full31", "tss") if f"reference_{a}" in out}
print(json.dumps({"clip": f"{nf} synthetic CIF frames, q=8", "runs": out, "byte_identical_to_reference": same}))
