"""Generates tests/golden/stats.json: the per-frame statistics block the UNMODIFIED reference encoder prints with `-l 1`
(PrintFrameStatistics p64.c:1299-1332 + Statistics stat.c:52-63) for seeded synthetic clips -- the lines between
START>Frame and END>Frame.  Run where /root/reference is mounted: `python tests/golden/make_stats_golden.py`."""
import json
import os
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import oracle as O  # noqa: E402
from p64_b200 import y4m  # noqa: E402

CASES = [("qcif4_q8_tss", y4m.IT_QCIF, 4, 5, dict(q=8)),
         ("cif3_q5_full31", y4m.IT_CIF, 3, 6, dict(q=5, full_search=True, search_limit=31)),
         ("qcif5_r96000_tss", y4m.IT_QCIF, 5, 8, dict(rate=96000)),          # Buffer Contents line, varying GQUANT, overflow overrides
         ("ntsc2_q12_intra", y4m.IT_NTSC, 2, 9, dict(q=12, intra_only=True))]


def frame_blocks(log):
    """lines from 'START>Frame' to 'END>Frame' inclusive, for every frame"""
    out, cur = [], None
    for line in log.splitlines():
        if line.startswith("START>Frame"):
            cur = []
        if cur is not None and "Buffer Overflow!" not in line:      # the per-MB overflow notices (p64.c:781-782) are not part of the block
            cur.append(line.rstrip())
        if line.startswith("END>Frame") and cur is not None:
            out.append(cur); cur = None
    return out


def main():
    res = {}
    tmp = tempfile.mkdtemp()
    for name, it, nf, seed, kw in CASES:
        clip = y4m.synth_clip(it, nf, seed)
        y4m.write_y4m(f"{tmp}/c.y4m", it, clip)
        log = O.ref_encode(f"{tmp}/c.y4m", f"{tmp}/o.p64", it, nf, extra=("-l", "1"), **kw)
        res[name] = dict(image_type=it, n_frames=nf, seed=seed, args=kw, frames=frame_blocks(log))
        print(name, len(res[name]["frames"]), "frames;", res[name]["frames"][-1][-2])
    json.dump(res, open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "stats.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
