"""Design tool (CPU, NumPy): how many packed-SAD rows would different EXACT pruning schedules of the exhaustive search
execute on the bench clip (reference = the oracle's reconstruction at q = 8)?  Prints executed / algorithmic per scheme.

    python tools/me_prune_sim.py [clip_bank_index] [frame]
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import oracle as O          # noqa: E402
from p64_b200 import y4m                # noqa: E402

IT = y4m.IT_CIF
W, H = y4m.DIMS[IT]


def partial_rows(ref, cur, bx, by):
    """-> P [31(dy)][31(dx)][16] cumulative row SADs, legal mask [31][31] (FastBME -i 31 range [-15,14] + me.c:212-213)"""
    x0, y0 = bx * 16, by * 16
    pad = np.zeros((H + 62, W + 62), np.int32)
    pad[31:31 + H, 31:31 + W] = ref
    win = pad[y0 + 31 - 15:y0 + 31 + 31, x0 + 31 - 15:x0 + 31 + 31]        # 46 x 46
    v = np.lib.stride_tricks.sliding_window_view(win, (16, 16))              # [31][31][16][16]
    c = cur[y0:y0 + 16, x0:x0 + 16].astype(np.int32)
    rows = np.abs(v - c).sum(axis=3)                                         # [31][31][16]
    P = np.cumsum(rows, axis=2)
    d = np.arange(-15, 16)
    okx = (x0 + d >= 0) & (x0 + d < W - 16) & (d <= 14)
    oky = (y0 + d >= 0) & (y0 + d < H - 16) & (d <= 14)
    legal = oky[:, None] & okx[None, :]
    return P, legal


def main():
    bank = int(sys.argv[1]) if len(sys.argv) > 1 else 0
    fno = int(sys.argv[2]) if len(sys.argv) > 2 else 4
    clip = y4m.synth_clip(IT, fno + 2, seed=1000 + bank, pan=((bank % 5) - 2, (bank % 3) - 1))
    enc = O.Encoder(IT)
    prev_me = None
    for f in range(fno):
        enc.encode_frame(clip[f], 8, O.ME_FULL, 31)
        if f:
            prev_me = enc.me_records().copy()
    ref = enc.recon()[:W * H].reshape(H, W).astype(np.int32)
    cur = clip[fno][:W * H].reshape(H, W).astype(np.int32)
    enc.encode_frame(clip[fno], 8, O.ME_FULL, 31)
    me = enc.me_records()
    tot_alg = 0
    acc = {}

    def add(k, v):
        acc[k] = acc.get(k, 0) + v

    best_hist = []
    for by in range(H // 16):
        for bx in range(W // 16):
            P, legal = partial_rows(ref, cur, bx, by)
            n = by * (W // 16) + bx
            full = P[:, :, 15]
            omv = full[15, 15]
            nleg = int(legal.sum())
            tot_alg += nleg * 16
            fl = np.where(legal, full, 1 << 30)
            best = min(int(fl.min()), int(omv))
            best_hist.append(best)
            hx, hy = (int(prev_me[n][0]), int(prev_me[n][1])) if prev_me is not None else (0, 0)
            hint = int(full[hy + 15, hx + 15]) if legal[hy + 15, hx + 15] else 1 << 30
            # ideal per-candidate pruning with the final best as the bound from the start: rows until partial > best
            need = (P <= best).sum(axis=2)                 # rows a candidate survives; executes min(need+1,16) rows
            rows_ideal = np.minimum(need + 1, 16)
            add("ideal(best known)", int(rows_ideal[legal].sum()))
            b0 = min(int(omv), hint)
            need0 = np.minimum((P <= b0).sum(axis=2) + 1, 16)
            add("ideal(bound=min(omv,hint))", int(need0[legal].sum()))
            # scheme A(R, checks): all candidates R rows; bound = min(omv, hint, full SAD of the argmin-partial candidate);
            # then per-candidate survivors continue in segments with checks at the given rows (per-candidate granularity)
            for R in (4, 5, 6, 7, 8):
                pr = np.where(legal, P[:, :, R - 1], 1 << 30)
                am = np.unravel_index(int(pr.argmin()), pr.shape)
                bound = min(b0, int(full[am]))
                ex = nleg * R
                surv = legal & (P[:, :, R - 1] <= bound)
                add(f"A{R}:survivors", int(surv.sum()))
                for nxt in ((R + 2, R + 4, 16) if R <= 10 else (16,)):
                    pass
                # survivors: sparse stage with checks every 2 rows (lane per survivor; batch of 32 leaves when all are hopeless:
                # approximated per candidate)
                r = R
                s = surv.copy()
                while r < 16 and s.any():
                    r2 = min(r + 2, 16)
                    ex += int(s.sum()) * (r2 - r)
                    s = s & (P[:, :, r2 - 1] <= bound)
                    r = r2
                add(f"A{R}:per-candidate(2-row checks)", ex)
                # chunk-granular dense continuation (3 chunks of 10 dy): a chunk continues while any candidate survives, check every 2 rows
                ex2 = nleg * R
                for c0 in range(0, 30, 10):
                    sl = slice(c0, c0 + 10)
                    lg = legal[sl]
                    if not lg.any():
                        continue
                    r = R
                    alive = (lg & (P[sl, :, R - 1] <= bound)).any()
                    while r < 16 and alive:
                        r2 = min(r + 2, 16)
                        ex2 += int(lg.sum()) * (r2 - r)
                        alive = (lg & (P[sl, :, r2 - 1] <= bound)).any()
                        r = r2
                add(f"A{R}:chunk-dense(2-row checks)", ex2)
            # current scheme: passes of 10 dy (31 dx), from the hinted pass outwards, checks after 4 and 8 rows
            ylo = int(np.argmax(legal.any(axis=1)))
            nd = int(legal.any(axis=1).sum())
            if legal.any(axis=0).sum() >= 17:
                passes = [(ylo + 10 * k, min(ylo + 10 * k + 10, ylo + nd)) for k in range((nd + 9) // 10)]
                kc = min(max((hy + 15 - ylo) // 10, 0), len(passes) - 1)
                order = [kc]
                for v in range(1, 2 * len(passes)):
                    k = kc + ((v + 1) >> 1) * (-1 if v & 1 else 1)
                    if 0 <= k < len(passes):
                        order.append(k)
                bound = int(omv)
                ex = 0
                for k in order:
                    a0, a1 = passes[k]
                    lg = legal[a0:a1]
                    ncand = int(lg.sum())
                    p4 = np.where(lg, P[a0:a1, :, 3], 1 << 30).min()
                    if p4 > bound:
                        ex += ncand * 4
                        continue
                    p8 = np.where(lg, P[a0:a1, :, 7], 1 << 30).min()
                    if p8 > bound:
                        ex += ncand * 8
                        continue
                    ex += ncand * 16
                    bound = min(bound, int(np.where(lg, full[a0:a1], 1 << 30).min()))
                add("current(v6)", ex)
                # option C: like v6, but a pass that survives its last check continues with its per-candidate survivors only
                for (c1, c2) in ((4, 8), (4, 7), (5, 8), (3, 6), (4, 6), (6, 6), (5, 7), (6, 8)):
                    bound = int(omv); ex = 0; nsurv = 0
                    for k in order:
                        a0, a1 = passes[k]
                        lg = legal[a0:a1]; ncand = int(lg.sum())
                        if c1 < c2:
                            if np.where(lg, P[a0:a1, :, c1 - 1], 1 << 30).min() > bound:
                                ex += ncand * c1; continue
                        if np.where(lg, P[a0:a1, :, c2 - 1], 1 << 30).min() > bound:
                            ex += ncand * c2; continue
                        ex += ncand * c2
                        # bound from the best partial candidate's full SAD
                        pr = np.where(lg, P[a0:a1, :, c2 - 1], 1 << 30)
                        am = np.unravel_index(int(pr.argmin()), pr.shape)
                        bound = min(bound, int(full[a0:a1][am]))
                        sv = lg & (P[a0:a1, :, c2 - 1] <= bound)
                        nsurv += int(sv.sum())
                        ex += int(sv.sum()) * (16 - c2)
                        bound = min(bound, int(np.where(sv, full[a0:a1], 1 << 30).min()))
                    add(f"C({c1},{c2})", ex); add(f"C({c1},{c2}):survivors", nsurv)
            else:
                add("current(v6)", nleg * 16)
                for (c1, c2) in ((4, 8), (4, 7), (5, 8), (3, 6), (4, 6), (6, 6), (5, 7), (6, 8)):
                    add(f"C({c1},{c2})", nleg * 16)
    print(f"bank {bank} frame {fno}: best SAD median {int(np.median(best_hist))}, mean {np.mean(best_hist):.0f}; mean ME val {me[:, 2].mean():.0f}")
    for k, v in acc.items():
        if "survivors" in k:
            print(f"  {k:42s} {v / 396:.1f} per MB")
        else:
            print(f"  {k:42s} {v / tot_alg:.3f}")


if __name__ == "__main__":
    main()
