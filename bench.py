#!/usr/bin/env python
"""bench.py -- CIF frames/s through the H.261 hot path (ME + decision + prediction + Chen DCT + quantise + recon).

    python bench.py [--gpus N] [--steps K] [--warmup W]            this repo's CUDA path (one JSON line)
    python bench.py --impl reference ...                           the UNMODIFIED reference CPU encoder, all host cores

Workload (BASELINE.json configs[4], the per-GPU share): 256 independent synthetic CIF streams per GPU,
fixed quantiser 8, exhaustive +-15 motion search (`-i 31`, FastBME), no rate control.  A step = one frame time
of every stream (256 CIF frames per GPU): steady-state inter frames; the intra first frame falls in the warm-up.
Multi-GPU: streams are partitioned across ranks, no collective on the data path (weak scaling).

  value  device-resident: source frames already in HBM (a ring of distinct frame sets larger than L2), outputs left
         in HBM; timed with CUDA events on the launching stream, max over ranks.
  e2e    the reference-facing C-ABI call with HOST buffers, p64b_ctx_submit_bits()/p64b_ctx_wait_bits(): H2D of the
         step's source frames from pinned memory + the same kernels + the device-side headers/VLC + D2H of every
         stream's finished H.261 bytes, every step (3 steps in flight).  `e2e_records` is the older variant that
         downloads every macroblock record and level instead (p64b_ctx_submit/p64b_ctx_wait, host VLC not included).
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import shutil
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

STREAMS_PER_GPU = 256
QUANT = 8
SEARCH_LIMIT = 31
RING = 6                      # distinct source frame sets resident in HBM (6 x 39 MB)
CLIP_BANK = 8                 # distinct seeded clips; stream s plays clip s % 8 with a phase offset
IT_CIF = 1
# algorithmic work per unit (DESIGN.md "Measurement")
SAD_OPS_PER_CIF_FRAME = 343473 * 64          # legal candidates (me.c:212-213) x 64 packed 4-byte SADs each
MB_BYTES_INTER = 384 + 384 + 384 + 384 + 8   # source + prediction + reconstruction + int8 levels + record
# DRAM traffic per launch (dram__bytes_read.sum + dram__bytes_write.sum) and warp-instruction counts from the committed `ncu --set full`
# captures of this very workload and of the code as committed (profiles/r02_ncu_me_search_kernel_v7.txt,
# profiles/r02_ncu_mb_encode_kernel_v6.txt), bytes / instructions per launch
NCU_CAPTURE = "profiles/r02_ncu_me_search_kernel_v7.txt, profiles/r02_ncu_mb_encode_kernel_v6.txt (round 2, code as committed)"
NCU_MB_WARP_INSTR = 69_068_928     # smsp__inst_executed.sum of mb_encode_kernel v6
NCU_ME_WARP_INSTR = 229_705_424    # smsp__inst_executed.sum of me_search_kernel v7
NCU_TRAFFIC = {"me_search_kernel": 55_187_968 + 1_454_592, "mb_encode_kernel": 81_264_640 + 33_761_024}


def _clocks_sampler(stop, out, gpu_index):
    """One streaming `nvidia-smi -lms` process (the profiling recipe's clocks line); rows are appended while the
    timed legs run."""
    q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
    try:
        p = subprocess.Popen(["nvidia-smi", "-i", str(gpu_index), f"--query-gpu={q}", "--format=csv,noheader,nounits", "-lms", "20"],
                             stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
    except Exception:
        return
    try:
        for line in p.stdout:
            f = [x.strip() for x in line.strip().split(",")]
            if len(f) >= 7:
                out.append(f)
            if stop.is_set():
                break
    finally:
        p.terminate()


def _summarise_clocks(samples):
    if not samples:
        return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unsampled"]}
    sm = sorted(int(float(s[0])) for s in samples)
    reasons = []
    for i, name in enumerate(["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]):
        if any(s[3 + i].lower().startswith("active") for s in samples):
            reasons.append(name)
    return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": int(float(samples[0][1])), "reasons": reasons,
            "samples": len(samples), "power_w_max": max(float(s[2]) for s in samples)}


_BANK = None


def make_sources(streams, n_sets: int) -> np.ndarray:
    """uint8 [n_sets, len(streams), frame_bytes]: set t holds frame (phase_s + t) of global stream s's clip."""
    from p64_b200 import y4m
    global _BANK
    if _BANK is None or _BANK[0] != n_sets:
        _BANK = (n_sets, [y4m.synth_clip(IT_CIF, n_sets + CLIP_BANK, seed=1000 + b, pan=((b % 5) - 2, (b % 3) - 1))
                          for b in range(CLIP_BANK)])
    bank = _BANK[1]
    out = np.empty((n_sets, len(streams), bank[0].shape[1]), np.uint8)
    for k, s in enumerate(streams):
        clip, phase = bank[s % CLIP_BANK], (s // CLIP_BANK) % CLIP_BANK
        out[:, k] = clip[phase:phase + n_sets]
    return out


# --------------------------------------------------------------------------------------------------------------
# reference arm / cpu_baseline: the unmodified reference encoder, one process per host core
# --------------------------------------------------------------------------------------------------------------
def run_reference_sample(frames_per_proc: int = 60, cores: int | None = None, search: str = "full"):
    """-> (frames_per_s, cores, sample description). Uses oracle/_ref/p64_ref_fs (FastBME, -i 31) when built,
    else reports kind 'port' from the C oracle."""
    from oracle import oracle as O
    from p64_b200 import y4m
    cores = cores or os.cpu_count() or 1
    tmp = tempfile.mkdtemp(dir="/dev/shm" if os.path.isdir("/dev/shm") else None)
    try:
        clip = y4m.synth_clip(IT_CIF, frames_per_proc, seed=1000)
        if O.have_ref():
            exe = os.path.join(O.REF_DIR, "p64_ref_fs" if search == "full" else "p64_ref")      # FastBME toggled | stock StepBME
            for c in range(cores):
                y4m.write_y4m(f"{tmp}/c{c}.y4m", IT_CIF, clip)
            cmds = [[exe, "-y4m", "-CIF", "-a", "0", "-b", str(frames_per_proc - 1), "-q", str(QUANT), "-i", str(SEARCH_LIMIT),
                     f"{tmp}/c{c}", "-s", f"{tmp}/o{c}.p64"] for c in range(cores)]
            taskset = shutil.which("taskset")
            t0 = time.perf_counter()
            procs = [subprocess.Popen(([taskset, "-c", str(c)] if taskset else []) + cmd, stdout=subprocess.DEVNULL,
                                      stderr=subprocess.DEVNULL) for c, cmd in enumerate(cmds)]
            rcs = [p.wait() for p in procs]
            dt = time.perf_counter() - t0
            if any(rcs):
                raise RuntimeError(f"reference encoder failed: {rcs}")
            kind = "reference"
            what = (f"{cores} pinned processes of the unmodified reference (oracle/_ref/{os.path.basename(exe)}: "
                    f"{'FastBME' if search == 'full' else 'stock StepBME'}, -q {QUANT} -i {SEARCH_LIMIT}), "
                    f"each encoding its own {frames_per_proc}-frame synthetic CIF Y4M from tmpfs, incl. its VLC and file I/O")
            return cores * frames_per_proc / dt, cores, kind, what
        # no compiled reference: time the C oracle port, single thread
        enc = O.Encoder(IT_CIF)
        n = min(frames_per_proc, 20)
        t0 = time.perf_counter()
        for f in range(n):
            enc.encode_frame(clip[f], QUANT, O.ME_FULL if search == "full" else 0, SEARCH_LIMIT)
        dt = time.perf_counter() - t0
        return n / dt, 1, "port", f"C oracle (oracle/p64_oracle.c), 1 thread, {n} synthetic CIF frames, hot path only (no VLC)"
    finally:
        shutil.rmtree(tmp, ignore_errors=True)


def main_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    vals = []
    if args.warmup > 0:
        run_reference_sample(20, search=args.search)
    t_all = time.perf_counter()
    for _ in range(args.steps):
        fps, cores, kind, what = run_reference_sample(60, search=args.search)
        vals.append(fps)
    dt = time.perf_counter() - t_all
    v = float(np.mean(vals))
    line = {"impl": "reference", "metric": "CIF frames/sec encoded (ME+DCT+Q+recon)", "value": v, "unit": "frames/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3 / max(1, args.steps),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
            "config": {"workload": "CIF 352x288 4:2:0 streams, fixed quantiser 8, " +
                                   ("full-search ME +-15 (-i 31)" if args.search == "full" else "stock three-step search (StepBME)") + ", no rate control; "
                                   "bounded sample: " + what},
            "cpu_baseline": {"value": v, "unit": "frames/s", "cores": cores, "kind": kind, "sample": what},
            "e2e": {"value": v, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))
    return 0


# --------------------------------------------------------------------------------------------------------------
# this repo's arm
# --------------------------------------------------------------------------------------------------------------
def main_cuda(args):
    import torch
    import torch.distributed as dist
    from p64_b200 import _lib
    from p64_b200.encoder import DeviceContext, make_step

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    # stdout carries exactly one JSON line: whatever libraries print there (NCCL's version banner ...) goes to stderr
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; p64_b200 has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    S, K, W = STREAMS_PER_GPU, args.steps, args.warmup
    ME_MODE = 1 if args.search == "full" else 0
    ctx = DeviceContext(IT_CIF, S, device=local)
    g = ctx.geom
    nmb, fb = g["num_mb"], g["frame_bytes"]
    stream = torch.cuda.Stream()
    ctx.set_cuda_stream(stream.cuda_stream)

    n_sets = RING
    from p64_b200 import shard
    my_streams = shard.stream_range(S * world, world, rank)      # weak scaling: 256 streams per GPU, no collective
    host_sets = make_sources(my_streams, n_sets)
    # pinned host ring for the e2e leg, device ring for the resident leg
    L = _lib.lib()
    pin = L.p64b_host_alloc(host_sets.nbytes)
    if not pin:
        raise SystemExit("pinned allocation failed")
    C.memmove(pin, host_sets.ctypes.data, host_sets.nbytes)
    dev_sets = torch.from_numpy(host_sets).cuda()
    d_mbs = torch.zeros(S * nmb * 8, dtype=torch.uint8, device="cuda")
    d_lv = torch.zeros(S * nmb * 384, dtype=torch.int8, device="cuda")
    set_bytes = S * fb

    def ring(i):   # ping-pong through the ring so consecutive steps are consecutive frames of every clip
        j = i % (2 * n_sets - 2)
        return j if j < n_sets else 2 * n_sets - 2 - j

    # The reference forces a macroblock intra once it has been inter for 132 frames (p64.c:772-773); every macroblock of a
    # stream starts intra in frame 0, so frame 133, 266, ... of a stream is a whole intra frame again (about ten times the
    # bits).  That is part of a long run (`--workload config5` crosses it), but the legs here time "one inter frame of every
    # stream": before a timed leg that such a frame would fall into, filler steps move the context past it (K <= 132).
    REFRESH = 133
    frames_seen = [0]

    def step_dev(i, first=False):
        ctx.encode_frames_dev(make_step(first, QUANT, ME_MODE, SEARCH_LIMIT), dev_sets.data_ptr() + ring(i) * set_bytes,
                              d_mbs.data_ptr(), d_lv.data_ptr())
        frames_seen[0] += 1

    def avoid_refresh(n):
        """n timed steps (+ their warm-up) are about to run on `ctx`: if frame 133 k of the streams would be among them, run filler
        steps until it is behind"""
        if n >= REFRESH - 1:
            return
        while (frames_seen[0] % REFRESH) + n >= REFRESH or frames_seen[0] % REFRESH == 0:
            step_dev(frames_seen[0])

    NOUT = int(os.environ.get("P64B_BENCH_INFLIGHT", "3"))      # steps in flight (the library's pipeline depth is 3)
    pin_outs = [(L.p64b_host_alloc(S * nmb * 8), L.p64b_host_alloc(S * nmb * 384)) for _ in range(NOUT)]
    if not all(a and b for a, b in pin_outs):
        raise SystemExit("pinned allocation failed")

    def run_host_steps(i0, n):
        """n pipelined steps through the host-buffer C-ABI call (p64b_ctx_submit / p64b_ctx_wait): every step uploads
        its source frames from pinned memory and downloads every record and level."""
        tickets = []
        for j in range(n):
            if j >= NOUT:
                ctx.wait(tickets[j - NOUT])          # the output set about to be reused has landed
            om, ol = pin_outs[j % NOUT]
            tickets.append(ctx.submit(make_step(False, QUANT, ME_MODE, SEARCH_LIMIT), pin + ring(i0 + j) * set_bytes, om, ol))
        for t in tickets[-NOUT:]:
            ctx.wait(t)
        frames_seen[0] += n

    def run_bits_steps(i0, n, c=None, base=None):
        """n pipelined steps through p64b_ctx_submit_bits / p64b_ctx_wait_bits: source frames up, stream bytes down.
        -> (bytes downloaded, bytes of H.261 stream produced)"""
        c = c or ctx
        base = base or pin
        tickets, down, used = [], 0, 0
        for j in range(n):
            if j >= NOUT:
                o = c.wait_bits_raw(tickets[j - NOUT]); down += o.downloaded_bytes; used += o.total_bytes
            tickets.append(c.submit_bits(make_step(False, QUANT, ME_MODE, SEARCH_LIMIT), (i0 + j) % 32, base + ring(i0 + j) * set_bytes))
        for t in tickets[-NOUT:]:
            o = c.wait_bits_raw(t); down += o.downloaded_bytes; used += o.total_bytes
        if c is ctx:
            frames_seen[0] += n
        return down, used

    # ---- device-resident leg -----------------------------------------------------------------------------
    with torch.cuda.stream(stream):
        for i in range(W):
            step_dev(i, first=(i == 0))
        avoid_refresh(K)
        barrier()
        samples, stop = [], threading.Event()
        # one sampler per job (rank 0's GPU): eight nvidia-smi pollers on an 8-GPU box only disturb the host side
        th = threading.Thread(target=_clocks_sampler if rank == 0 else (lambda *a: None), args=(stop, samples, local), daemon=True)
        th.start()
        launches0 = ctx.launches
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for i in range(K):
            step_dev(W + i)
        e1.record(stream)
        barrier()
        launches = ctx.launches - launches0
        ms = e0.elapsed_time(e1)
        # per-kernel durations (CUDA events around every launch, on the launching stream), same K steps again
        avoid_refresh(K)
        ctx.profile(True)
        ctx.me_executed(reset=True)
        for i in range(K):
            step_dev(W + K + i)
        prof = ctx.profile_read()
        me_executed = ctx.me_executed()        # packed SAD ops the sweeps really issued in those K launches (device counter)
        ctx.profile(False)
        # ---- the same resident step INCLUDING the device-side headers + VLC (what the reference arm's number includes too)
        avoid_refresh(K + 3)
        for i in range(3):
            ctx.encode_bits_dev(make_step(False, QUANT, ME_MODE, SEARCH_LIMIT), (W + i) % 32, dev_sets.data_ptr() + ring(W + i) * set_bytes)
        barrier()
        v0, v1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        v0.record(stream)
        for i in range(K):
            ctx.encode_bits_dev(make_step(False, QUANT, ME_MODE, SEARCH_LIMIT), (W + 3 + i) % 32, dev_sets.data_ptr() + ring(W + 3 + i) * set_bytes)
        v1.record(stream)
        barrier()
        ms_vlc_resident = v0.elapsed_time(v1)
        frames_seen[0] += K + 3
        # ---- sustained: >= 3 s of the same device-resident steps back to back (does the clock hold under seconds of 80 %-ALU
        # integer load?); the clock sampler keeps running, its rows from this window are summarised separately
        sus = None
        if not args.no_sustained:
            n_sus = max(K, int(args.sustained_s * 1e3 / max(ms / K, 1e-3)))       # (seconds long: crosses the refresh frames, 1 in 133)
            s0 = len(samples)
            barrier()
            u0, u1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            u0.record(stream)
            for i in range(n_sus):
                step_dev(W + i)
                if i % 512 == 511:
                    stream.synchronize()            # (keeps the launch queue bounded; < 0.1 % of the window)
            u1.record(stream)
            barrier()
            sus = {"ms": u0.elapsed_time(u1), "steps": n_sus, "rows": samples[s0:]}
        # ---- e2e legs: host buffers through the C-ABI calls -------------------------------------------------
        avoid_refresh(K + 3)
        run_host_steps(W + 2 * K, 3)
        barrier()
        t0 = time.perf_counter()
        run_host_steps(W + 2 * K + 3, K)
        barrier()
        ms_e2e_rec = (time.perf_counter() - t0) * 1e3   # host wall clock between device-wide synchronisations (3 streams)
        avoid_refresh(K + 4)
        run_bits_steps(W + 3 * K + 3, 4)
        barrier()
        ctx.profile(True)
        t0 = time.perf_counter()
        sc0 = int(L.p64b_ctx_second_copies(ctx.h))
        bits_down, bits_used = run_bits_steps(W + 3 * K + 7, K)
        barrier()
        ms_e2e = (time.perf_counter() - t0) * 1e3
        second_copies = int(L.p64b_ctx_second_copies(ctx.h)) - sc0
        prof_bits = ctx.profile_read()
        ctx.profile(False)
        # ---- what the PCIe link gives: the step's upload alone, back to back, pinned -> HBM (the e2e legs are upload-bound)
        gb = C.c_double()
        _lib.check(L.p64b_measure_h2d(local, C.c_void_p(pin), C.c_size_t(set_bytes), 10, C.byref(gb)))
        h2d_gbs = gb.value
        # the same with every rank uploading at the same time (N > 1): what the box's host side gives all GPUs together
        h2d_conc = h2d_gbs
        if world > 1:
            barrier()
            _lib.check(L.p64b_measure_h2d(local, C.c_void_p(pin), C.c_size_t(set_bytes), 30, C.byref(gb)))
            h2d_conc = gb.value
            barrier()
        # ---- host link probes with every rank active at the same time (between barriers): uploads from ONE buffer (host
        # cache friendly), uploads cycling the ring the e2e leg reads (6 x 39 MB per rank: DRAM), downloads alone, and both
        # directions at once with the e2e leg's sizes -- what the end-to-end number is to be read against.
        down_bytes = max(int(bits_down // max(K, 1)), 1 << 16)
        down_buf = L.p64b_host_alloc(down_bytes)
        ups = (C.c_void_p * n_sets)(*[pin + k * set_bytes for k in range(n_sets)])
        link_local = []
        for sets, mode in ((1, 1), (n_sets, 1), (n_sets, 2), (n_sets, 3)):
            u, d = C.c_double(), C.c_double()
            barrier()
            _lib.check(L.p64b_measure_link(local, ups, sets, C.c_size_t(set_bytes), C.c_void_p(down_buf), C.c_size_t(down_bytes), 24, mode,
                                           C.byref(u), C.byref(d)))
            link_local += [u.value, d.value]
        barrier()
        L.p64b_host_free(down_buf)
        # ---- the headline e2e at N > 1: streams partitioned in proportion to each GPU's share of the host links
        bal = None
        if world > 1 and not args.equal_partition:
            bal = balanced_e2e_leg(L, dist, world, rank, local, S, K, n_sets, ring, ME_MODE, NOUT, barrier, link_local[2])
        # ---- attribution experiments (--experiments): the same e2e leg (a) from write-combined pinned memory, (b) with ONE
        # process driving all GPUs from N threads while the other ranks idle
        exp_local = [0.0, 0.0]
        if args.experiments:
            wc = L.p64b_host_alloc_flags(host_sets.nbytes, 1)
            if wc:
                C.memmove(wc, host_sets.ctypes.data, host_sets.nbytes)
                avoid_refresh(K + 4)
                run_bits_steps(W + 5 * K, 4, base=wc)
                barrier()
                t0 = time.perf_counter()
                run_bits_steps(W + 5 * K + 4, K, base=wc)
                barrier()
                exp_local[0] = (time.perf_counter() - t0) * 1e3
                L.p64b_host_free(wc)
            if world > 1:
                barrier()
                if rank == 0:
                    exp_local[1] = single_process_leg(world, S, K, W, host_sets, set_bytes, ring, ME_MODE, NOUT)
                barrier()
        # ---- BASELINE configs[2]: the same streams under rate control (-r), buffer model on the device ----------
        rc_line = None
        if not args.no_rate_control:
            rc_line = rate_control_leg(L, local, S, K, pin, set_bytes, ring, ME_MODE, barrier)
        extra = 0
        while rank == 0 and len(samples) < 5 and extra < 400:        # short runs: keep the same load up until the sampler has rows
            step_dev(W + extra)
            extra += 1
            if extra % 20 == 0:
                torch.cuda.synchronize()
        torch.cuda.synchronize()
        stop.set()
        th.join(timeout=2)

    sus_ms, ms_vlc_resident = shard.max_over_ranks([sus["ms"] if sus else 0.0, ms_vlc_resident], dist if world > 1 else None, device="cuda")
    rc_ms = [rc_line["ms_dev"], rc_line["ms_host"]] if rc_line else [0.0, 0.0]
    red = shard.max_over_ranks([ms, ms_e2e, ms_e2e_rec] + rc_ms + [-h2d_conc] + [-x for x in link_local] + exp_local + [bal[0] if bal else 0.0],
                               dist if world > 1 else None, device="cuda")
    ms, ms_e2e, ms_e2e_rec, rc_ms[0], rc_ms[1], neg_conc = red[:6]
    ms_bal = red[-1]
    red = red[:-1]
    bal_down = bal_used = 0
    if bal:
        t = torch.tensor([float(bal[2]), float(bal[3])], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        bal_down, bal_used = [int(x) for x in t.cpu()]
    link_min = [-x for x in red[6:6 + len(link_local)]]       # the slowest rank's rates
    exp_ms = red[6 + len(link_local):]
    link_sum = link_local
    if world > 1:
        t = torch.tensor(link_local, dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        link_sum = [float(x) for x in t.cpu()]
    h2d_conc_min = -neg_conc          # the slowest rank's upload rate while all ranks upload

    frames = world * S * K
    value = frames / (ms * 1e-3)
    e2e_value = frames / (ms_e2e * 1e-3)
    e2e_equal = None
    if bal:           # headline = the balanced partition; the equal split stays in the line next to it
        e2e_equal = {"value": e2e_value, "unit": "frames/s", "ms_per_step": ms_e2e / K, "partition": [S] * world,
                     "note": "the same leg with 256 streams on every GPU: the job waits for the GPU whose share of the host uplinks is smallest"}
        e2e_value = frames / (ms_bal * 1e-3)
    line = None
    if rank == 0:
        peak_ops, clk = C.c_double(), C.c_double()
        _lib.check(L.p64b_measure_sad_peak(local, C.byref(peak_ops), C.byref(clk)))
        me_ms = prof["me"][0] / max(1, prof["me"][1])
        mb_ms = prof["mb"][0] / max(1, prof["mb"][1])
        me_ops = (SAD_OPS_PER_CIF_FRAME if ME_MODE else 396 * 33 * 64) * S      # tss: at most 33 probes per macroblock
        mb_bytes = MB_BYTES_INTER * nmb * S
        hbm_peak, peak_src = 6650.0, "fallback"
        try:
            mp = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
            hbm_peak, peak_src = float(mp["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
        me_roof = {"kernel": "me_search_kernel", "bound": "int_issue", "achieved": me_ops / (me_ms * 1e-3) / 1e9,
                   "peak": peak_ops.value / 1e9, "unit": "G packed-SAD ops/s", "frac": (me_ops / (me_ms * 1e-3)) / peak_ops.value,
                   "traffic": NCU_TRAFFIC["me_search_kernel"], "traffic_source": NCU_CAPTURE, "avg_launch_ms": me_ms, "launches_timed": prof["me"][1],
                   "issue": {"warp_instructions_per_launch_ncu": NCU_ME_WARP_INSTR, "frac_of_issue_peak": NCU_ME_WARP_INSTR / (me_ms * 1e-3) / (148 * 4 * clk.value * 1e6)},
                   "peak_source": "measured live: VABSDIFF4.U8.ACC issue-rate probe (p64b_measure_sad_peak)",
                   "executed": {"packed_sad_ops_per_launch": me_executed / max(1, prof["me"][1]),
                                "rate_g_ops_s": me_executed / max(1, prof["me"][1]) / (me_ms * 1e-3) / 1e9,
                                "frac_of_peak": me_executed / max(1, prof["me"][1]) / (me_ms * 1e-3) / peak_ops.value,
                                "share_of_algorithmic": me_executed / max(1, prof["me"][1]) / me_ops,
                                "note": "exact elimination (ComputeError's `error >= MV` exit, me.c:122-170): a chunk of 10 x 32 candidates is dropped when all are "
                                        "strictly above the best full SAD after 4 or 8 of 16 rows; of a chunk that survives only the candidates still at or below "
                                        "it are finished (one per lane, from a list in shared memory).  `achieved`/`frac` count the ALGORITHMIC ops (every legal "
                                        "candidate in full) per the bench contract and can exceed 1 on matchable content; `executed` is what the ALU pipe really "
                                        "issued (device counter, includes masked surplus lanes)"},
                   "algorithmic": (f"{SAD_OPS_PER_CIF_FRAME} packed SAD ops per CIF frame (343473 legal candidates x 64) x {S} frames per launch" if ME_MODE
                                   else f"three-step search: at most 33 probes x 64 packed SAD ops per macroblock (upper bound) x {396 * S} macroblocks per launch")}
        mb_roof = {"kernel": "mb_encode_kernel", "bound": "hbm", "achieved": mb_bytes / (mb_ms * 1e-3) / 1e9, "peak": hbm_peak,
                   "unit": "GB/s", "frac": (mb_bytes / (mb_ms * 1e-3) / 1e9) / hbm_peak, "traffic": NCU_TRAFFIC["mb_encode_kernel"], "avg_launch_ms": mb_ms,
                   "launches_timed": prof["mb"][1], "peak_source": peak_src,
                   "issue": {"warp_instructions_per_launch_ncu": NCU_MB_WARP_INSTR, "ncu_capture": NCU_CAPTURE,
                             "frac_of_issue_peak": NCU_MB_WARP_INSTR / (mb_ms * 1e-3) / (148 * 4 * clk.value * 1e6),
                             "note": "the bound that actually holds: warp-instructions (ncu smsp__inst_executed.sum of the committed capture) per "
                                     "launch time vs 148 SMs x 4 schedulers x 1 instruction per clock"},
                   "note": "integer-issue bound, not HBM bound: ncu shows the ALU pipe (shifts, byte permutes, min/max, shift-adds) busy 58 % "
                           "and the FMA pipe (IMAD, IDP.4A) 33 % at 68 % issue utilisation, 3630 instructions per 8x8 block (DESIGN.md 3.2)",
                   "algorithmic": f"{MB_BYTES_INTER} B per inter macroblock (384 source + 384 prediction + 384 reconstruction + 384 int8 levels + 8 record) x {nmb * S} macroblocks per launch"}
        dominant = me_roof if me_ms >= mb_ms else mb_roof
        cpu = None
        if world == 1:
            fps, cores, kind, what = run_reference_sample(60, search=args.search)
            cpu = {"value": fps, "unit": "frames/s", "cores": cores, "kind": kind, "sample": what}
        line = {"metric": "CIF frames/sec encoded (ME+DCT+Q+recon)", "value": value, "unit": "frames/s", "n_gpus": world,
                "steps": K, "warmup": W, "ms_per_step": ms / K, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "u8", "data": "synthetic",
                "config": {"workload": f"{S} independent synthetic CIF 352x288 4:2:0 streams per GPU (BASELINE configs[4] per-GPU share), "
                                       f"one step = one inter frame of every stream, fixed quantiser {QUANT}, " +
                                       (f"full-search ME +-15 (-i {SEARCH_LIMIT})" if ME_MODE else "stock three-step search (StepBME)") + ", no rate control",
                           "streams_per_gpu": S, "frames_per_step": world * S, "parallelism": f"streams partitioned over {world} GPU(s), no collective",
                           "l2": f"inputs larger than L2: ring of {n_sets} source sets ({n_sets * set_bytes >> 20} MiB) + frame stores + outputs = {(n_sets + 3) * set_bytes >> 20} MiB per GPU"},
                "roofline": dominant, "roofline_kernels": {"me_search_kernel": me_roof, "mb_encode_kernel": mb_roof},
                "kernel_share_of_step": {"me_search_kernel": me_ms / (me_ms + mb_ms), "mb_encode_kernel": mb_ms / (me_ms + mb_ms)},
                "cpu_baseline": cpu,
                "value_with_vlc": {"value": frames / (ms_vlc_resident * 1e-3), "unit": "frames/s", "ms_per_step": ms_vlc_resident / K,
                                   "note": "device-resident like `value`, but the step also writes the headers and the VLC on the device (p64b_ctx_encode_bits_dev): "
                                           "like for like with the reference arm, whose number includes its VLC"},
                "e2e": {"value": e2e_value, "unit": "frames/s",
                        "h2d_bytes_per_step": S * fb if not bal else world * S * fb, "d2h_bytes_per_step": bits_down // K if not bal else bal_down // K,
                        "stream_bytes_per_step": bits_used // K if not bal else bal_used // K, "ms_per_step": (ms_bal if bal else ms_e2e) / K,
                        "bytes_are": "per GPU" if not bal else "whole job (all GPUs)",
                        "second_copies_equal_partition_leg_rank0": second_copies,
                        "partition": [S] * world if not bal else bal[1], "partition_from_link_probe": None if not bal else bal[4],
                        "partition_note": None if not bal else "streams per GPU, proportional to each GPU's measured share of the host links (probe, "
                                                               "then one trial run) -- found before the timed region; stream contents and bytes do not depend on it",
                        "equal_partition": e2e_equal,
                        "vlc_kernels_ms_per_step": prof_bits["vlc"][0] / max(1, prof_bits["vlc"][1]),
                        "upload_alone_gbs": h2d_gbs, "upload_share_of_step_equal_partition": (S * fb / (h2d_gbs * 1e9)) / (ms_e2e * 1e-3 / K),
                        "upload_all_ranks_at_once_gbs_per_gpu_min": h2d_conc_min,
                        "upload_bound_frames_s": world * h2d_conc_min * 1e9 / fb,
                        "link": {"note": "all ranks at once, GB/s per GPU [slowest rank, mean]; `ring` cycles the 6 source sets the e2e leg reads "
                                         "(DRAM-sourced), `one_buffer` repeats one set (host-cache friendly); duplex = uploads and downloads "
                                         "of the e2e leg's sizes at the same time",
                                 "upload_one_buffer": [link_min[0], link_sum[0] / world], "upload_ring": [link_min[2], link_sum[2] / world],
                                 "download_alone": [link_min[5], link_sum[5] / world],
                                 "duplex_upload": [link_min[6], link_sum[6] / world], "duplex_download": [link_min[7], link_sum[7] / world],
                                 "bound_frames_s_from_ring_upload": link_sum[2] * 1e9 / fb,
                                 "bound_frames_s_from_duplex_upload": link_sum[6] * 1e9 / fb},
                        "efficiency_vs_duplex_upload_bound": e2e_value / max(link_sum[6] * 1e9 / fb, 1.0),
                        "efficiency_vs_ring_upload_bound": e2e_value / max(link_sum[2] * 1e9 / fb, 1.0),
                        "api": "p64b_ctx_submit_bits/p64b_ctx_wait_bits (host source frames in, finished H.261 stream bytes out; "
                               "headers + VLC on the device; pinned buffers; 3 steps in flight)"},
                "e2e_records": {"value": frames / (ms_e2e_rec * 1e-3), "unit": "frames/s", "h2d_bytes_per_step": S * fb,
                                "d2h_bytes_per_step": S * nmb * (384 + 8), "ms_per_step": ms_e2e_rec / K,
                                "api": "p64b_ctx_submit/p64b_ctx_wait (records + levels out, host VLC NOT included)"},
                "gpu_launches": int(launches), "clocks": _summarise_clocks(samples)}
        if sus:
            rows = sus["rows"]
            mhz = sorted(int(float(r_[0])) for r_ in rows) if rows else []
            line["sustained"] = {"value": world * S * sus["steps"] / (sus_ms * 1e-3), "unit": "frames/s", "seconds": sus_ms * 1e-3, "steps": sus["steps"],
                                 "ms_per_step": sus_ms / sus["steps"], "vs_value": (world * S * sus["steps"] / (sus_ms * 1e-3)) / value,
                                 "sm_mhz_median": mhz[len(mhz) // 2] if mhz else None, "sm_mhz_min": mhz[0] if mhz else None,
                                 "power_w_max": max(float(r_[2]) for r_ in rows) if rows else None, "clock_samples": len(rows),
                                 "note": "the device-resident leg repeated back to back for seconds (rank 0's GPU sampled every 20 ms)"}
        if args.experiments:
            line["e2e_experiments"] = {
                "write_combined_source": {"value": frames / (exp_ms[0] * 1e-3) if exp_ms[0] else None, "unit": "frames/s"},
                "one_process_n_threads": {"value": frames / (exp_ms[1] * 1e-3) if exp_ms[1] else None, "unit": "frames/s",
                                          "note": "rank 0 alone drives all GPUs (one thread + one p64b_ctx per GPU), the other ranks idle"}}
        if rc_line:
            line["e2e_rate_control"] = {
                "value": world * S * K / (rc_ms[0] * 1e-3), "unit": "frames/s", "ms_per_step": rc_ms[0] / K,
                "h2d_bytes_per_step": S * fb, "d2h_bytes_per_step": rc_line["down"] // K, "stream_bytes_per_step": rc_line["used"] // K,
                "rate": RATE, "gquant_range_last_step": rc_line["gquant"], "overflow_mbs": rc_line["overflows"],
                "kernel_launches_per_step": rc_line["launches"] // K,
                "api": "p64b_ctx_set_rate_control + p64b_ctx_submit_bits/p64b_ctx_wait_bits (BASELINE configs[2]: -r; per-GOB GQUANT "
                       "and the overflow override chosen on the device from exact bit counts; 3 steps in flight)",
                "host_round_trip_path": {"value": world * S * rc_line["host_steps"] / (rc_ms[1] * 1e-3), "unit": "frames/s",
                                         "ms_per_step": rc_ms[1] / rc_line["host_steps"], "steps": rc_line["host_steps"],
                                         "api": "p64b_enc_encode with host_vlc=1: p64b_ctx_frame_begin / 12 x p64b_ctx_encode_gob / "
                                                "p64b_ctx_frame_end, rate control + VLC on the host cores"}}
        sys.stdout.flush()
        os.dup2(real_stdout, 1)
        print(json.dumps(line), flush=True)
    ctx.close()
    L.p64b_host_free(pin)
    for a, b in pin_outs:
        L.p64b_host_free(a); L.p64b_host_free(b)
    if world > 1:
        dist.destroy_process_group()
    return 0


def split_proportional(total, weights, floor=8):
    """integer shares of `total` proportional to `weights` (largest remainder), each at least `floor`"""
    w = np.maximum(np.asarray(weights, np.float64), 1e-9)
    raw = w / w.sum() * total
    shares = np.maximum(np.floor(raw).astype(int), floor)
    order = np.argsort(-(raw - np.floor(raw)))
    i = 0
    while shares.sum() < total:
        shares[order[i % len(order)]] += 1; i += 1
    while shares.sum() > total:
        k = int(np.argmax(shares)); shares[k] -= 1
    return [int(x) for x in shares]


def balanced_e2e_leg(L, dist, world, rank, local, S, K, n_sets, ring, me_mode, NOUT, barrier, link_rate):
    """The end-to-end leg with the job's world*S streams partitioned over the GPUs IN PROPORTION TO WHAT EACH GPU'S HOST LINK
    DELIVERS while all of them upload (the box's GPUs share PCIe uplinks unevenly: the equal split waits for the slowest one).
    Streams are independent, so any partition gives the same bytes; there is still no exchange between ranks.  The split is
    found before the timed region: from the link probe first, then corrected once from a short trial run.
    -> (ms of K steps [max over ranks is taken by the caller], shares, bytes downloaded, stream bytes, trial shares)"""
    import torch
    from p64_b200.encoder import DeviceContext, make_step

    def gather(x):
        t = torch.zeros(world, dtype=torch.float64, device="cuda")
        t[rank] = x
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return [float(v) for v in t.cpu()]

    def run(c, base, set_bytes, i0, n, first):
        tickets, down, used = [], 0, 0
        for j in range(n):
            if j >= NOUT:
                o = c.wait_bits_raw(tickets[j - NOUT]); down += o.downloaded_bytes; used += o.total_bytes
            tickets.append(c.submit_bits(make_step(first and j == 0, QUANT, me_mode, SEARCH_LIMIT), (i0 + j) % 32, base + ring(i0 + j) * set_bytes))
        for t in tickets[-NOUT:]:
            o = c.wait_bits_raw(t); down += o.downloaded_bytes; used += o.total_bytes
        return down, used

    shares = split_proportional(world * S, gather(link_rate))
    trial = list(shares)
    out = None
    for it in range(2):
        my, start = shares[rank], sum(shares[:rank])
        hs = make_sources(range(start, start + my), n_sets)
        c = DeviceContext(IT_CIF, my, device=local)
        pin = L.p64b_host_alloc(hs.nbytes)
        C.memmove(pin, hs.ctypes.data, hs.nbytes)
        sb = my * hs.shape[2]
        steps = K if it else max(16, K // 2)
        run(c, pin, sb, 0, 6, True)
        barrier()
        t0 = time.perf_counter()
        down, used = run(c, pin, sb, 6, steps, False)
        torch.cuda.synchronize()
        mine = (time.perf_counter() - t0) * 1e3
        barrier()
        ms = (time.perf_counter() - t0) * 1e3
        c.close()
        L.p64b_host_free(pin)
        if it == 0:       # per-stream cost of every rank in the trial -> corrected split
            times = gather(mine)
            shares = split_proportional(world * S, [shares[r] / max(times[r], 1e-6) for r in range(world)])
        else:
            out = (ms, shares, down, used, trial, my)
    return out


def single_process_leg(world, S, K, W, host_sets, set_bytes, ring, me_mode, NOUT):
    """The e2e leg with ONE process driving every GPU of the job: a thread + a p64b_ctx per GPU (ctypes releases the GIL in
    the calls), each with its own pinned source ring.  -> ms for K steps of all GPUs (wall clock between thread barriers)."""
    from p64_b200 import _lib
    from p64_b200.encoder import DeviceContext, make_step
    L = _lib.lib()
    ctxs, pins = [], []
    for d in range(world):
        ctxs.append(DeviceContext(IT_CIF, S, device=d))
        p = L.p64b_host_alloc(host_sets.nbytes)
        C.memmove(p, host_sets.ctypes.data, host_sets.nbytes)      # (the same content on every GPU: this leg measures the host side)
        pins.append(p)
    bar = threading.Barrier(world + 1)

    def worker(d):
        c, base = ctxs[d], pins[d]

        def run(i0, n):
            tickets = []
            for j in range(n):
                if j >= NOUT:
                    c.wait_bits_raw(tickets[j - NOUT])
                tickets.append(c.submit_bits(make_step(i0 + j == 0, QUANT, me_mode, SEARCH_LIMIT), (i0 + j) % 32, base + ring(i0 + j) * set_bytes))
            for t in tickets[-NOUT:]:
                c.wait_bits_raw(t)
        run(0, max(W, 4))
        bar.wait()
        run(max(W, 4), K)
        bar.wait()

    ths = [threading.Thread(target=worker, args=(d,)) for d in range(world)]
    for t in ths:
        t.start()
    bar.wait()
    t0 = time.perf_counter()
    bar.wait()
    ms = (time.perf_counter() - t0) * 1e3
    for t in ths:
        t.join()
    for c in ctxs:
        c.close()
    for p in pins:
        L.p64b_host_free(p)
    return ms


RATE = 384000                 # -r of BASELINE configs[2]


def rate_control_leg(L, local, S, K, pin, set_bytes, ring, me_mode, barrier):
    """BASELINE configs[2] (CIF inter encode with rate control) for the same S streams, end to end with host buffers:
    (a) buffer model on the device, (b) the per-GOB host round-trip path of the sequence encoder, a few steps."""
    from p64_b200.encoder import DeviceContext, Encoder, make_step
    iq = min(max(10000000 // RATE, 1), 31)
    ctx = DeviceContext(IT_CIF, S, device=local)
    ctx.set_rate_control(RATE)
    NOUT = int(os.environ.get("P64B_BENCH_INFLIGHT", "3"))

    def run(i0, n, first):
        tickets, down, used = [], 0, 0
        for j in range(n):
            if j >= NOUT:
                o = ctx.wait_bits_raw(tickets[j - NOUT]); down += o.downloaded_bytes; used += o.total_bytes
            tickets.append(ctx.submit_bits(make_step(first and j == 0, iq, me_mode, SEARCH_LIMIT), (i0 + j) % 32, pin + ring(i0 + j) * set_bytes))
        for t in tickets[-NOUT:]:
            o = ctx.wait_bits_raw(t); down += o.downloaded_bytes; used += o.total_bytes
        return down, used, o

    run(0, 6, True)
    barrier()
    l0 = ctx.launches
    t0 = time.perf_counter()
    down, used, o = run(6, K, False)
    barrier()
    ms_dev = (time.perf_counter() - t0) * 1e3
    launches = ctx.launches - l0
    gq = np.ctypeslib.as_array(o.gquant, (S,)); ov = np.ctypeslib.as_array(o.overflows, (S,))
    out = {"ms_dev": ms_dev, "down": int(down), "used": int(used), "gquant": [int(gq.min()), int(gq.max())],
           "overflows": int(ov.sum()), "launches": int(launches)}
    ctx.close()
    # (b) host round trips: 12 synchronous per-GOB calls per frame, VLC and rate control on the host cores
    hs = max(2, min(K, 8))
    enc = Encoder(IT_CIF, S, rate=RATE, me_mode=me_mode, search_limit=SEARCH_LIMIT, host_vlc=True, device=local)
    frames = [np.ctypeslib.as_array(C.cast(pin + ring(i) * set_bytes, C.POINTER(C.c_uint8)), (set_bytes,)) for i in range(hs + 2)]
    enc.encode(frames[0]); enc.encode(frames[1])
    barrier()
    t0 = time.perf_counter()
    for i in range(hs):
        enc.encode(frames[2 + i])
    barrier()
    out["ms_host"] = (time.perf_counter() - t0) * 1e3
    out["host_steps"] = hs
    enc.close()
    return out


def main_me1024(args):
    """BASELINE configs[3]: batched full-search +-15 SAD over 1024 synthetic CIF frame pairs resident in HBM, against the
    measured VABSDIFF4 issue peak.  Not the default workload (that is the stream batch); one JSON line of the same shape."""
    import torch
    from p64_b200 import _lib, y4m
    from p64_b200.encoder import DeviceContext
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; p64_b200 has no CPU fallback")
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    P, K, W = 1024, args.steps, args.warmup
    w, h = y4m.DIMS[IT_CIF]
    bank = [y4m.random_pair(IT_CIF, 50 + b, shift=((b * 5) % 29 - 14, (b * 7) % 23 - 11), noise=4) for b in range(16)]
    ref = np.stack([bank[p % 16][0] for p in range(P)]); cur = np.stack([bank[p % 16][1] for p in range(P)])
    ctx = DeviceContext(IT_CIF, 1)
    stream = torch.cuda.Stream()
    ctx.set_cuda_stream(stream.cuda_stream)
    r, c = torch.from_numpy(ref).cuda(), torch.from_numpy(cur).cuda()
    out = torch.zeros(P * 396 * 8, dtype=torch.int32, device="cuda")
    torch.cuda.synchronize()
    samples, stop = [], threading.Event()
    with torch.cuda.stream(stream):
        for _ in range(W):
            ctx.motion_estimation_dev(r.data_ptr(), c.data_ptr(), P, 1, SEARCH_LIMIT, out.data_ptr())
        torch.cuda.synchronize()
        th = threading.Thread(target=_clocks_sampler, args=(stop, samples, 0), daemon=True)
        th.start()
        # The pass order of the early exit takes a hint from the output array (the macroblock's previous vector).  Timed
        # twice: COLD, the array zeroed before every launch (no information: the headline), and WARM, the array left as
        # the previous launch wrote it (the same pairs again = a perfect predictor, like steady motion in a stream).
        def timed(zero_first):
            ctx.me_executed(reset=True)
            evs = []
            for _ in range(K):
                if zero_first:
                    out.zero_()
                a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a0.record(stream)
                ctx.motion_estimation_dev(r.data_ptr(), c.data_ptr(), P, 1, SEARCH_LIMIT, out.data_ptr())
                a1.record(stream)
                evs.append((a0, a1))
            torch.cuda.synchronize()
            return sum(x.elapsed_time(y) for x, y in evs), ctx.me_executed() / K
        ms, executed = timed(True)
        ms_warm, executed_warm = timed(False)
        extra = 0
        while len(samples) < 5 and extra < 2000:
            ctx.motion_estimation_dev(r.data_ptr(), c.data_ptr(), P, 1, SEARCH_LIMIT, out.data_ptr())
            extra += 1
            if extra % 20 == 0:
                torch.cuda.synchronize()
        torch.cuda.synchronize()
        stop.set()
        th.join(timeout=2)
    L = _lib.lib()
    peak_ops, clk = C.c_double(), C.c_double()
    _lib.check(L.p64b_measure_sad_peak(0, C.byref(peak_ops), C.byref(clk)))
    alg = SAD_OPS_PER_CIF_FRAME * P
    line = {"metric": "CIF frame pairs/sec, full-search ME +-15 (BASELINE configs[3])", "value": P * K / (ms * 1e-3), "unit": "frame pairs/s",
            "n_gpus": 1, "steps": K, "warmup": W, "ms_per_step": ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "u8", "data": "synthetic",
            "config": {"workload": "ME microbenchmark: 1024 CIF luma pairs resident in HBM (16 distinct seeded pairs: uniform noise and a copy "
                                   "shifted by up to +-14 with noise +-4), exhaustive search -i 31, one launch per step",
                       "l2": f"inputs larger than L2: {2 * P * w * h >> 20} MiB of luma planes per launch"},
            "roofline": {"kernel": "me_search_kernel", "bound": "int_issue", "achieved": alg / (ms / K * 1e-3) / 1e9, "peak": peak_ops.value / 1e9,
                         "unit": "G packed-SAD ops/s", "frac": alg / (ms / K * 1e-3) / peak_ops.value, "traffic": None,
                         "executed": {"packed_sad_ops_per_launch": executed, "frac_of_peak": executed / (ms / K * 1e-3) / peak_ops.value,
                                      "share_of_algorithmic": executed / alg},
                         "with_previous_vectors_as_hint": {"value": P * K / (ms_warm * 1e-3), "ms_per_step": ms_warm / K,
                                                           "share_of_algorithmic": executed_warm / alg,
                                                           "note": "output array left as the previous launch wrote it: the same pairs again, i.e. a perfect "
                                                                   "predictor for the pass order (steady motion in a stream); the headline zeroes it before every launch"},
                         "algorithmic": f"{SAD_OPS_PER_CIF_FRAME} packed SAD ops per CIF pair x {P} pairs per launch"},
            "cpu_baseline": None, "e2e": None, "gpu_launches": K, "clocks": _summarise_clocks(samples)}
    ctx.close()
    sys.stdout.flush()
    os.dup2(real_stdout, 1)
    print(json.dumps(line), flush=True)
    return 0



# --------------------------------------------------------------------------------------------------------------
# BASELINE configs[1]: QCIF intra-only (DCT + quantise + IDCT path only, no ME) on 1 B200
# --------------------------------------------------------------------------------------------------------------
IT_QCIF = 2
QCIF_STREAMS = 1024           # 1024 x 38 016 B = 37 MiB per source set (4 sets in the ring: larger than L2 with the outputs)
MB_BYTES_INTRA = 384 + 384 + 384 + 8          # source + reconstruction + int8 levels + record (no prediction)


def run_reference_intra_sample(frames_per_proc=100, cores=None):
    """the unmodified reference, `-QCIF -o -q 8 < test.intra` (every macroblock intra), one pinned process per host core"""
    from oracle import oracle as O
    from p64_b200 import y4m
    cores = cores or os.cpu_count() or 1
    if not O.have_ref():
        return None
    tmp = tempfile.mkdtemp(dir="/dev/shm" if os.path.isdir("/dev/shm") else None)
    try:
        clip = y4m.synth_clip(IT_QCIF, frames_per_proc, seed=4321)
        exe = os.path.join(O.REF_DIR, "p64_ref")
        for c in range(cores):
            y4m.write_y4m(f"{tmp}/c{c}.y4m", IT_QCIF, clip)
        taskset = shutil.which("taskset")
        t0 = time.perf_counter()
        procs = []
        for c in range(cores):
            cmd = [exe, "-y4m", "-QCIF", "-a", "0", "-b", str(frames_per_proc - 1), "-q", str(QUANT), "-o", f"{tmp}/c{c}", "-s", f"{tmp}/o{c}.p64"]
            procs.append(subprocess.Popen(([taskset, "-c", str(c)] if taskset else []) + cmd, stdin=open(os.path.join(O.REF_DIR, "test.intra"), "rb"),
                                          stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL))
        rcs = [p.wait() for p in procs]
        dt = time.perf_counter() - t0
        if any(rcs):
            raise RuntimeError(f"reference encoder failed: {rcs}")
        return {"value": cores * frames_per_proc / dt, "unit": "frames/s", "cores": cores, "kind": "reference",
                "sample": f"{cores} pinned processes of the unmodified reference (oracle/_ref/p64_ref -QCIF -o -q {QUANT} < test.intra), each "
                          f"encoding its own {frames_per_proc}-frame synthetic QCIF Y4M from tmpfs, incl. its VLC and file I/O"}
    finally:
        shutil.rmtree(tmp, ignore_errors=True)


def main_qcif_intra(args):
    import torch
    from p64_b200 import _lib, y4m
    from p64_b200.encoder import DeviceContext, make_step
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; p64_b200 has no CPU fallback")
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    S, K, W, NSETS = QCIF_STREAMS, args.steps, args.warmup, 4
    L = _lib.lib()
    ctx = DeviceContext(IT_QCIF, S)
    g = ctx.geom
    nmb, fb = g["num_mb"], g["frame_bytes"]
    stream = torch.cuda.Stream()
    ctx.set_cuda_stream(stream.cuda_stream)
    bank = [y4m.synth_clip(IT_QCIF, NSETS + 8, seed=4321 + b, pan=((b % 5) - 2, (b % 3) - 1)) for b in range(8)]
    host_sets = np.empty((NSETS, S, fb), np.uint8)
    for s_ in range(S):
        host_sets[:, s_] = bank[s_ % 8][(s_ // 8) % 8:(s_ // 8) % 8 + NSETS]
    pin = L.p64b_host_alloc(host_sets.nbytes)
    C.memmove(pin, host_sets.ctypes.data, host_sets.nbytes)
    dev_sets = torch.from_numpy(host_sets).cuda()
    d_mbs = torch.zeros(S * nmb * 8, dtype=torch.uint8, device="cuda")
    d_lv = torch.zeros(S * nmb * 384, dtype=torch.int8, device="cuda")
    set_bytes = S * fb
    step = make_step(False, QUANT, 0, 15, force_intra=True)
    samples, stop = [], threading.Event()
    with torch.cuda.stream(stream):
        for i in range(W):
            ctx.encode_frames_dev(make_step(i == 0, QUANT, 0, 15, force_intra=True), dev_sets.data_ptr() + (i % NSETS) * set_bytes, d_mbs.data_ptr(), d_lv.data_ptr())
        torch.cuda.synchronize()
        th = threading.Thread(target=_clocks_sampler, args=(stop, samples, 0), daemon=True)
        th.start()
        l0 = ctx.launches
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for i in range(K):
            ctx.encode_frames_dev(step, dev_sets.data_ptr() + ((W + i) % NSETS) * set_bytes, d_mbs.data_ptr(), d_lv.data_ptr())
        e1.record(stream)
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        launches = ctx.launches - l0
        ctx.profile(True)
        for i in range(K):
            ctx.encode_frames_dev(step, dev_sets.data_ptr() + ((W + i) % NSETS) * set_bytes, d_mbs.data_ptr(), d_lv.data_ptr())
        prof = ctx.profile_read()
        ctx.profile(False)
        # e2e: host frames in, stream bytes out
        def run_bits(i0, n):
            tickets, down, used = [], 0, 0
            for j in range(n):
                if j >= 3:
                    o = ctx.wait_bits_raw(tickets[j - 3]); down += o.downloaded_bytes; used += o.total_bytes
                tickets.append(ctx.submit_bits(step, (i0 + j) % 32, pin + ((i0 + j) % NSETS) * set_bytes))
            for t in tickets[-3:]:
                o = ctx.wait_bits_raw(t); down += o.downloaded_bytes; used += o.total_bytes
            return down, used
        run_bits(0, 4)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        down, used = run_bits(4, K)
        torch.cuda.synchronize()
        ms_e2e = (time.perf_counter() - t0) * 1e3
        extra = 0
        while len(samples) < 5 and extra < 3000:
            ctx.encode_frames_dev(step, dev_sets.data_ptr(), d_mbs.data_ptr(), d_lv.data_ptr()); extra += 1
            if extra % 50 == 0:
                torch.cuda.synchronize()
        torch.cuda.synchronize()
        stop.set(); th.join(timeout=2)
    mb_ms = prof["mb"][0] / max(1, prof["mb"][1])
    hbm_peak, peak_src = 6650.0, "fallback"
    try:
        hbm_peak, peak_src = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        pass
    mb_bytes = MB_BYTES_INTRA * nmb * S
    line = {"metric": "QCIF frames/sec encoded intra-only (DCT+Q+IDCT path, BASELINE configs[1])", "value": S * K / (ms * 1e-3), "unit": "frames/s", "n_gpus": 1,
            "steps": K, "warmup": W, "ms_per_step": ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
            "config": {"workload": f"{S} independent synthetic QCIF 176x144 4:2:0 streams, every macroblock intra (`-o < test.intra`), fixed quantiser {QUANT}, no ME; "
                                   "one step = one frame of every stream", "streams_per_gpu": S, "frames_per_step": S,
                       "l2": f"inputs larger than L2: ring of {NSETS} source sets ({NSETS * set_bytes >> 20} MiB) + frame stores + outputs = {(NSETS + 3) * set_bytes >> 20} MiB"},
            "roofline": {"kernel": "mb_encode_kernel", "bound": "hbm", "achieved": mb_bytes / (mb_ms * 1e-3) / 1e9, "peak": hbm_peak, "unit": "GB/s",
                         "frac": mb_bytes / (mb_ms * 1e-3) / 1e9 / hbm_peak, "traffic": None, "avg_launch_ms": mb_ms, "launches_timed": prof["mb"][1], "peak_source": peak_src,
                         "note": "integer-issue bound like the inter case (DESIGN.md 3.2); intra macroblocks skip the prediction fetch and the loop filter",
                         "algorithmic": f"{MB_BYTES_INTRA} B per intra macroblock (384 source + 384 reconstruction + 384 int8 levels + 8 record) x {nmb * S} macroblocks per launch"},
            "cpu_baseline": run_reference_intra_sample(),
            "e2e": {"value": S * K / (ms_e2e * 1e-3), "unit": "frames/s", "h2d_bytes_per_step": S * fb, "d2h_bytes_per_step": down // K,
                    "stream_bytes_per_step": used // K, "ms_per_step": ms_e2e / K,
                    "api": "p64b_ctx_submit_bits/p64b_ctx_wait_bits (host source frames in, finished H.261 stream bytes out; pinned buffers; 3 steps in flight)"},
            "gpu_launches": int(launches), "clocks": _summarise_clocks(samples)}
    ctx.close()
    L.p64b_host_free(pin)
    sys.stdout.flush()
    os.dup2(real_stdout, 1)
    print(json.dumps(line), flush=True)
    return 0


# --------------------------------------------------------------------------------------------------------------
# BASELINE configs[4] at full length: 256 streams per GPU x 300 frames through the sequence encoder, every stream hashed
# --------------------------------------------------------------------------------------------------------------
def main_config5(args):
    """300 CIF frames of 256 streams per GPU (2048 on 8 GPUs) through the public sequence encoder (p64b_enc_*: host frames in,
    .p64 bytes out, device-side VLC), fixed quantiser 8, exhaustive search -i 31.  Crosses the forced-intra refresh
    (LastIntra > 131, p64.c:772-773) and nine temporal-reference wraps inside a batch.  Every stream is md5'd; streams that play
    the same clip at the same phase must be byte-identical; one stream per distinct (clip, phase) pair up to `--sample` is
    byte-compared with the unmodified reference encoder (oracle/_ref/p64_ref_fs) run on that stream's frames."""
    import hashlib
    import torch
    import torch.distributed as dist
    from oracle import oracle as O
    from p64_b200 import shard, y4m
    from p64_b200.encoder import Encoder
    world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0")); local = int(os.environ.get("LOCAL_RANK", "0"))
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; p64_b200 has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    S, NF = STREAMS_PER_GPU, args.frames
    streams = list(shard.stream_range(S * world, world, rank))
    bank = [y4m.synth_clip(IT_CIF, NF + CLIP_BANK, seed=1000 + b, pan=((b % 5) - 2, (b % 3) - 1)) for b in range(CLIP_BANK)]
    keys = [(s_ % CLIP_BANK, (s_ // CLIP_BANK) % CLIP_BANK) for s_ in streams]
    enc = Encoder(IT_CIF, S, q=QUANT, me_mode=1, search_limit=SEARCH_LIMIT, device=local)
    fb = bank[0].shape[1]
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t_enc = 0.0
    t_all = time.perf_counter()
    for f in range(NF):
        frame = enc.staging()                   # the reader's role: the next frame of every stream goes straight into the encoder's pinned staging
        for k, (b, ph) in enumerate(keys):
            frame[k] = bank[b][ph + f]
        t0 = time.perf_counter()
        enc.encode(frame)                       # host frames in, 3 frames in flight
        t_enc += time.perf_counter() - t0
    t0 = time.perf_counter()
    enc.finish()
    t_enc += time.perf_counter() - t0
    wall = time.perf_counter() - t_all
    data = [enc.data(k) for k in range(S)]
    enc.close()
    md5s = [hashlib.md5(d).hexdigest() for d in data]
    by_key = {}
    consistent = True
    for k, key in enumerate(keys):
        consistent &= by_key.setdefault(key, md5s[k]) == md5s[k]
    # byte-compare a sample with the unmodified reference (one stream per distinct (clip, phase) pair, this rank's first ones)
    compared, equal = [], True
    if O.have_ref() and args.sample > 0:
        todo = list(by_key)[:args.sample]
        tmp = tempfile.mkdtemp(dir="/dev/shm" if os.path.isdir("/dev/shm") else None)
        try:
            procs = []
            for i, (b, ph) in enumerate(todo):
                y4m.write_y4m(f"{tmp}/c{i}.y4m", IT_CIF, bank[b][ph:ph + NF])
                cmd = [os.path.join(O.REF_DIR, "p64_ref_fs"), "-y4m", "-CIF", "-a", "0", "-b", str(NF - 1), "-q", str(QUANT), "-i", str(SEARCH_LIMIT), f"{tmp}/c{i}", "-s", f"{tmp}/o{i}.p64"]
                procs.append(subprocess.Popen(cmd, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL))
            for i, (key, pr) in enumerate(zip(todo, procs)):
                pr.wait()
                ref = open(f"{tmp}/o{i}.p64", "rb").read()
                k = keys.index(key)
                ok = ref == data[k]
                equal &= ok
                compared.append({"stream": streams[k], "clip": key[0], "phase": key[1], "bytes": len(ref), "md5": hashlib.md5(ref).hexdigest(), "equal": bool(ok)})
        finally:
            shutil.rmtree(tmp, ignore_errors=True)
    t_enc_max, wall_max = shard.max_over_ranks([t_enc, wall], dist if world > 1 else None, device="cuda")
    all_md5 = [md5s]
    flags = [bool(consistent), bool(equal)]
    all_cmp = [compared]
    if world > 1:
        all_md5 = [None] * world; dist.all_gather_object(all_md5, md5s)
        fl = [None] * world; dist.all_gather_object(fl, flags); flags = [all(x[0] for x in fl), all(x[1] for x in fl)]
        all_cmp = [None] * world; dist.all_gather_object(all_cmp, compared)
    if rank == 0:
        flat = [m for r_ in all_md5 for m in r_]
        frames = world * S * NF
        line = {"metric": "CIF frames/sec encoded (ME+DCT+Q+recon)", "value": frames / t_enc_max, "unit": "frames/s", "n_gpus": world, "steps": NF, "warmup": 0,
                "ms_per_step": t_enc_max * 1e3 / NF, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
                "config": {"workload": f"BASELINE configs[4] at full length: {world * S} independent synthetic CIF streams ({S} per GPU) x {NF} frames through the sequence "
                                       f"encoder p64b_enc_* (host frames in, .p64 bytes out; device-side headers + VLC), fixed quantiser {QUANT}, full-search ME +-15 (-i {SEARCH_LIMIT}); "
                                       "first frame intra, forced-intra refresh after 132 inter frames, TR wraps every 32 frames",
                           "streams_per_gpu": S, "frames_per_stream": NF, "parallelism": f"streams partitioned over {world} GPU(s), no collective"},
                "value_is": "end to end: seconds inside p64b_enc_encode/p64b_enc_finish (max over ranks); the frames are assembled in the encoder's pinned staging buffer (p64b_enc_staging, as the p64b command's Y4M reader does) -> upload -> kernels -> bytes",
                "wall_s_including_input_assembly": wall_max,
                "streams": len(flat), "md5_of_all_stream_md5s": hashlib.md5("".join(flat).encode()).hexdigest(),
                "same_input_same_bytes": flags[0],
                "byte_compared_with_reference": {"all_equal": flags[1], "streams": [c for r_ in all_cmp for c in r_],
                                                 "reference": "oracle/_ref/p64_ref_fs -y4m -CIF -q 8 -i 31 on the stream's own frames"},
                "e2e": {"value": frames / t_enc_max, "unit": "frames/s", "h2d_bytes_per_step": S * fb, "d2h_bytes_per_step": sum(len(d) for d in data) // NF},
                "cpu_baseline": None, "gpu_launches": None}
        sys.stdout.flush()
        os.dup2(real_stdout, 1)
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()
    return 0 if (flags[0] and flags[1]) else 1


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="cuda", choices=["cuda", "reference"])
    ap.add_argument("--search", default="full", choices=["full", "tss"],
                    help="full = exhaustive FastBME -i 31 (the north-star configuration, default); tss = the stock three-step StepBME")
    ap.add_argument("--workload", default="streams", choices=["streams", "me1024", "qcif_intra", "config5"],
                    help="streams = the stream batch (default, BASELINE configs[4] per-GPU share); me1024 = the ME microbenchmark of configs[3] (1 GPU); "
                         "qcif_intra = configs[1] (QCIF intra-only, 1 GPU); config5 = configs[4] at full length (300 frames per stream, every stream hashed, "
                         "a sample byte-compared with the reference encoder)")
    ap.add_argument("--frames", type=int, default=300, help="config5: frames per stream")
    ap.add_argument("--sample", type=int, default=8, help="config5: streams per rank byte-compared with the reference encoder")
    ap.add_argument("--no-rate-control", action="store_true", help="skip the extra rate-control (-r) end-to-end leg")
    ap.add_argument("--no-sustained", action="store_true", help="skip the multi-second sustained leg")
    ap.add_argument("--sustained-s", type=float, default=3.0, help="length of the sustained leg in seconds")
    ap.add_argument("--equal-partition", action="store_true", help="N > 1: keep 256 streams on every GPU in the end-to-end leg (default: balance by link share)")
    ap.add_argument("--experiments", action="store_true", help="extra end-to-end attribution legs (write-combined source, one process driving all GPUs)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "cuda" else args.warmup
    if args.impl == "cuda" and args.workload == "me1024":
        return main_me1024(args)
    if args.impl == "cuda" and args.workload == "qcif_intra":
        return main_qcif_intra(args)
    if args.impl == "cuda" and args.workload == "config5":
        return main_config5(args)
    return main_reference(args) if args.impl == "reference" else main_cuda(args)


if __name__ == "__main__":
    sys.exit(main())
