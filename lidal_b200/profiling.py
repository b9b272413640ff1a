"""Live roofline accounting for bench.py (CUDA events on the launching stream; nothing here runs under a profiler)."""
from __future__ import annotations

import json
import os

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
FALLBACK = {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}   # B200_PROFILING.md fallback


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        d["source"] = "measured (MEASURED_PEAKS.json)"
        return d
    return dict(FALLBACK, source="fallback (B200_PROFILING.md)")


class ConvTrace:
    """Installed as functional.CONV_TRACE: records one (start, end) event pair per lb_conv_fwd launch."""

    def __init__(self):
        self.records = []

    def begin(self):
        ev = torch.cuda.Event(enable_timing=True)
        ev.record()
        return ev

    def end(self, ev0, nbr, n_out, k_vol, c_in, c_out, tc):
        ev1 = torch.cuda.Event(enable_timing=True)
        ev1.record()
        self.records.append((ev0, ev1, nbr, int(n_out), k_vol, c_in, c_out, bool(tc)))

    def summarise(self):
        torch.cuda.synchronize()
        rows = []
        pair_cache = {}
        for ev0, ev1, nbr, n_out, k, cin, cout, tc in self.records:
            if nbr is None:
                pairs = n_out
            else:
                key = (nbr.data_ptr(), tuple(nbr.shape), n_out)
                if key not in pair_cache:
                    pair_cache[key] = int((nbr[:, :n_out] >= 0).sum().item())
                pairs = pair_cache[key]
            rows.append(dict(ms=ev0.elapsed_time(ev1), pairs=pairs, n_out=n_out, k=k, cin=cin, cout=cout, tc=tc,
                             flops=2.0 * pairs * cin * cout, dense_flops=2.0 * n_out * k * cin * cout))
        return rows


def conv_roofline(run, resident, steps=3):
    """Events around every sparse-conv launch for ``steps`` steps.  achieved = algorithmic FLOPs (2 * map pairs * Cin *
    Cout, summed over the tcgen05 launches) / their summed device time."""
    from .compat.nn import functional as F
    peaks = measured_peaks()
    run(*resident[0])
    torch.cuda.synchronize()
    trace = ConvTrace()
    F.CONV_TRACE = trace
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    try:
        e0.record()
        for i in range(steps):
            run(*resident[i % len(resident)])
        e1.record()
    finally:
        F.CONV_TRACE = None
    rows = trace.summarise()
    step_ms = e0.elapsed_time(e1) / steps
    tc = [r for r in rows if r["tc"]]
    simt = [r for r in rows if not r["tc"]]
    tc_ms, tc_fl = sum(r["ms"] for r in tc), sum(r["flops"] for r in tc)
    peak = float(peaks.get("bf16_tflops_sustained", FALLBACK["bf16_tflops_sustained"]))
    achieved = tc_fl / (tc_ms * 1e-3) / 1e12 if tc_ms > 0 else 0.0
    best = max(tc, key=lambda r: r["flops"] / max(r["ms"], 1e-6)) if tc else None
    traffic, traffic_src = None, None
    for name in ("r02_conv_traffic.json", "r01_conv_traffic.json"):       # dram__bytes_read.sum + dram__bytes_write.sum per launch
        tpath = os.path.join(ROOT, "profiles", name)                        # from a committed ncu pass of the bench command
        if os.path.exists(tpath):
            tj = json.load(open(tpath))
            traffic, traffic_src = tj["traffic_bytes_per_launch"], f"profiles/{name} (ncu, per launch)"
            break
    alg_bytes = sum(2.0 * (r["n_out"] * r["cin"] + r["n_out"] * r["cout"]) + 4.0 * r["k"] * r["n_out"] + 2.0 * r["k"] * r["cin"] * r["cout"]
                    for r in tc) / max(len(tc), 1)
    if os.environ.get("LIDAL_LAYER_TABLE"):
        import sys
        per = len(rows) // steps
        print(f"{'k':>3} {'cin':>4} {'cout':>4} {'n_out':>8} {'pairs':>9} {'dens':>5} {'ms':>7} {'alg TF/s':>8} {'dense TF/s':>10} tc", file=sys.stderr)
        for r in rows[:per]:
            print(f"{r['k']:3d} {r['cin']:4d} {r['cout']:4d} {r['n_out']:8d} {r['pairs']:9d} {r['pairs'] / max(r['n_out'] * r['k'], 1):5.2f} "
                  f"{r['ms']:7.3f} {r['flops'] / r['ms'] / 1e9:8.1f} {r['dense_flops'] / r['ms'] / 1e9:10.1f} {int(r['tc'])}", file=sys.stderr)
    return {
        "bound": "tensor", "kernel": "lb::conv_tc_kernel (tcgen05 implicit-GEMM sparse conv)",
        "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak, "traffic": traffic,
        "traffic_source": traffic_src, "algorithmic_bytes_per_launch": alg_bytes,
        "peak_source": peaks["source"] + ", sustained bf16 (kernel timed inside a long step)",
        "launches_per_step": len(tc) / steps, "avg_launch_ms": tc_ms / max(len(tc), 1),
        "algorithmic_gflop_per_step": tc_fl / steps / 1e9,
        "dense_gflop_per_step": sum(r["dense_flops"] for r in tc) / steps / 1e9,
        "kernel_ms_per_step": tc_ms / steps, "share_of_step": (tc_ms / steps) / step_ms,
        "simt_conv_ms_per_step": sum(r["ms"] for r in simt) / steps, "instrumented_step_ms": step_ms,
        "best_layer": None if best is None else {k: best[k] for k in ("cin", "cout", "k", "n_out", "pairs", "ms")}
        | {"tflops": best["flops"] / (best["ms"] * 1e-3) / 1e12},
        "how": "CUDA events on the launching stream around each conv launch; algorithmic FLOPs = 2*pairs*Cin*Cout from the live kernel maps",
    }


def stage_rooflines(run, resident, steps=3):
    """``roofline_by_stage``: achieved HBM GB/s of the non-conv stages of a step (map build, voxelize, devoxelize, point
    queries) = algorithmic bytes / CUDA-event time on the launching stream; peak = measured copy
    bandwidth."""
    from . import engine
    hbm = float(measured_peaks().get("hbm_gbs", FALLBACK["hbm_gbs"]))
    run(*resident[0])
    torch.cuda.synchronize()
    engine.STAGE_TRACE = engine.StageTrace()
    try:
        for i in range(steps):
            run(*resident[i % len(resident)])
        summ = engine.STAGE_TRACE.summarise()
    finally:
        engine.STAGE_TRACE = None
    out = {}
    for name, d in summ.items():
        gbs = d["bytes"] / (d["ms"] * 1e-3) / 1e9 if d["ms"] > 0 else 0.0
        out[name] = {"bound": "hbm", "ms_per_step": d["ms"] / steps, "launch_groups_per_step": d["calls"] / steps,
                     "algorithmic_bytes_per_step": d["bytes"] / steps, "achieved": gbs, "peak": hbm, "unit": "GB/s", "frac": gbs / hbm}
    return out


def scoring_extras(ms_per_step: float, dev, n_frames: int = 25, n_cls: int = 19, engine=None, kind: str = "SK"):
    """Rooflines of the scoring-side kernels on one SK/NU-shaped frame against a resident 24-frame window: achieved HBM GB/s
    (algorithmic bytes of SURVEY.md section 8d / CUDA-event time) against the measured copy peak.  The end-to-end frames/s
    number is measured by bench.py's LiDAL workload (``lidal``), not here."""
    from . import score, synth
    peaks = measured_peaks()
    seq = synth.GpuSequence(n_frames, kind, seed=77, device=dev)
    sc = score.SequenceScorer(dev)
    ev = lambda: torch.cuda.Event(enable_timing=True)   # noqa: E731
    frames = [seq.frame(i) for i in range(n_frames)]
    xyz = [score.register_points(fr[0], fr[1]) for fr in frames]
    probs = [torch.softmax(torch.sin(x.float() @ torch.randn(3, n_cls, device=dev, generator=torch.Generator(device=dev).manual_seed(5)) * 0.35) * 2.0
                           + 0.6 * torch.randn(x.shape[0], n_cls, device=dev, generator=torch.Generator(device=dev).manual_seed(300 + i)), 1)
             for i, x in enumerate(xyz)]
    torch.cuda.synchronize()
    g0, g1 = ev(), ev()
    g0.record()
    for i in range(n_frames):
        sc.add_frame(xyz[i], probs[i], frames[i][2], frames[i][3])
    g1.record()
    fid = n_frames // 2
    sc.score_frame_device(fid)
    torch.cuda.synchronize()
    reps = 10
    s0, s1 = ev(), ev()
    s0.record()
    for _ in range(reps):
        sc.score_frame_device(fid)
    s1.record()
    torch.cuda.synchronize()
    score_ms = s0.elapsed_time(s1) / reps
    _, _, cnt = sc.score_points(fid)
    matched = int(cnt.sum().item())
    npts = sc.frames[fid].n
    nn_pts = sum(sc.frames[n].n for n in score.neighbour_ids(fid, n_frames))
    # query prob once + per neighbour frame its coordinates + the matched prob rows + outputs
    alg_bytes = npts * n_cls * 4 + npts * 24 + nn_pts * 24 + matched * n_cls * 4 + npts * 12
    nv = 8 * 92000
    logits = torch.randn(nv, n_cls, device=dev)
    inv = torch.cat([torch.randint(0, 92000, (npts,), device=dev) + v * 92000 for v in range(8)])
    score.tta_tail(logits, inv, 8)
    t0, t1 = ev(), ev()
    t0.record()
    for _ in range(reps):
        score.tta_tail(logits, inv, 8)
    t1.record()
    torch.cuda.synchronize()
    tail_ms = t0.elapsed_time(t1) / reps
    tail_bytes = 8 * npts * (n_cls * 4 + 8) + npts * (n_cls * 4 + 8)
    grid_ms = g0.elapsed_time(g1) / n_frames
    grid_bytes = npts * (24 + 40 + 16 * 2)                 # read xyz, write the sorted copies + slots (2 slots per point)
    hbm = float(peaks.get("hbm_gbs", FALLBACK["hbm_gbs"]))
    roof = lambda b, ms: {"bound": "hbm", "ms": ms, "achieved": b / (ms * 1e-3) / 1e9, "peak": hbm, "unit": "GB/s",   # noqa: E731
                          "frac": b / (ms * 1e-3) / 1e9 / hbm, "algorithmic_bytes": b}
    return {
        "tta_tail_ms": tail_ms, "interframe_score_ms": score_ms, "frame_grid_build_ms": grid_ms, "points_per_frame": npts,
        "matched_pairs": matched,
        "score_roofline": roof(alg_bytes, score_ms), "tta_roofline": roof(tail_bytes, tail_ms), "grid_roofline": roof(grid_bytes, grid_ms),
        "note": "one query frame against a resident 24-frame window; scoring computes exact float64 distances and double-precision "
                "log by design (bit-exact matches), so it is issue-bound well below the HBM roof",
    }
