#!/usr/bin/env python
"""denovo3D candidates/sec (solve + score) on N B200s vs the reference CPU path.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

Workload (BASELINE.json configs[1]): one synthetic amyloid-like filament image,
256x256 px at 1.3 A/px (true twist -1.2 deg, rise 4.75 A); dense 1000 twist x 50
rise grid; interpolation "nn", algorithm {"model": "lsq"}, cosine score.  A
"step" = one batch of consecutive grid candidates (twist-major, default 4
twists x 50 rises = 200 candidates per GPU); successive steps take successive
batches, so the inputs of every step are new and far larger than L2.
Ranks split the candidates with no data-path collective (weak scaling); rank
results are all-gathered over NCCL at the end of every step (score tile +
local top-K).

JSON line keys follow the driver contract (see the task statement / DESIGN.md).
"""

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

if os.environ.get("NCCL_DEBUG", "").upper() == "VERSION":
    os.environ["NCCL_DEBUG"] = "WARN"  # NCCL prints its version banner on stdout; rank 0's stdout is ONE JSON line

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

APIX = 1.3
N_IMG = 256
TRUE_TWIST, TRUE_RISE = -1.2, 4.75
N_TWIST, N_RISE = 1000, 50
TWISTS = np.linspace(-3.0, -0.2, N_TWIST)
RISES = np.linspace(4.4, 5.1, N_RISE)
WORKLOAD = ("cfg2: synthetic amyloid-like filament 256x256 px @1.3 A/px (true twist -1.2 deg, rise 4.75 A), "
            "1000 twist x 50 rise grid, nn interpolation, model=lsq, cosine score")


def synthetic_filament(n=N_IMG, apix=APIX, twist=TRUE_TWIST, rise=TRUE_RISE, seed=7, n_atoms=30, diameter=100.0,
                       ball_radius=3.0):
    """Helical assembly of Gaussian balls projected along y (same recipe as the
    reference's utils.simulate_helical_projection:31-189, restated; the
    reference itself is not available on the GPU box)."""
    rng = np.random.default_rng(seed)
    # asymmetric unit: a planar-ish chain of atoms inside the tube cross-section
    t = np.cumsum(rng.normal(0, 1, (n_atoms, 2)), axis=0)
    t -= t.mean(axis=0)
    t *= (diameter / 2 * 0.8) / np.abs(t).max()
    au = np.stack([t[:, 0], t[:, 1], rng.normal(0, 0.3, n_atoms)], axis=1)  # x, y, z (A)
    length = n * apix
    hmax = int(np.ceil(length / 2 / rise)) + 2
    img = np.zeros((n, n), dtype=np.float64)
    yy = (np.arange(n) - n // 2) * apix
    xx = (np.arange(n) - n // 2) * apix
    sig2 = (ball_radius / 1.5) ** 2
    for h in range(-hmax, hmax + 1):
        a = np.deg2rad(twist * h)
        ca, sa = np.cos(a), np.sin(a)
        x = ca * au[:, 0] - sa * au[:, 1]
        z = au[:, 2] + h * rise
        # image rows = x (across the filament), columns = z (along the axis); projection along y
        gy = np.exp(-(yy[:, None] - x[None, :]) ** 2 / (2 * sig2))
        gx = np.exp(-(xx[:, None] - z[None, :]) ** 2 / (2 * sig2))
        img += gy @ gx.T
    return np.ascontiguousarray(img, dtype=np.float32)


def grid_tasks():
    from helicon_b200.grid import build_tasks

    tasks, ntot = build_tasks(N_IMG, N_IMG, APIX, TWISTS, RISES, csyms=(1,), reconstruct_length_rise=3,
                              target_apix3d=0, sym_oversample=-1)
    return tasks


# ---------------------------------------------------------------------------
# clocks sampling during the timed region (B200_PROFILING.md)
# ---------------------------------------------------------------------------
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, device_index):
        self.idx = device_index
        self.lines = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200", "-i", str(self.idx)],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thr = threading.Thread(target=self._read, daemon=True)
            self.thr.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=["unavailable"])
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for nm, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return dict(sm_mhz=float(np.median(sm)) if sm else None, sm_max_mhz=float(max(mx)) if mx else None,
                    reasons=sorted(reasons), samples=len(sm))


# ---------------------------------------------------------------------------
# CPU arm: the oracle port of the reference's path (bounded sample)
# ---------------------------------------------------------------------------
def _cpu_candidate(args):
    """One candidate through the oracle port on one core: full matrix build, scipy
    LSMR for `iters` iterations (the first stage of the reference's lsq_linear
    call, solver_linear_regression.py:258-270), reprojection + cosine score.
    Returns (seconds_build, seconds_per_iteration, seconds_score)."""
    os.environ["OMP_NUM_THREADS"] = "1"  # the reference forces this at import (lib/transforms.py:14)
    import warnings

    warnings.filterwarnings("ignore")
    from scipy.sparse import vstack
    from scipy.sparse.linalg import lsmr

    from oracle import denovo3d_oracle as O

    img, twist, rise_px, L3, target, iters = args
    N = img.shape[0]
    t0 = time.perf_counter()
    A, b, pid = O.build_A_data_matrix_fast(img, 1.0, twist, rise_px, 1, N, N, N, 0, L3, target)
    As, bs = O.build_A_helical_sym_matrix(L3, N, N, twist, rise_px, 1, 0.0, N // 2 - 1, target, "nn")
    AA = vstack((A, As)).tocsr()
    bb = np.concatenate((b, bs))
    t1 = time.perf_counter()
    r = lsmr(AA, bb, maxiter=iters, atol=1e-4, btol=1e-4)
    t2 = time.perf_counter()
    x = r[0].astype(np.float32)
    score = O.cosine_similarity(A.dot(x), b)
    t3 = time.perf_counter()
    return (t1 - t0, (t2 - t1) / max(1, r[2]), t3 - t2, int(r[2]))


def cpu_sample(img, tasks, n_parallel, iters, itn_full):
    """Time `n_parallel` candidates concurrently (one per core) and scale the LSMR
    part to `itn_full` iterations per candidate."""
    from concurrent.futures import ProcessPoolExecutor

    from helicon_b200.planner import MAX_EQUATIONS

    g = tasks[0].geom
    ndisk = int(np.count_nonzero(
        (np.add.outer((np.arange(g["D3"]) - g["D3"] // 2) ** 2, (np.arange(g["D3"]) - g["D3"] // 2) ** 2))
        < (g["D3"] // 2 - 1) ** 2))
    target = min(MAX_EQUATIONS, int(max(g["D2"] * g["L2"], g["L3"] * ndisk) * g["sym_oversample"]))
    jobs = [(img, t.twist, t.rise / g["apix3d"], g["L3"], target, iters) for t in tasks[:n_parallel]]
    t0 = time.perf_counter()
    if n_parallel == 1:
        outs = [_cpu_candidate(jobs[0])]
    else:
        with ProcessPoolExecutor(max_workers=n_parallel) as ex:
            outs = list(ex.map(_cpu_candidate, jobs))
    wall = time.perf_counter() - t0
    per_cand = [tb + tit * itn_full + ts for tb, tit, ts, _ in outs]
    # candidates run concurrently: throughput = n / (slowest scaled candidate)
    return len(jobs) / max(per_cand), wall, outs


# ---------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=4)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=200, help="candidates per GPU per step")
    ap.add_argument("--positive", type=int, default=0,
                    help="positive_constraint passed to the solver (0 = unbounded LSMR path; -1 = reference default)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-positive-rule", action="store_true", help="skip the secondary positive_constraint=-1 measurement")
    ap.add_argument("--no-trilinear", action="store_true", help="skip the secondary interpolation='linear' measurement")
    ap.add_argument("--no-pipeline", action="store_true", help="prepare and solve batches strictly one after the other")
    ap.add_argument("--e2e-batches", type=int, default=3, help="batches per search_grid() call of the e2e measurement")
    ap.add_argument("--cpu-iters", type=int, default=200,
                    help="scipy-LSMR iterations timed in the cpu_baseline sample (200 -> ~15-20 s of CPU work)")
    ap.add_argument("--ref-iters", type=int, default=16, help="--impl reference: LSMR iterations timed per sampled candidate")
    ap.add_argument("--ref-budget-s", type=float, default=200.0, help="--impl reference: wall-clock budget of the run")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    img = synthetic_filament()
    tasks = grid_tasks()
    itn_typical = 300

    if args.impl == "reference":
        if rank != 0:
            return 0
        cores = len(os.sched_getaffinity(0))
        small = synthetic_filament(n=64, apix=APIX * 4)
        for _ in range(args.warmup):  # warm-up: imports, page cache, worker start-up (small problem)
            _cpu_candidate((small, -1.3, 4.75 / (APIX * 4), 4, 30000, 5))
        vals, t_steps = [], []
        iters = args.ref_iters
        t_budget, t_start = args.ref_budget_s, time.perf_counter()
        for s in range(args.steps):
            sel = tasks[(s * cores * 37) % (len(tasks) - cores):][:cores]
            v, wall, outs = cpu_sample(img, sel, cores, iters, itn_typical)
            vals.append(v); t_steps.append(wall)
            # keep the whole run inside the budget: fewer LSMR iterations per sample for the remaining steps
            left = t_budget - (time.perf_counter() - t_start)
            if s + 1 < args.steps:
                t_it = max(o[1] for o in outs)
                t_fix = max(o[0] + o[2] for o in outs)
                iters = int(max(4, min(args.ref_iters, (left / (args.steps - s - 1) - t_fix) / max(t_it, 1e-3))))
        value = float(np.mean(vals))
        line = dict(
            impl="reference", metric="denovo3D candidates/sec (solve+score)", value=value, unit="candidates/s",
            n_gpus=args.gpus, steps=args.steps, warmup=args.warmup, ms_per_step=float(np.mean(t_steps) * 1e3),
            higher_is_better=True, scaling="weak", vs_baseline=None, dtype="f32/f64 mixed", data="synthetic",
            config=dict(workload=WORKLOAD, positive_constraint=args.positive),
            cpu_baseline=dict(value=value, unit="candidates/s", cores=cores, kind="port",
                              sample=f"{cores} grid candidates per step, one per core in parallel processes: full matrix "
                                     f"build + up to {args.ref_iters} scipy-LSMR iterations each (fewer when the "
                                     f"{args.ref_budget_s:.0f} s budget of the run runs out), LSMR time scaled to "
                                     f"{itn_typical} iterations (typical for this grid) + score"),
            e2e=dict(value=value, unit="candidates/s", h2d_bytes_per_step=0, d2h_bytes_per_step=0),
        )
        print(json.dumps(line))
        return 0

    import torch

    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist

        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    from helicon_b200.engine import Batch, Problem
    from helicon_b200.grid import search_grid
    from helicon_b200.planner import MAX_EQUATIONS, CandidateSpec, positive_rule

    g = tasks[0].geom
    assert all(t.geom["L3"] == g["L3"] for t in tasks), "cfg2 grid is expected to share one L3"
    stream = torch.cuda.current_stream()
    prob = Problem(img, g["s"], g["D2"], g["L2"], g["D3"], 0.0, g["D3"] // 2 - 1, device=local_rank, stream=stream)
    n3 = g["L3"] * prob.ndisk
    target = min(MAX_EQUATIONS, int(max(g["D2"] * g["L2"], n3) * g["sym_oversample"]))

    def specs_for(step):
        # batch index over the whole job: step-major, then rank
        bi = (step * world + rank) * args.batch
        sel = [tasks[(bi + i) % len(tasks)] for i in range(args.batch)]
        return sel, [CandidateSpec(t.twist, t.rise / g["apix3d"], 1, target, target,
                                   positive_rule(args.positive, t.rise / g["apix3d"], t.twist, g["L3"])) for t in sel]

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def gather_scores(scores_np):
        t = torch.from_numpy(scores_np).cuda()
        if dist is None:
            return t
        out = [torch.empty_like(t) for _ in range(world)]
        dist.all_gather(out, t)
        return torch.cat(out)

    from helicon_b200.grid import BatchPipeline

    pipe = BatchPipeline(device=local_rank, pipelined=not args.no_pipeline)

    def run_steps(steps, profile=False):
        """Steps run back to back; the host planning + GPU setup of step s+1 (worker thread, second stream) overlaps
        the solve of step s.  Everything -- planning, map/row builds, solve, score, NCCL gather -- is inside."""
        out = []
        sels = [specs_for(s) for s in steps]
        for i, batch in pipe.run(prob, g["L3"], [sp for _, sp in sels]):
            res = batch.solve(profile=int(profile))
            tm = batch.timing()
            md = [batch.rows_padded(c)[0] for c in range(0, batch.nc, max(1, batch.nc // 8))]
            batch.close()
            allsc = gather_scores(res["score"].astype(np.float32))
            top = torch.topk(allsc, min(10, allsc.numel()))
            out.append((res, tm, float(np.mean(md)), top))
        return out

    # ---- kernel-path timing: image resident, candidates -> scores -------------
    run_steps(range(args.warmup))
    sampler = ClockSampler(local_rank)
    barrier()
    sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    launches, itn_sum, ncand = 0, 0, 0
    fwd_ms, fwd_launches, fwd_bytes, iter_bytes = 0.0, 0, 0.0, 0.0
    adj_ms, upd_ms, sym_ms, scal_ms, lsmr_ms = 0.0, 0.0, 0.0, 0.0, 0.0
    for res, tm, md_mean, top in run_steps(range(args.warmup, args.warmup + args.steps), profile=True):
        launches += tm["launches"]
        itn_sum += int(res["itn"].sum()); ncand += len(res)
        fwd_ms += tm["fwd_data_ms"]; fwd_launches += tm["fwd_data_launches"]
        adj_ms += tm["adj_ms"]; upd_ms += tm["update_ms"]; sym_ms += tm["fwd_sym_ms"]; scal_ms += tm["scalar_ms"]
        lsmr_ms += tm["lsmr_ms"]
        # algorithmic bytes of the forward projector: read v (4n) + read/write u (8 m_data) per
        # candidate-iteration, summed over the iterations each candidate was active
        fwd_bytes += float(res["itn"].sum()) * (4.0 * n3 + 8.0 * md_mean)
        # whole LSMR iteration: B_iter = 56 n + 12 m bytes with m = data + symmetry rows
        iter_bytes += float(np.sum(res["itn"].astype(np.float64) * (56.0 * n3 + 12.0 * (md_mean + res["n_sym_rows"]))))
    torch.cuda.synchronize()  # the batches run on the library's own streams
    e1.record(stream)
    barrier()
    clocks = sampler.stop()
    t_ms = e0.elapsed_time(e1)
    tt = torch.tensor([t_ms], device="cuda")
    if dist is not None:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    t_ms = float(tt.item())
    total_cands = ncand * world
    value = total_cands / (t_ms / 1e3)
    pipe.close()

    # ---- end-to-end through the public API: host image in, host scores out ----
    e2e_vals = []
    e2e_bytes_in = img.nbytes
    per_rank = args.batch
    n_tw = max(1, per_rank // N_RISE) * args.e2e_batches  # twists per search_grid call and rank
    launches_e2e = 0
    from helicon_b200.distributed import gather_grid_results

    for s in range(2):
        # every rank is handed the same twist list (world x the per-rank share) and solves its round-robin shard;
        # the score tiles and local top-K are all-gathered (NCCL) inside the timed region
        bi = (args.warmup + args.steps) * world * per_rank + s * n_tw * world * N_RISE
        tw_idx = [(bi // N_RISE + q) % N_TWIST for q in range(n_tw * world)]
        barrier()
        t0 = time.perf_counter()
        out = search_grid(np.array(img, copy=True), APIX, TWISTS[tw_idx], RISES, positive_constraint=args.positive,
                          device=local_rank, stream=stream, batch_candidates=per_rank, pipelined=not args.no_pipeline,
                          shard=(rank, world))
        launches_local = out["launches"]
        out = gather_grid_results(out, top_k=10, dist=dist, device="cuda")
        best = float(np.nanmax(out["scores"]))  # host read of the step's result
        barrier()
        dt = time.perf_counter() - t0
        tt = torch.tensor([dt], device="cuda")
        if dist is not None:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        if s > 0:
            e2e_vals.append(out["n_candidates"] / float(tt.item()))
            launches_e2e = launches_local
    e2e_val = float(np.mean(e2e_vals))
    d2h = int(out["scores"].size * 4 + out["itn"].size * 4)

    # ---- the same call with the reference's DEFAULT positive-constraint rule (SLR:352-355): for this amyloid-like
    # geometry every candidate then also runs the bounded TRF branch of scipy's lsq_linear (float64) --------------
    posrule = None
    if args.positive == 0 and not args.no_positive_rule:
        bi = (args.warmup + args.steps) * world * per_rank + 2 * n_tw * world * N_RISE
        tw_idx = [(bi // N_RISE + q) % N_TWIST for q in range(2 * world)]
        barrier()
        t0 = time.perf_counter()
        outp = search_grid(np.array(img, copy=True), APIX, TWISTS[tw_idx], RISES, positive_constraint=-1,
                           device=local_rank, stream=stream, batch_candidates=100, pipelined=not args.no_pipeline,
                           shard=(rank, world))
        outp = gather_grid_results(outp, top_k=10, dist=dist, device="cuda")
        barrier()
        tt = torch.tensor([time.perf_counter() - t0], device="cuda")
        if dist is not None:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        posrule = dict(value=outp["n_candidates"] / float(tt.item()), unit="candidates/s",
                       candidates=int(outp["n_candidates"]),
                       bounded_fraction=float(np.mean((outp["flags"][np.isfinite(outp["scores"])] & 4) != 0)),
                       note="search_grid(positive_constraint=-1), host image in -> host scores out; one un-warmed call")

    # ---- secondary: trilinear interpolation (the app's default mode) through the same call: explicit GPU-built rows,
    # one candidate at a time; a few candidates on rank 0 only, un-warmed -------------------
    trilinear = None
    if rank == 0 and not args.no_trilinear:
        t0 = time.perf_counter()
        outl = search_grid(np.array(img, copy=True), APIX, TWISTS[[398, 402]], RISES[[24, 26]], positive_constraint=0,
                           device=local_rank, stream=stream, interpolation="linear")
        dtl = time.perf_counter() - t0
        trilinear = dict(value=outl["n_candidates"] / dtl, unit="candidates/s", candidates=int(outl["n_candidates"]),
                         n_gpus=1, mean_lsmr_iterations=float(outl["itn"].mean()), best_score=float(np.nanmax(outl["scores"])),
                         note="search_grid(interpolation='linear', positive_constraint=0) on one GPU: row build on the GPU "
                              "+ LSMR on the explicit CSR (192 M entries per candidate at this shape), one un-warmed call")

    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return 0

    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback 6650 GB/s (B200_PROFILING.md)"
    achieved = fwd_bytes / (fwd_ms / 1e3) / 1e9 if fwd_ms > 0 else 0.0
    # DRAM traffic of the dominant kernel from the committed ncu capture (profiles/r1_traffic.json), scaled from the
    # capture's candidates per launch to this run's mean ACTIVE candidates per launch
    traffic = None
    traffic_note = "no capture"
    try:
        tj = json.load(open(os.path.join(ROOT, "profiles", "r1_traffic.json")))
        per_cand = tj["dram_bytes_per_launch"]["k_fwd_data"] / tj["candidates_per_launch"]
        traffic = per_cand * itn_sum / max(1, fwd_launches)
        traffic_note = ("dram__bytes_read.sum + dram__bytes_write.sum of k_fwd_data from " + tj["source"] +
                        f": {per_cand/1e6:.2f} MB per candidate-pass x {itn_sum / max(1, fwd_launches):.1f} active "
                        "candidates per launch in this run")
    except Exception:
        pass
    line = dict(
        metric="denovo3D candidates/sec (solve+score)", value=value, unit="candidates/s", n_gpus=world,
        steps=args.steps, warmup=args.warmup, ms_per_step=t_ms / args.steps, higher_is_better=True, scaling="weak",
        vs_baseline=None, dtype="f32 (u,v,h) / f64 (x,hbar), as scipy executes LSMR", data="synthetic",
        config=dict(workload=WORKLOAD, step=f"{args.batch} consecutive grid candidates per GPU", L3=g["L3"],
                    pipelined=not args.no_pipeline,
                    unknowns_per_candidate=n3, positive_constraint=args.positive,
                    cache="inputs of every step are new candidates; per-step working set >> L2 (126 MB)",
                    mean_lsmr_iterations=itn_sum / max(1, ncand)),
        clocks=clocks, gpu_launches=int(launches),
        e2e=dict(value=e2e_val, unit="candidates/s", h2d_bytes_per_step=int(e2e_bytes_in), d2h_bytes_per_step=d2h,
                 candidates_per_call=int(out["n_candidates"]), best_score=best,
                 note="search_grid(): host image -> Problem upload, host planning, solve, scores copied back; bytes are per "
                      "search_grid() call"),
        roofline=dict(bound="hbm", kernel="k_fwd_data (forward projector u <- A v - alpha u)", achieved=achieved,
                      peak=peak, unit="GB/s", frac=achieved / peak if peak else None, traffic=traffic,
                      traffic_source=traffic_note,
                      algorithmic_bytes_per_launch=fwd_bytes / max(1, fwd_launches),
                      on_chip="the kernel is bound on chip, not by HBM: ncu (profiles/r1_summary.md) 82 % of the LSU data "
                              "pipe, 9.9 TB/s L2->L1 (179 MB of gathers per candidate-pass served by L2), DRAM traffic = "
                              "algorithmic bytes",
                      peak_source=peak_src,
                      avg_launch_ms=fwd_ms / max(1, fwd_launches),
                      note="achieved = algorithmic bytes (4n + 8 m_data per active candidate-iteration) / summed "
                           "CUDA-event time of the kernel's launches inside the timed region"),
        roofline_iteration=dict(
            bound="hbm", achieved=iter_bytes / (lsmr_ms / 1e3) / 1e9 if lsmr_ms > 0 else 0.0, peak=peak, unit="GB/s",
            frac=(iter_bytes / (lsmr_ms / 1e3) / 1e9 / peak) if lsmr_ms > 0 and peak else None,
            note="all kernels of the LSMR phase together: algorithmic bytes B_iter = 56 n + 12 m per active "
                 "candidate-iteration (SURVEY 8d) / device time of the LSMR phase"),
        kernel_share=dict(lsmr_phase_ms=lsmr_ms, fwd_data_ms=fwd_ms, fwd_sym_ms=sym_ms, adjoint_ms=adj_ms,
                          update_ms=upd_ms, scalar_ms=scal_ms),
    )
    if posrule is not None:
        line["e2e_positive_rule_default"] = posrule
    if trilinear is not None:
        line["e2e_trilinear"] = trilinear
    if not args.no_cpu_baseline:
        sel = [tasks[(args.warmup * args.batch + 17) % len(tasks)]]
        itn_ref = int(round(itn_sum / max(1, ncand)))
        v, wall, outs = cpu_sample(img, sel, 1, args.cpu_iters, itn_ref)
        line["cpu_baseline"] = dict(
            value=v, unit="candidates/s", cores=1, kind="port",
            sample=f"1 grid candidate (twist={sel[0].twist}, rise={sel[0].rise:.4f}): full matrix build "
                   f"({outs[0][0]:.1f} s) + {outs[0][3]} scipy-LSMR iterations ({outs[0][1]*1e3:.0f} ms each), LSMR time "
                   f"scaled to the {itn_ref} iterations the GPU path averaged + score; wall {wall:.1f} s")
    print(json.dumps(line))
    if dist is not None:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
