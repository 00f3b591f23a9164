#!/usr/bin/env python
"""bench.py — SSRS hot path on B200: track-steps/s (+ updraft/potential field time per grid).

    python bench.py --gpus 1 --steps 5 --warmup 3                       # N=1: BASELINE configs[1]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 ... bench.py --gpus N ...
    python bench.py --impl reference ...      # the reference's own Python path on the host cores
    python bench.py --workload config4 ...    # config2 | config3 | config4 | config5 (BASELINE.json configs[1..4])

A "step" is one pass of stages 3+4 (batched track stepping + presence accumulation, plus the presence all-reduce when
N > 1) over one batch of tracks on fields that are already resident in HBM; every step produces its own presence map
in its own buffer.  Consecutive steps are issued on a small ring of CUDA streams, so the long tail of one batch (a launch
lasts as long as its longest track) overlaps the bulk of the next ones; the field stages (1: updraft stencil, 2: potential
solve) are timed once per run and reported under "fields".
Workloads (BASELINE.json configs):
  config2 (default at N=1)  uniform mode, 6000x5000 DEM at 10 m, 100k tracks per step and GPU (weak scaling if N>1)
  config3 (default at N>1)  as config2 with 1M tracks per step in total, block-partitioned over the N GPUs (strong scaling)
  config4                   snapshot mode: wind at a jittered 2 km lattice of sites interpolated to the grid on the GPU,
                            per-cell-wind stencil, 1M tracks per step in total
  config5                   seasonal mode: a step is one wind CASE on a 12000x10000 DEM — stencil, potential solve
                            (row-sharded over the ranks, or whole cases per rank with --seasonal-mode case_parallel),
                            100k tracks, presence map — so `value` is whole-pipeline track-steps/s
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
# stdout carries exactly one JSON line.  NCCL (and anything else below Python) may write banners straight to file
# descriptor 1 — NCCL_DEBUG_FILE does not cover its version line — so descriptor 1 is pointed at stderr for the whole run
# and the JSON line goes through a private duplicate of the original stdout.
os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
sys.stdout.flush()
_JSON_FD = os.dup(1)
os.dup2(2, 1)


def emit(line: dict) -> None:
    os.write(_JSON_FD, (json.dumps(line) + "\n").encode())


METRIC = "track-steps/sec"
UNIT = "track-steps/s"
STEP_BYTES = 76          # 72 B of 3x3 gathers on two f32 fields + 4 B presence atomic (no trajectory store)
STENCIL_BYTES = 20       # 4 B DEM + 4 outputs x 4 B
# Per-launch figures of the stepping kernel from the `ncu --set full` capture of this command's default workload
# (profiles/r02_ncu_step_tracks_*.txt): DRAM bytes (dram__bytes_read.sum + dram__bytes_write.sum) and L2 bytes
# (lts__t_bytes.sum), both per 1e9 track-steps (one 100k-track batch), and the measured L2 bandwidth of the box
# (tools/l2_peak.py, profiles/r02_l2_peak.txt).  None until measured.
STEP_DRAM_BYTES_PER_STEP = None
STEP_L2_BYTES_PER_STEP = None
L2_PEAK_GBS = None
try:
    with open(os.path.join(ROOT, "profiles", "r02_stepping_traffic.json")) as _f:
        _t = json.load(_f)
        STEP_DRAM_BYTES_PER_STEP, STEP_L2_BYTES_PER_STEP, L2_PEAK_GBS = _t["dram_bytes_per_step"], _t["l2_bytes_per_step"], _t["l2_peak_gbs"]
except (OSError, KeyError, ValueError):
    pass


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ssrs_b200", choices=["ssrs_b200", "reference"])
    ap.add_argument("--workload", default="auto", choices=["auto", "config2", "config3", "config4", "config5"])
    ap.add_argument("--rows", type=int, default=0)
    ap.add_argument("--cols", type=int, default=0)
    ap.add_argument("--resolution", type=float, default=10.0)
    ap.add_argument("--tracks-per-gpu", type=int, default=0, help="tracks per step and GPU (weak scaling)")
    ap.add_argument("--tracks-total", type=int, default=0, help="tracks per step over all GPUs (strong scaling)")
    ap.add_argument("--seed", type=int, default=2021)
    ap.add_argument("--streams", type=int, default=12, help="steps in flight (ring of CUDA streams and presence buffers)")
    ap.add_argument("--mode", default="phased", choices=["phased", "single", "walk"],
                    help="stepping entry point: ssrs_step_tracks_phased (default), ssrs_step_tracks, ssrs_walk_tracks")
    ap.add_argument("--seasonal-mode", default="sharded", choices=["sharded", "case_parallel"])
    ap.add_argument("--cpu-sample-tracks", type=int, default=0, help="tracks in the CPU baseline sample (0 = auto)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--emit-fields", default="", help="(internal) write {updraft, potential} of the workload to this .npz and exit")
    return ap.parse_args()


def resolve_workload(a, world):
    name = a.workload
    if name == "auto":
        name = "config2" if world == 1 else "config3"
    big = name == "config5"
    rows = a.rows or (10000 if big else 5000)
    cols = a.cols or (12000 if big else 6000)
    w = {"name": name, "rows": rows, "cols": cols, "res": a.resolution,
         "wind": {"config2": "uniform", "config3": "uniform", "config4": "snapshot", "config5": "seasonal"}[name]}
    if name == "config2":
        per = a.tracks_per_gpu or (a.tracks_total // world if a.tracks_total else 100_000)
        w.update(n_total=per * world, scaling="weak")
    elif name in ("config3", "config4"):
        tot = a.tracks_total or (a.tracks_per_gpu * world if a.tracks_per_gpu else 1_000_000)
        w.update(n_total=tot, scaling="strong")
    else:
        tot = a.tracks_total or (a.tracks_per_gpu * world if a.tracks_per_gpu else 100_000)
        w.update(n_total=tot, scaling="strong")
    return w


def workload_config(w, world):
    """`config` of the JSON line: the workload and nothing run-specific, so that it is identical in both arms (the driver
    compares the two arms' `config`); what a run did beyond that is under the line's `run` key."""
    mb = w["rows"] * w["cols"] * 4 / 1e6
    l2 = (f"inputs_exceed_l2 (fields {2 * mb:.0f} MB + presence {mb:.0f} MB > 126 MB; every timed step is another realisation "
          f"and writes another presence raster)") if 3 * mb > 126 else \
         f"inputs fit L2 ({3 * mb:.1f} MB): a custom small grid, not a BASELINE workload"
    return {"workload": workload_text(w, world), "grid": [w["rows"], w["cols"]], "tracks_per_step": w["n_total"], "l2": l2}


def workload_text(w, world):
    wind = {"uniform": "uniform mode, wind 10 m/s from 270 deg", "snapshot": "snapshot mode, wind at a jittered 2 km lattice of "
            "sites interpolated to the grid", "seasonal": "seasonal mode, one sampled wind condition per step"}[w["wind"]]
    return (f"{w['name']}: {wind}, synthetic {w['cols']}x{w['rows']} DEM at {w['res']:g} m, {w['n_total']} northbound tracks "
            f"per step over {world} GPU(s)")


def start_cells(a_or_w, n_total, seed=None):
    """track_start_region=(5,55,1,2) km scaled to the grid, drawn like get_starting_indices('random')."""
    from ssrs_b200.movmodel import get_starting_indices
    if isinstance(a_or_w, dict):
        rows, cols, res = a_or_w["rows"], a_or_w["cols"], a_or_w["res"]
    else:
        rows, cols, res = a_or_w.rows, a_or_w.cols, a_or_w.resolution
        seed = a_or_w.seed if seed is None else seed
    width_km = (cols * res / 1000.0, rows * res / 1000.0)
    region = (width_km[0] * 5 / 60, width_km[0] * 55 / 60, width_km[1] * 1 / 50, width_km[1] * 2 / 50)
    st = np.random.get_state()
    np.random.seed(seed)
    r, c = get_starting_indices(n_total, region, "random", width_km, res)
    np.random.set_state(st)
    return r, c


class ClockSampler(threading.Thread):
    """Samples nvidia-smi clocks / throttle reasons while the timed region runs."""

    FIELDS = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
              "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.rows, self.stop_flag = index, [], False

    def run(self):
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits", "-i",
                                      str(self.index)], capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([x.strip() for x in out.split(",")])
            except (OSError, subprocess.SubprocessError):
                pass
            time.sleep(0.1)

    def summary(self):
        sm = [float(r[0]) for r in self.rows if r and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in self.rows for i in range(4) if len(r) > 2 + i and r[2 + i].startswith("Active")})
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


# ---------------------------------------------------------------------------------------------------------------------
# CPU legs (test infrastructure: oracle/)
# ---------------------------------------------------------------------------------------------------------------------
def cpu_port_rate(U, P, shape, sr, sc, n_sample, threads, seed):
    """C restatement of the reference's stepper (oracle/ssrs_oracle.c) on `threads` host threads."""
    from oracle import oracle_c as OC
    st = np.stack([sr[:n_sample], sc[:n_sample]], 1).astype(np.int32)
    t0 = time.perf_counter()
    out = OC.step_tracks(U, P, shape, st, 0.0, 1, 1.0, seed=seed, want_presence=True, nthreads=threads, fast=True)
    dt = time.perf_counter() - t0
    return out["total_steps"] / dt, out["total_steps"], dt


def cpu_reference_rate(U, P, sr, sc, n_sample, procs, seed):
    """The reference's own generate_simulated_tracks through its pool pattern (oracle/ref_cpu.py)."""
    from oracle import ref_cpu
    steps, dt, procs = ref_cpu.pool_track_steps(U, P, sr[:n_sample], sc[:n_sample], 0.0, procs=procs, seed=seed)
    return steps / dt, steps, dt, procs


def cpu_baseline_block(U, P, shape, sr, sc, a):
    """`cpu_baseline` of the JSON line: the reference's Python stepper on all host cores on a bounded sample of the
    step's tracks and fields (kind "reference"); the C port's rate rides along as `port`.  Without the reference's
    modules (neither /root/reference nor the staged oracle/_ref) the port is the baseline and says so."""
    from oracle import ref_loader
    threads = os.cpu_count() or 1
    n_port = min(len(sr), 4096 * threads)
    rate_p, steps_p, dt_p = cpu_port_rate(U, P, shape, sr, sc, n_port, threads, a.seed)
    port = {"value": rate_p, "unit": UNIT, "cores": threads, "kind": "port",
            "sample": f"{n_port} of the step's tracks on the same fields, {steps_p} track-steps in {dt_p:.1f} s "
                      f"(oracle/ssrs_oracle.c, OpenMP)"}
    if not ref_loader.available():
        port["note"] = "reference modules unavailable (no /root/reference, no staged oracle/_ref): C port only"
        return port
    # ~1e4 track-steps/s/core: size the sample for about 15 s
    mean_len = max(1.0, steps_p / n_port)
    n_ref = a.cpu_sample_tracks or int(max(threads, min(len(sr), 15.0 * 1.0e4 * threads / mean_len)))
    rate, steps, dt, procs = cpu_reference_rate(U, P, sr, sc, n_ref, threads, a.seed)
    return {"value": rate, "unit": UNIT, "cores": procs, "kind": "reference",
            "sample": f"{n_ref} of the step's tracks on the same fields, {steps} track-steps in {dt:.1f} s: the unmodified "
                      f"ssrs/movmodel.py generate_simulated_tracks through multiprocess.Pool({procs}).map "
                      f"(the pattern of ssrs/simulator.py:360-369)",
            "port": port}


# ---------------------------------------------------------------------------------------------------------------------
# fields on the device
# ---------------------------------------------------------------------------------------------------------------------
def wind_rasters(w, torch, case=None):
    """(wspeed, wdirn) for the stencil: scalars in uniform mode; CUDA rasters interpolated from the synthetic site
    lattice otherwise (case = (speed0, dirn0) of a seasonal condition)."""
    if w["wind"] == "uniform":
        return 10.0, 270.0, {}
    from ssrs_b200 import layers
    from ssrs_b200.synth import synthetic_wind_lattice
    s0, d0 = case if case is not None else (8.0, 270.0)
    xl, yl, spd, drn = synthetic_wind_lattice(w["rows"], w["cols"], w["res"], spacing_m=2000.0, seed=7, speed0=s0, dirn0=d0)
    tri = w.setdefault("_tri", layers.delaunay_triangles(xl, yl))
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    ev[0].record()
    ws, wd = layers.interpolate_wind_to_grid(xl, yl, spd, drn, 0.0, 0.0, w["res"], (w["rows"], w["cols"]), triangles=tri)
    ev[1].record()
    torch.cuda.synchronize()
    return ws, wd, {"wind_sites": int(len(xl)), "wind_interp_ms": ev[0].elapsed_time(ev[1])}


def build_fields_gpu(a, torch, world=1, w=None, check_sharded=False):
    """Stage 1 (+2) on the device; returns (updraft, potential, info).  With world > 1 the potential solve is
    row-sharded over the ranks (ssrs_potential_solve_sharded, NCCL halo exchanges)."""
    from ssrs_b200 import layers
    from ssrs_b200.potential import solve_potential_device
    from ssrs_b200.synth import synthetic_dem
    if w is None:
        w = {"rows": a.rows, "cols": a.cols, "res": a.resolution, "wind": "uniform"}
    rows, cols, res = w["rows"], w["cols"], w["res"]
    z = torch.from_numpy(synthetic_dem(rows, cols, res)).cuda()
    ws, wd, info = wind_rasters(w, torch)
    per_cell = w["wind"] != "uniform"
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    for _ in range(3):
        out = layers.updraft_fields(z, res, ws, wd, 0.75)
    torch.cuda.synchronize()
    reps = 10
    ev[0].record()
    for _ in range(reps):
        out = layers.updraft_fields(z, res, ws, wd, 0.75)
    ev[1].record()
    torch.cuda.synchronize()
    ms = ev[0].elapsed_time(ev[1]) / reps
    peak, _ = measured_peak()
    nbytes = STENCIL_BYTES + (8 if per_cell else 0)
    gbs = nbytes * rows * cols / (ms * 1e-3) / 1e9
    info["updraft_ms"] = ms
    info["updraft_roofline"] = {"bound": "hbm", "achieved": gbs, "peak": peak, "unit": "GB/s", "frac": gbs / peak,
                                "bytes_per_cell": nbytes}
    up = out["updraft"]
    # warm-up on a small grid: CUDA loads each kernel lazily on its first launch
    zs = torch.from_numpy(synthetic_dem(256, 320, res, seed=1)).cuda()
    ks = layers.updraft_fields(zs, res, 10.0, 270.0, 0.75, want=("updraft",))["updraft"]
    solve_potential_device(ks, 0.0)
    torch.cuda.synchronize()
    sharded = world > 1
    if sharded:
        import torch.distributed as dist
        from ssrs_b200 import dist as D
        D.native_comm()                       # communicator creation is a one-time cost of the process, not of a solve
    # first solve at this size also grows the workspace arena (reported separately); the second is the per-grid time a
    # multi-case run (seasonal mode: one solve per wind case) sees
    for key in ("potential_first_ms", "potential_ms"):
        if sharded:
            dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        pot, stats = solve_potential_device(up, 0.0, sharded=sharded)
        torch.cuda.synchronize()
        info[key] = (time.perf_counter() - t0) * 1e3
    info["potential_stats"] = stats
    info["potential_sharded_over"] = world
    if sharded:
        info["potential_halo_mode"] = D.halo_mode()       # 'peer': halos stored into the neighbours' memory; 'nccl': send/recv
    if sharded and check_sharded:
        # correctness of the row-sharded solve, outside any timed region: against this GPU's own single-GPU solve
        ref, _ = solve_potential_device(up, 0.0, sharded=False)
        ulp = float(np.spacing(np.float32(1000.0)))
        diff = float((pot.double() - ref.double()).abs().max().item()) / ulp
        info["sharded_vs_single_max_ulp"] = diff
        if not diff <= 2.0:
            raise RuntimeError(f"row-sharded potential differs from the single-GPU solve by {diff:.2f} float32 ulp (> 2)")
    return up, pot, info


# ---------------------------------------------------------------------------------------------------------------------
# stepping: a ring of streams, one presence buffer and workspace per slot
# ---------------------------------------------------------------------------------------------------------------------
class StepRing:
    def __init__(self, torch, shape, n_rank, slots, mode, world, fields=None):
        from ssrs_b200 import _native as N
        from ssrs_b200 import movmodel as mm
        self.torch, self.mm, self.shape, self.mode, self.world = torch, mm, shape, mode, world
        self.slots = max(1, slots)
        self.streams = [torch.cuda.Stream() for _ in range(self.slots)]
        self.presence = [torch.zeros(shape, dtype=torch.int32, device="cuda") for _ in range(self.slots)]
        wsb = int(N.load().ssrs_walk_workspace_bytes(n_rank))
        self.workspace = [torch.empty(wsb, dtype=torch.uint8, device="cuda") for _ in range(self.slots)] if mode != "single" else None
        self.reduce_stream = torch.cuda.Stream() if world > 1 else None
        self.free = [None] * self.slots          # event after which slot i's presence buffer may be overwritten
        self.tables = None
        if mode == "walk":
            tb = int(N.load().ssrs_walk_table_bytes(*shape))
            self.tables = [torch.empty(tb, dtype=torch.uint8, device="cuda") for _ in range(min(self.slots, 3))]

    def issue(self, i, fields, sr, sc, seed, track_id0, total, kev=None, after=None):
        """Step i on slot i % slots; returns the slot.  `after`: event the step's stream must wait for first.
        sr: host start rows with sc the columns, or a device int32 [n, 2] tensor with sc None."""
        torch, mm = self.torch, self.mm
        k = i % self.slots
        s = self.streams[k]
        if after is not None:
            s.wait_event(after)
        if self.free[k] is not None:
            s.wait_event(self.free[k])
        with torch.cuda.stream(s):
            self.presence[k].zero_()
            if kev is not None:
                kev[0].record(s)
            extra = {}
            if self.mode == "phased":
                extra = dict(phased=True, workspace=self.workspace[k])
            elif self.mode == "walk":
                tab = mm.build_transition_table(fields, 0.0, out=self.tables[i % len(self.tables)])
                extra = dict(walk=True, table=tab, workspace=self.workspace[k])
            mm.simulate_tracks_batch(0.0, sr, sc, self.shape, fields=fields, seed=seed, track_id0=track_id0,
                                     presence=self.presence[k], total_steps=total, **extra)
            if kev is not None:
                kev[1].record(s)
            done = torch.cuda.Event()
            done.record(s)
        if self.world > 1:
            from ssrs_b200 import dist as D
            # all ranks issue the all-reduces in step order on one side stream: launch i+1 steps while map i reduces
            self.reduce_stream.wait_event(done)
            with torch.cuda.stream(self.reduce_stream):
                D.presence_allreduce(self.presence[k])          # ssrs_presence_allreduce: the library's NCCL communicator
                done = torch.cuda.Event()
                done.record(self.reduce_stream)
        self.free[k] = done
        return k, done

    def join(self):
        cur = self.torch.cuda.current_stream()
        for s in self.streams:
            cur.wait_stream(s)
        if self.reduce_stream is not None:
            cur.wait_stream(self.reduce_stream)


def phase_launches(rows, cols, mode):
    """Kernel launches of one stepping call."""
    from ssrs_b200 import _native as N
    if mode == "single":
        return 1
    return int(N.load().ssrs_step_phase_count(rows, cols, 0)) + (1 if mode == "walk" else 0)


def main():
    a = parse()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    w = resolve_workload(a, world)

    if a.impl == "reference":
        if rank != 0:
            return 0
        return reference_arm(a, w, world)

    import torch
    import torch.distributed as dist
    from ssrs_b200 import dist as D
    from ssrs_b200 import movmodel as mm
    from ssrs_b200 import _native as N

    N.load()                     # fail loudly if the CUDA library is missing
    N.require_cuda()
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    if a.emit_fields:
        up, pot, _ = build_fields_gpu(a, torch, 1, w)
        np.savez(a.emit_fields, updraft=up.cpu().numpy(), potential=pot.cpu().numpy())
        return 0
    if w["name"] == "config5":
        return seasonal_arm(a, w, rank, world, torch, dist, D, mm)

    rows, cols = w["rows"], w["cols"]
    shape = (rows, cols)
    n_total = w["n_total"]
    lo, hi = D.shard_range(n_total, rank, world)
    n_rank = hi - lo
    sr_all, sc_all = start_cells(w, n_total, a.seed)
    sr, sc = sr_all[lo:hi], sc_all[lo:hi]
    up, pot, finfo = build_fields_gpu(a, torch, world, w, check_sharded=True)
    fields = mm.interleave_fields(up, pot)
    total = torch.zeros(1, dtype=torch.int64, device="cuda")
    ring = StepRing(torch, shape, n_rank, a.streams, a.mode, world)
    # every slot of the ring is warmed once: a stream's first launches allocate (and so synchronise) in its memory pool
    warm = max(a.warmup, 3, ring.slots)
    start_h = torch.from_numpy(np.stack([sr, sc], 1).astype(np.int32)).pin_memory()
    start_d = start_h.to("cuda")                 # the device-timed steps read the start cells from HBM like the fields

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for i in range(warm):
        ring.issue(i, fields, start_d, None, a.seed, lo, total)
    ring.join()
    barrier()
    # one launch alone, between synchronisations: the latency of a single batch
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ring.issue(0, fields, start_d, None, a.seed, lo, total, kev=(e0, e1))
    ring.join()
    barrier()
    alone_ms = e0.elapsed_time(e1)
    total.zero_()
    sampler = ClockSampler(local)
    sampler.start()
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    kev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(a.steps)]
    barrier()
    t0.record()
    for i in range(a.steps):
        # every timed step is another realisation of the same workload (seed + 1 + step): a launch lasts as long as its
        # longest track, and one realisation's maximum (90k..104k steps) would decide the whole figure
        ring.issue(i, fields, start_d, None, a.seed + 1 + i, lo, total, kev=kev[i], after=t0 if i < ring.slots else None)
    ring.join()
    t1.record()
    barrier()
    sampler.stop_flag = True
    ms_total = t0.elapsed_time(t1)
    launch_ms = float(np.mean([s.elapsed_time(e) for s, e in kev]))
    steps_rank = int(total.item())
    t = torch.tensor([ms_total, float(steps_rank), launch_ms, alone_ms], dtype=torch.float64, device="cuda")
    if world > 1:
        tmax = t.clone()
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        tsum = t.clone()
        dist.all_reduce(tsum, op=dist.ReduceOp.SUM)
        ms_total, steps_all, launch_ms, alone_ms = float(tmax[0]), float(tsum[1]), float(tmax[2]), float(tmax[3])
    else:
        steps_all = float(steps_rank)
    value = steps_all / (ms_total * 1e-3)

    # ---- end-to-end through the public API with HOST buffers (H2D + D2H inside the timed region), same ring ----------
    up_h = torch.empty(shape, dtype=torch.float32).pin_memory()
    pot_h = torch.empty(shape, dtype=torch.float32).pin_memory()
    up_h.copy_(up)
    pot_h.copy_(pot)
    slots = ring.slots
    pres_h = [torch.empty(shape, dtype=torch.int32).pin_memory() for _ in range(slots)]
    dev_in = [(torch.empty(shape, dtype=torch.float32, device="cuda"), torch.empty(shape, dtype=torch.float32, device="cuda"),
               torch.empty((n_rank, 2), dtype=torch.int32, device="cuda")) for _ in range(slots)]      # landing buffers per slot
    tot_e2e = torch.zeros(1, dtype=torch.int64, device="cuda")
    copied = [None] * slots

    def e2e_step(i, seed, after=None):
        k = i % slots
        s = ring.streams[k]
        if after is not None:
            s.wait_event(after)
        if copied[k] is not None:
            s.wait_event(copied[k])              # the slot's previous map has left for the host
        with torch.cuda.stream(s):
            u_d, p_d, st_d = dev_in[k]
            u_d.copy_(up_h, non_blocking=True)
            p_d.copy_(pot_h, non_blocking=True)
            st_d.copy_(start_h, non_blocking=True)
            f_d = mm.interleave_fields(u_d, p_d)
        ring.issue(i, f_d, st_d, None, seed, lo, tot_e2e)          # steps, (all-reduces)
        out_stream = ring.reduce_stream if world > 1 else s
        with torch.cuda.stream(out_stream):
            pres_h[k].copy_(ring.presence[k], non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(out_stream)
        copied[k] = ev
        ring.free[k] = ev

    for i in range(slots):
        e2e_step(i, a.seed)
    ring.join()
    barrier()
    tot_e2e.zero_()
    e2, e3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e2.record()
    for i in range(a.steps):
        e2e_step(i, a.seed + 1 + i, after=e2 if i < slots else None)
    ring.join()
    e3.record()
    barrier()
    ms_e2e = e2.elapsed_time(e3)
    t2 = torch.tensor([ms_e2e, float(tot_e2e.item())], dtype=torch.float64, device="cuda")
    if world > 1:
        m = t2.clone(); dist.all_reduce(m, op=dist.ReduceOp.MAX)
        s_ = t2.clone(); dist.all_reduce(s_, op=dist.ReduceOp.SUM)
        ms_e2e, steps_e2e = float(m[0]), float(s_[1])
    else:
        steps_e2e = float(t2[1])
    e2e_value = steps_e2e / (ms_e2e * 1e-3)
    h2d = 2 * rows * cols * 4 + n_rank * 8
    d2h = rows * cols * 4

    if rank == 0:
        peak, peak_src = measured_peak()
        per_launch_ms = ms_total / a.steps             # launches overlap: the region's time per launch
        achieved = STEP_BYTES * (steps_rank / a.steps) / (per_launch_ms * 1e-3) / 1e9
        nl = phase_launches(rows, cols, a.mode)
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": a.steps, "warmup": a.warmup,
            "ms_per_step": ms_total / a.steps, "higher_is_better": True, "scaling": w["scaling"], "vs_baseline": None,
            "dtype": "f64 probabilities on f32 fields", "data": "synthetic",
            "config": workload_config(w, world),
            "run": {"tracks_per_gpu": n_rank, "track_steps_per_step": steps_all / a.steps,
                    "rng": "philox4x32-10 keyed by (seed, global track id, step); timed step i uses seed + 1 + i (another realisation per step)",
                    "stepping": {"phased": "ssrs_step_tracks_phased (survivors compacted between phases)",
                                 "single": "ssrs_step_tracks (one launch per batch)",
                                 "walk": "ssrs_transition_table + ssrs_walk_tracks (table rebuilt every step)"}[a.mode],
                    "steps_in_flight": ring.slots,
                    "warmup_executed": f"{warm + 1} untimed steps: max(--warmup, one per ring slot), then one batch alone "
                                       f"for launch_ms_alone",
                    "parallelism": f"tracks block-partitioned over {world} GPU(s), fields replicated, presence all-reduce "
                                   f"per step on a side stream"},
            "fields": finfo,
            "roofline": {"bound": "hbm", "kernel": "step_tracks_kernel", "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak,
                         "traffic": (STEP_DRAM_BYTES_PER_STEP * steps_rank / a.steps) if STEP_DRAM_BYTES_PER_STEP else None,
                         "traffic_unit": "bytes per step of this bench = one phased launch (ncu dram__bytes_read.sum + dram__bytes_write.sum)",
                         "peak_source": peak_src, "bytes_per_track_step": STEP_BYTES,
                         "kernel_ms": per_launch_ms, "launch_ms_alone": alone_ms, "launch_ms_in_flight": launch_ms,
                         "note": "steps overlap on a ring of streams, so kernel_ms is the timed region divided by its "
                                 "launches; launch_ms_alone is one batch between synchronisations (it lasts as long as its "
                                 "longest track), launch_ms_in_flight the mean event-to-event time of a batch in the ring. "
                                 "Not HBM-bound by construction (SURVEY §8d): the gathers are served by L1/L2; "
                                 "HBM fraction reported as the contract asks, L2 ratio under roofline_l2"},
            "roofline_l2": ({"bound": "l2", "achieved": STEP_L2_BYTES_PER_STEP * (steps_rank / a.steps) / (per_launch_ms * 1e-3) / 1e9,
                             "peak": L2_PEAK_GBS, "unit": "GB/s",
                             "frac": STEP_L2_BYTES_PER_STEP * (steps_rank / a.steps) / (per_launch_ms * 1e-3) / 1e9 / L2_PEAK_GBS,
                             "source": "lts__t_bytes.sum per track-step from the ncu capture, L2 peak measured by tools/ubench/l2bw.cu "
                                       "(profiles/r02_stepping_traffic.json)"} if STEP_L2_BYTES_PER_STEP and L2_PEAK_GBS else None),
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": ms_e2e / a.steps},
            "gpu_launches": a.steps * (nl + (1 if world > 1 else 0)),     # stepping kernels (+ the NCCL all-reduce) per step
            "clocks": sampler.summary(),
        }
        if not a.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline_block(up.cpu().numpy(), pot.cpu().numpy(), shape, sr, sc, a)
        emit(line)
    if world > 1:
        dist.barrier()
        D.destroy_native_comm()
        dist.destroy_process_group()
    return 0


# ---------------------------------------------------------------------------------------------------------------------
# config5: seasonal sweep — a step is one wind case through the whole pipeline
# ---------------------------------------------------------------------------------------------------------------------
def seasonal_arm(a, w, rank, world, torch, dist, D, mm):
    from ssrs_b200 import layers
    from ssrs_b200.potential import solve_potential_device
    from ssrs_b200.synth import seasonal_wind_conditions, synthetic_dem
    rows, cols, res = w["rows"], w["cols"], w["res"]
    shape = (rows, cols)
    case_parallel = a.seasonal_mode == "case_parallel" and world > 1
    n_total = w["n_total"]
    lo, hi = (0, n_total) if case_parallel else D.shard_range(n_total, rank, world)
    n_rank = hi - lo
    sr_all, sc_all = start_cells(w, n_total, a.seed)
    sr, sc = sr_all[lo:hi], sc_all[lo:hi]
    z = torch.from_numpy(synthetic_dem(rows, cols, res)).cuda()
    start_d = torch.from_numpy(np.stack([sr, sc], 1).astype(np.int32)).cuda()
    warm = max(a.warmup, 1)
    ncase = a.steps + warm
    spd, drn = seasonal_wind_conditions(ncase, seed=11)
    if world > 1 and not case_parallel:
        D.native_comm()
    ring = StepRing(torch, shape, n_rank, min(a.streams, 4), a.mode, 1 if case_parallel else world)
    total = torch.zeros(1, dtype=torch.int64, device="cuda")
    stage = {"wind_interp_ms": 0.0, "updraft_ms": 0.0, "potential_ms": 0.0}
    iters = []

    def one_case(i, timed):
        if case_parallel and i % world != rank:
            return
        t0 = time.perf_counter()
        ws, wd, _ = wind_rasters(w, torch, case=(float(spd[i]), float(drn[i])))
        t1 = time.perf_counter()
        up = layers.updraft_fields(z, res, ws, wd, 0.75, want=("updraft",))["updraft"]
        torch.cuda.synchronize()
        t2 = time.perf_counter()
        pot, stats = solve_potential_device(up, 0.0, sharded=(world > 1 and not case_parallel))
        torch.cuda.synchronize()
        t3 = time.perf_counter()
        fields = mm.interleave_fields(up, pot)
        ev = torch.cuda.Event()
        ev.record()
        ring.issue(i, fields, start_d, None, a.seed + 1 + i, lo, total, after=ev)   # asynchronous: overlaps the next case's fields
        if timed:
            stage["wind_interp_ms"] += (t1 - t0) * 1e3
            stage["updraft_ms"] += (t2 - t1) * 1e3
            stage["potential_ms"] += (t3 - t2) * 1e3
            iters.append(stats["iterations"])

    for i in range(warm):
        one_case(i, False)
    ring.join()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    total.zero_()
    sampler = ClockSampler(int(os.environ.get("LOCAL_RANK", "0")))
    sampler.start()
    t0 = time.perf_counter()
    for i in range(warm, ncase):
        one_case(i, True)
    ring.join()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    wall = time.perf_counter() - t0
    sampler.stop_flag = True
    t = torch.tensor([wall, float(total.item())], dtype=torch.float64, device="cuda")
    if world > 1:
        m = t.clone(); dist.all_reduce(m, op=dist.ReduceOp.MAX)
        s_ = t.clone(); dist.all_reduce(s_, op=dist.ReduceOp.SUM)
        wall, steps_all = float(m[0]), float(s_[1])
    else:
        steps_all = float(t[1])
    if rank == 0:
        mine = max(1, len(iters))
        line = {"metric": METRIC, "value": steps_all / wall, "unit": UNIT, "n_gpus": world, "steps": a.steps, "warmup": warm,
                "ms_per_step": wall / a.steps * 1e3, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
                "dtype": "f64 probabilities on f32 fields", "data": "synthetic",
                "config": {"workload": workload_text(w, world), "grid": [rows, cols], "tracks_per_step": n_total,
                           "seasonal_mode": "case_parallel" if case_parallel else "sharded",
                           "step": "one wind case: site interpolation, per-cell-wind stencil, threshold, potential solve, "
                                   "stepping, presence map (whole pipeline inside the timed region)",
                           "track_steps_per_step": steps_all / a.steps},
                "fields": {"per_case_ms_rank0": {k: v / mine for k, v in stage.items()}, "iterations": iters},
                "e2e": None, "gpu_launches": None, "clocks": sampler.summary()}
        emit(line)
    if world > 1:
        dist.barrier()
        D.destroy_native_comm()
        dist.destroy_process_group()
    return 0


# ---------------------------------------------------------------------------------------------------------------------
# the reference arm: the reference's own CPU implementation of the path on the host cores
# ---------------------------------------------------------------------------------------------------------------------
def reference_fields(a, w):
    """Fields for the CPU arm without loading the product into this process.  Stage 1: the numpy restatement of the
    reference (oracle_np; the reference's own np.vectorize threshold needs minutes at 3e7 cells).  Stage 2: the
    reference's algorithm (SuperLU, oracle_np.solve_potential) where it can run (<= 1.2e6 cells); above that no CPU solve
    exists, so a CHILD process runs the GPU solve and hands the raster over as a file.  No fallback."""
    from oracle import oracle_np as O
    from ssrs_b200.synth import synthetic_dem
    rows, cols, res = w["rows"], w["cols"], w["res"]
    if w["wind"] != "uniform":
        raise SystemExit("bench.py --impl reference runs the uniform-wind workloads (config2/config3)")
    z = synthetic_dem(rows, cols, res)
    t0 = time.perf_counter()
    _, _, _, K = O.updraft_pipeline(z, res, 10.0, 270.0, 0.75)
    stencil_s = time.perf_counter() - t0
    U = K.astype(np.float32)
    if rows * cols <= 1_200_000:
        t0 = time.perf_counter()
        P = O.solve_potential(U.astype(np.float64), 0.0)
        return U, P, f"reference algorithm (SuperLU on the row-normalised system), {time.perf_counter() - t0:.1f} s", stencil_s
    with tempfile.TemporaryDirectory() as tmp:
        path = os.path.join(tmp, "fields.npz")
        cmd = [sys.executable, os.path.abspath(__file__), "--emit-fields", path, "--workload", w["name"], "--rows", str(rows),
               "--cols", str(cols), "--resolution", str(res)]
        env = {k: v for k, v in os.environ.items() if k not in ("RANK", "WORLD_SIZE", "LOCAL_RANK")}
        r = subprocess.run(cmd, capture_output=True, text=True, env=env)
        if r.returncode != 0 or not os.path.exists(path):
            raise RuntimeError("the reference arm needs the potential at this grid size from the GPU solver (the reference's "
                               "SuperLU solve cannot run above ~1e6 cells) and the child process failed:\n" + r.stderr[-2000:])
        with np.load(path) as f:
            P = f["potential"]
    return U, P, "ssrs_b200 GPU solve in a child process (the reference's SuperLU solve cannot run at this size)", stencil_s


def reference_arm(a, w, world):
    from oracle import ref_loader
    threads = os.cpu_count() or 1
    U, P, pot_src, stencil_s = reference_fields(a, w)
    shape = (w["rows"], w["cols"])
    n_total = w["n_total"]
    sr, sc = start_cells(w, n_total, a.seed)
    use_ref = ref_loader.available()
    # bounded sample per step: one track per process costs `dt0` of wall time (the reference's pool pickles its closure —
    # the two rasters — once per task chunk, simulator.py:361-369, which dominates small samples); a step gets as many
    # rounds of that as fit a total budget of ~150 s for the whole run
    if use_ref:
        probe = min(n_total, threads)
        _, steps0, dt0, procs = cpu_reference_rate(U, P, sr, sc, probe, threads, a.seed)
        per_step = 150.0 / max(1, a.steps + min(a.warmup, 1))
        n_sample = a.cpu_sample_tracks or int(min(n_total, probe * max(1, int(per_step / max(dt0, 1e-3)))))
        run = lambda lo: cpu_reference_rate(U, P, sr[lo:], sc[lo:], n_sample, threads, a.seed)[:3]
        kind, how = "reference", (f"unmodified ssrs/movmodel.py generate_simulated_tracks through multiprocess.Pool({threads}).map "
                                  f"(ssrs/simulator.py:360-369)")
    else:
        n_sample = a.cpu_sample_tracks or min(n_total, 2048 * threads)
        run = lambda lo: cpu_port_rate(U, P, shape, sr[lo:], sc[lo:], n_sample, threads, a.seed)
        kind, how = "port", f"oracle/ssrs_oracle.c with OpenMP on {threads} threads (reference modules unavailable)"
    for _ in range(min(a.warmup, 1)):
        run(0)
    steps_done, secs = 0, 0.0
    for i in range(a.steps):
        lo = (i * n_sample) % max(1, n_total - n_sample + 1)
        _, nsteps, dt = run(lo)
        steps_done += nsteps
        secs += dt
    value = steps_done / secs
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": a.steps,
            "warmup": a.warmup, "ms_per_step": secs / a.steps * 1e3, "higher_is_better": True, "scaling": w["scaling"],
            "vs_baseline": None, "dtype": "f64 probabilities on f32 fields", "data": "synthetic",
            "config": workload_config(w, world),
            "run": {"potential": pot_src, "stencil_numpy_s": stencil_s},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": kind,
                             "sample": f"{n_sample} of the step's tracks per step ({steps_done // max(1, a.steps)} track-steps), {how}"},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    emit(line)
    return 0


if __name__ == "__main__":
    sys.exit(main())
