#!/usr/bin/env python
"""bench.py -- cell-updates/s of the SSPRK3 dycore step (BASELINE.json metric) on N B200s of one node.

    python bench.py --gpus 1 --steps 10 --warmup 3
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
           bench.py --gpus N --steps K --warmup W
    python bench.py --impl reference ...      # the reference's own CPU implementation on the host cores

Workload (BASELINE.json configs[1]): dry Euler dycore (WENO5 + SSPRK3; the water-vapour tracer the reference cannot
run without, so N = 6 variables), synthetic supercell-shaped 512 x 512 x 128 fp64 grid PER GPU (weak scaling: the
global grid is nproc_x*512 x nproc_y*512 x 128 with the reference's x-y decomposition).  A "step" is one
dycore.time_step(coupler, dt) = coupler->dycore conversion, 3 fused RK stages, dycore->coupler conversion.
The state (6 fields x 268 MB) is far larger than L2 (126 MB), so no flush is needed between iterations.

One JSON line on stdout (rank 0).  `value` = device-timed throughput with the state resident in HBM; `e2e` = the same
metric through the host-buffer C-ABI call (pinned host arrays, H2D + step + D2H inside the timed region).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

NX_LOC, NY_LOC, NZ = 512, 512, 128
DX = 1000.0
ZLEN = 20000.0
NUM_TRACERS = 1
NVAR = 5 + NUM_TRACERS
BYTES_PER_CELL_UPDATE = 64 * NVAR          # 8 N doubles per cell per SSPRK3 step
CPU_SAMPLE = dict(nx=256, ny=256, nz=128)  # bounded sample of the same workload for the CPU arm: config 2's dz, dt and levels,
                                           # a quarter of its columns (about 10 s of work for 16 host cores at 6 steps)


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi samples during the timed region (recipe of B200_PROFILING.md)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "20"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for l in self.proc.stdout:
            self.lines.append(l.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, power, reasons = [], None, [], set()
        for l in self.lines:
            f = [x.strip() for x in l.split(",")]
            if len(f) < 9:
                continue
            try:
                if float(f[1]) < 600.0:      # idle sample (SM clock parked at ~120 MHz before the first kernel): not "under load"
                    continue
                sm.append(float(f[1])); mx = float(f[2]); power.append(float(f[3]))
            except ValueError:
                continue
            for name, v in zip(["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"], f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx, "samples": len(sm),
                "power_w_max": max(power) if power else None, "reasons": sorted(reasons)}


def decomposition(nranks, rank):
    """The reference's x-y rank grid (model/core/coupler.h:127-179) for a 3-D run."""
    import math
    npy = int(math.ceil(math.sqrt(nranks)))
    while npy >= 1 and nranks % npy != 0:
        npy -= 1
    npx = nranks // npy
    return npx, npy, rank % npx, rank // npx


def ref_driver(omp=True):
    exe = os.path.join(ROOT, "oracle", "_ref", "ref_driver_omp" if omp else "ref_driver")
    return exe if os.path.exists(exe) else None


def cpu_reference_run(steps, warmup, sample=CPU_SAMPLE):
    """Time the reference's own CPU implementation of the same step (dycore only, vapour tracer) on the host cores.
    Returns (cell_updates_per_s, cores, kind, sample_description)."""
    ncores = os.cpu_count() or 1
    nx, ny, nz = sample["nx"], sample["ny"], sample["nz"]
    desc = "dycore.time_step, supercell %dx%dx%d fp64 (config 2's dx=dy=1000 m, dz, dt and 128 levels; a quarter of its columns), %d steps" % (nx, ny, nz, steps)
    exe = ref_driver(omp=True)
    if exe is not None:
        env = dict(os.environ, OMP_NUM_THREADS=str(ncores), GATOR_INITIAL_MB="4096")

        def run(nsteps):
            out = subprocess.run([exe, "run", "nx=%d" % nx, "ny=%d" % ny, "nz=%d" % nz, "xlen=%g" % (nx * DX),
                                  "ylen=%g" % (ny * DX), "zlen=%g" % ZLEN, "tracers=vapor", "steps=%d" % nsteps, "time=1"],
                                 env=env, capture_output=True, text=True, check=True).stdout
            j = [json.loads(l) for l in out.splitlines() if l.startswith("{") and "seconds" in l][0]
            return j["seconds"]
        # the driver times its step loop only; warm-up = a separate short run that pages the binary in
        if warmup > 0:
            run(1)
        sec = run(steps)
        return nx * ny * nz * steps / sec, ncores, "reference", desc + " (oracle/_ref/ref_driver_omp, YAKL OpenMP backend, %d threads)" % ncores
    # no compiled reference on this box: the plain-C port, one core
    import numpy as np
    import _oracle as O
    from miniweatherml_b200.supercell import supercell_column
    nx, ny, nz = 48, 48, 64
    bg, col = supercell_column(nz, ZLEN)
    f = np.stack([np.broadcast_to(col[n][:, None, None], (nz, ny, nx)) for n in
                  ["density_dry", "uvel", "vvel", "wvel", "temp", "water_vapor"]]).copy()
    O.perturb_thermal(f[4], 0, 0, DX, DX, ZLEN / nz, nx * DX, ny * DX)
    p = O.make_params(nx, ny, nz, nx * DX, ny * DX, ZLEN, 1)
    dt = 0.6 * min(DX, ZLEN / nz) / 430.0
    O.dycore_step(p, bg, f, dt, steps=max(warmup, 1) if warmup else 0)
    t0 = time.time()
    O.dycore_step(p, bg, f, dt, steps=steps)
    sec = time.time() - t0
    return nx * ny * nz * steps / sec, 1, "port", "oracle/mw_oracle.c dycore step, supercell %dx%dx%d fp64, %d steps, 1 thread" % (nx, ny, nz, steps)


def workload_config(world, dt=None):
    """The `config` object of the benchmark line: the same for both arms (the CPU arm adds the sample it timed)."""
    npx, npy, _, _ = decomposition(world, 0)
    nxg, nyg = NX_LOC * npx, NY_LOC * npy
    if dt is None:
        dt = 0.6 * min(DX, ZLEN / NZ) / 430.0                      # DYC:70-77
    return {"workload": "BASELINE configs[1]: dry Euler dycore (WENO5 + SSPRK3, vapour tracer, N=6), synthetic "
                        "supercell, %dx%dx%d fp64 per GPU" % (NX_LOC, NY_LOC, NZ),
            "global_grid": [nxg, nyg, NZ], "decomposition": "%dx%d (x,y)" % (npx, npy), "dt": dt,
            "l2_policy": "state (1.6 GB per GPU) larger than L2, no flush needed"}


def run_reference_arm(args, rank):
    if rank != 0:
        return
    steps = max(1, min(args.steps, 6))
    v, cores, kind, desc = cpu_reference_run(steps, min(args.warmup, 1))
    line = {"impl": "reference", "metric": "cell-updates/s per SSPRK3 step", "value": v, "unit": "cell-updates/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": None,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": dict(workload_config(max(args.gpus, 1)), sample=desc),
            "cpu_baseline": {"value": v, "unit": "cell-updates/s", "cores": cores, "kind": kind, "sample": desc},
            "e2e": {"value": v, "unit": "cell-updates/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def run_config3(args):
    """BASELINE configs[2], weak-scaling form: supercell + Kessler microphysics (+ sponge layer + column
    nudging, the canonical step loop of experiments/supercell_example/driver.cpp:66-79) on 1024 x 1024 x 128 fp64 cells
    per GPU, N = 8 variables.  Prints one extra JSON report line: full-step and dycore-only cell-updates/s."""
    import torch
    import torch.distributed as dist
    import miniweatherml_b200 as mw
    from miniweatherml_b200.supercell import supercell_column
    from miniweatherml_b200 import distributed as mwd
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    nxl = int(os.environ.get("MW_C3_NX", "1024")); nyl = int(os.environ.get("MW_C3_NY", "1024")); nz = NZ
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    comm = None
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
        comm = mwd.create_comm(dist, rank, world, dev)
    npx, npy, px, py = decomposition(world, rank)
    nxg, nyg = nxl * npx, nyl * npy
    T = 3
    cfg = mw.make_config(nxl, nyl, nz, nxg * DX, nyg * DX, ZLEN, T, nx_glob=nxg, ny_glob=nyg, i_beg=px * nxl,
                         j_beg=py * nyl, nproc_x=npx, nproc_y=npy, px=px, py=py)
    dy = mw.Dycore(cfg)
    if comm is not None:
        dy.attach_comm(comm)
    bg, col = supercell_column(nz, ZLEN)
    dy.set_background(bg)
    names = ["density_dry", "uvel", "vvel", "wvel", "temp", "water_vapor"]
    fields = [torch.tensor(col[n], device=dev)[:, None, None].expand(nz, nyl, nxl).contiguous() for n in names]
    fields += [torch.zeros((nz, nyl, nxl), device=dev, dtype=torch.float64) for _ in range(2)]
    precl = torch.zeros((nyl, nxl), device=dev, dtype=torch.float64)
    f5 = [fields[0], fields[1], fields[2], fields[4], fields[5]]
    column = mw.column_average(f5, nxy_glob=nxg * nyg, comm=comm)
    mw.perturb_temperature(fields[4], px * nxl, py * nyl, DX, DX, ZLEN / nz, nxg * DX, nyg * DX)
    dt = dy.compute_time_step()
    dz = ZLEN / nz
    dy.enable_timing(True)

    def step():
        dy.time_step(fields, dt)
        mw.kessler_step(fields[4], fields[0], fields[5], fields[6], fields[7], precl, dz, dt, comm=comm)
        mw.sponge_layer(fields, dz, ZLEN, dt, nxy_glob=nxg * nyg, comm=comm)
        mw.nudge_to_column(f5, column, dt, nxy_glob=nxg * nyg, comm=comm)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        step()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        step()
    e1.record()
    barrier()
    s_ms, s_n, dyc_ms = dy.last_timing()
    t = torch.tensor([e0.elapsed_time(e1), dyc_ms], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms, dyc_ms = t.tolist()
    finite = all(bool(torch.isfinite(f).all()) for f in fields)
    if rank == 0:
        peak, peak_src = measured_peak()
        cells = nxg * nyg * nz
        full = cells * args.steps / (ms * 1e-3)
        dyc = cells / (dyc_ms * 1e-3)
        bytes_full = 64 * 8 + 72           # dycore 64 N (N = 8) + Kessler 72 B per cell
        print(json.dumps({
            "report": "config3", "metric": "cell-updates/s per SSPRK3 step", "unit": "cell-updates/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "value": full, "ms_per_step": ms / args.steps,
            "dycore_only_value": dyc, "dycore_ms_per_step": dyc_ms, "stage_kernel_ms": s_ms / max(s_n, 1),
            "scaling": "weak", "dtype": "f64", "data": "synthetic",
            "config": {"workload": "BASELINE configs[2] (weak form): supercell + Kessler + sponge + nudging, N=8, "
                                   "%dx%dx%d fp64 per GPU" % (nxl, nyl, nz),
                       "global_grid": [nxg, nyg, nz], "decomposition": "%dx%d (x,y)" % (npx, npy), "dt": dt,
                       "state_finite": finite},
            "hbm_frac_full_step": full / world * bytes_full / 1e9 / peak,
            "hbm_frac_dycore": dyc / world * 64 * 8 / 1e9 / peak, "peak": peak, "peak_source": peak_src,
            "mem_gb": torch.cuda.max_memory_allocated() / 1e9}), flush=True)
    dy.close()
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--workload", default="config2", choices=["config2", "config3"],
                    help="config2 (default, the contract line): dry dycore 512x512x128 per GPU; config3: supercell + Kessler "
                         "+ sponge + nudging (N=8), 1024x1024x128 per GPU (SURVEY 8(d) weak-scaling size), extra report line")
    args = ap.parse_args()
    if args.workload == "config3" and args.impl != "reference":
        return run_config3(args)
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference_arm(args, rank)
        return
    assert args.warmup >= 3 or args.steps <= 2, "timing rules: at least 3 warm-up steps"

    import numpy as np
    import torch
    import torch.distributed as dist
    import miniweatherml_b200 as mw
    from miniweatherml_b200.supercell import supercell_column
    from miniweatherml_b200 import distributed as mwd

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    comm = None
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
        comm = mwd.create_comm(dist, rank, world, dev)
    npx, npy, px, py = decomposition(world, rank)
    nxg, nyg = NX_LOC * npx, NY_LOC * npy
    cfg = mw.make_config(NX_LOC, NY_LOC, NZ, nxg * DX, nyg * DX, ZLEN, NUM_TRACERS, nx_glob=nxg, ny_glob=nyg,
                         i_beg=px * NX_LOC, j_beg=py * NY_LOC, nproc_x=npx, nproc_y=npy, px=px, py=py)
    dy = mw.Dycore(cfg)
    if comm is not None:
        dy.attach_comm(comm)
    bg, col = supercell_column(NZ, ZLEN)
    dy.set_background(bg)
    names = ["density_dry", "uvel", "vvel", "wvel", "temp", "water_vapor"]
    fields = [torch.tensor(col[n], device=dev)[:, None, None].expand(NZ, NY_LOC, NX_LOC).contiguous() for n in names]
    mw.perturb_temperature(fields[4], px * NX_LOC, py * NY_LOC, DX, DX, ZLEN / NZ, nxg * DX, nyg * DX)
    dt = dy.compute_time_step()
    dy.enable_timing(True)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()                  # runs through warm-up and the timed region (same load); idle samples are dropped
    for _ in range(args.warmup):
        dy.time_step(fields, dt)
    barrier()
    l0 = dy.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    stage_ms, n_stage = 0.0, 0
    e0.record()
    for _ in range(args.steps):
        dy.time_step(fields, dt)
    e1.record()
    barrier()
    # per-launch time of the dominant kernel (k_stage), CUDA events recorded inside the library on the same stream,
    # taken over the last timed step
    s_ms, s_n, _ = dy.last_timing()
    launches = dy.launch_count() - l0
    clocks = sampler.stop() if rank == 0 else None
    ms = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms = ms.item()
    finite = all(bool(torch.isfinite(f).all()) for f in fields)
    cells_glob = nxg * nyg * NZ
    value = cells_glob * args.steps / (ms * 1e-3)

    # ---- end to end through the host-buffer entry point (pinned host memory, H2D + step + D2H per step) ----------
    e2e_value = None
    if args.e2e_steps > 0:
        host = [torch.empty((NZ, NY_LOC, NX_LOC), dtype=torch.float64).pin_memory() for _ in names]
        for h, f in zip(host, fields):
            h.copy_(f)
        hnp = [h.numpy() for h in host]
        dy.time_step_host(hnp, dt)                       # warm-up (allocates the staging buffers)
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.e2e_steps):
            dy.time_step_host(hnp, dt)
        barrier()
        e2e_s = torch.tensor([time.perf_counter() - t0], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(e2e_s, op=dist.ReduceOp.MAX)
        e2e_value = cells_glob * args.e2e_steps / e2e_s.item()
    field_bytes = NZ * NY_LOC * NX_LOC * 8

    if rank == 0:
        peak, peak_src = measured_peak()
        cells_loc = NX_LOC * NY_LOC * NZ
        # algorithmic bytes of one stage launch: stage 1 reads q, writes q1 (2N doubles/cell), stages 2,3 read q,q0 and
        # write (3N): average over the three launches of a step = 8N/3 doubles per cell
        alg_bytes_per_launch = cells_loc * 8.0 * NVAR * 8.0 / 3.0
        k_ms = s_ms / max(s_n, 1)
        achieved = alg_bytes_per_launch / (k_ms * 1e-3) / 1e9
        traffic = None
        fp64 = None
        pj = os.path.join(ROOT, "profiles", "stage_kernel_bench.json")
        if os.path.exists(pj):
            try:
                pr = json.load(open(pj))
                traffic = pr.get("dram_bytes_per_launch")
                # FP64-pipe view of the same launch: fp64 instructions per cell-stage counted by ncu (committed capture)
                # x cells / live CUDA-event duration, against the DFMA issue rate measured by tools/fp64_peak.cu
                fi = pr.get("fp64_thread_instr_per_cell_stage")
                pk = pr.get("fp64_peak_thread_instr_per_s")
                if fi and pk:
                    ach = fi * cells_loc / (k_ms * 1e-3)
                    fp64 = {"bound": "fp64-pipe", "achieved": ach / 1e12, "peak": pk / 1e12, "unit": "T fp64-instr/s",
                            "frac": ach / pk, "fp64_instr_per_cell_stage": fi,
                            "ncu_pipe_fp64_cycles_active_pct": pr.get("ncu_pipe_fp64_cycles_active_pct"),
                            "peak_source": "measured DFMA issue rate (tools/fp64_peak.cu, profiles/r01a_device_peaks.jsonl)"}
            except Exception:
                pass
        line = {"metric": "cell-updates/s per SSPRK3 step", "value": value, "unit": "cell-updates/s", "n_gpus": world,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": dict(workload_config(world, dt), state_finite=finite),
                "clocks": clocks,
                "e2e": {"value": e2e_value, "unit": "cell-updates/s", "h2d_bytes_per_step": NVAR * field_bytes,
                        "d2h_bytes_per_step": NVAR * field_bytes, "steps": args.e2e_steps,
                        "api": "mw_dycore_time_step_host (pinned host buffers)"},
                "gpu_launches": int(launches),
                "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                             "traffic": traffic, "kernel": "k_stage_ws<1,16,8,SEG>", "kernel_ms": k_ms, "peak_source": peak_src,
                             "algorithmic_bytes_per_launch": alg_bytes_per_launch,
                             "note": "the fused stage kernel is FP64-pipe-bound, not HBM-bound (DESIGN.md section 4): "
                                     "the binding roofline is reported under fp64_pipe",
                             "fp64_pipe": fp64,
                             "whole_step_frac": value / world * BYTES_PER_CELL_UPDATE / 1e9 / peak}}
        if not args.no_cpu_baseline:
            try:
                v, cores, kind, desc = cpu_reference_run(4, 1)
                line["cpu_baseline"] = {"value": v, "unit": "cell-updates/s", "cores": cores, "kind": kind, "sample": desc}
            except Exception as e:                                  # the baseline is a report, never a reason to fail
                line["cpu_baseline"] = {"value": None, "unit": "cell-updates/s", "cores": 0, "kind": "unavailable",
                                        "sample": "failed: %r" % (e,)}
        print(json.dumps(line), flush=True)
    dy.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
