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
CPU_SAMPLE = dict(nx=256, ny=256, nz=128)  # in-line cpu_baseline: config 2's dz, dt and levels, a quarter of its columns
                                           # (about 10 s of work for 16 host cores)
CPU_FULL = dict(nx=NX_LOC, ny=NY_LOC, nz=NZ)   # --impl reference: the whole per-GPU grid of the benchmark line
REF_MAX_TIMED_STEPS = 4                    # ~6 s per step on 16 cores: 1 warm-up + 4 timed steps stay within two minutes


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


NCU_METRICS = ["smsp__sass_thread_inst_executed_op_dfma_pred_on.sum", "smsp__sass_thread_inst_executed_op_dmul_pred_on.sum",
               "smsp__sass_thread_inst_executed_op_dadd_pred_on.sum", "smsp__thread_inst_executed.sum",
               "dram__bytes_read.sum", "dram__bytes_write.sum", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_elapsed",
               "gpu__time_duration.sum"]


def counters_child():
    """Run by `ncu` from measure_counters(): two dycore steps of the benchmark workload on one GPU, nothing else."""
    import torch
    import miniweatherml_b200 as mw
    from miniweatherml_b200.supercell import supercell_column
    dev = torch.device("cuda", 0)
    cfg = mw.make_config(NX_LOC, NY_LOC, NZ, NX_LOC * DX, NY_LOC * DX, ZLEN, NUM_TRACERS)
    dy = mw.Dycore(cfg)
    bg, col = supercell_column(NZ, ZLEN)
    dy.set_background(bg)
    names = ["density_dry", "uvel", "vvel", "wvel", "temp", "water_vapor"]
    fields = [torch.tensor(col[n], device=dev)[:, None, None].expand(NZ, NY_LOC, NX_LOC).contiguous() for n in names]
    mw.perturb_temperature(fields[4], 0, 0, DX, DX, ZLEN / NZ, NX_LOC * DX, NY_LOC * DX)
    dt = dy.compute_time_step()
    for _ in range(2):
        dy.time_step(fields, dt)
    torch.cuda.synchronize()
    dy.close()


def measure_counters(timeout_s=240):
    """Instruction mix and DRAM traffic of the dominant kernel, counted on THIS box (outside the timed region): the
    three stage launches of one step under `ncu --metrics ...` in a child process.  Returns a dict or {"error": ...}."""
    import csv
    import io
    import shutil
    ncu = shutil.which("ncu") or "/usr/local/cuda/bin/ncu"
    if not os.path.exists(ncu):
        return {"error": "ncu not found"}
    cmd = [ncu, "--metrics", ",".join(NCU_METRICS), "--clock-control", "none", "-k", "regex:k_stage", "-s", "3", "-c", "3",
           "--csv", sys.executable, os.path.abspath(__file__), "--counters-child"]
    try:
        env = dict(os.environ)
        for k in ("RANK", "WORLD_SIZE", "LOCAL_RANK", "MASTER_ADDR", "MASTER_PORT"):
            env.pop(k, None)
        out = subprocess.run(cmd, capture_output=True, text=True, timeout=timeout_s, env=env).stdout
    except Exception as e:
        return {"error": "ncu child failed: %r" % (e,)}
    rows = [r for r in csv.reader(io.StringIO(out)) if len(r) > 5]
    hdr = next((r for r in rows if "Metric Name" in r), None)
    if hdr is None:
        return {"error": "no metrics in the ncu output: " + out[-300:].replace("\n", " ")}
    iN, iV, iK, iI = hdr.index("Metric Name"), hdr.index("Metric Value"), hdr.index("Kernel Name"), hdr.index("ID")
    acc, ids, kern = {}, set(), None
    for r in rows:
        if r is hdr or len(r) <= max(iN, iV):
            continue
        try:
            v = float(r[iV].replace(",", ""))
        except ValueError:
            continue
        acc[r[iN]] = acc.get(r[iN], 0.0) + v
        ids.add(r[iI]); kern = r[iK]
    n = max(len(ids), 1)
    if "smsp__thread_inst_executed.sum" not in acc:
        return {"error": "metrics missing: " + ",".join(sorted(acc))}
    cells = NX_LOC * NY_LOC * NZ
    f64 = sum(acc.get(m, 0.0) for m in NCU_METRICS[:3])
    return {"launches_counted": n, "kernel": kern.split("(")[0] if kern else None,
            "fp64_thread_instr_per_cell_stage": f64 / n / cells,
            "thread_instr_per_cell_stage": acc["smsp__thread_inst_executed.sum"] / n / cells,
            "dram_bytes_per_launch": (acc.get("dram__bytes_read.sum", 0.0) + acc.get("dram__bytes_write.sum", 0.0)) / n,
            "pipe_fp64_cycles_active_pct": acc.get("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_elapsed", 0.0) / n,
            "ncu_kernel_ms_cold": acc.get("gpu__time_duration.sum", 0.0) / n / 1e6,
            "source": "ncu --metrics (this run, this box), mean of the %d stage launches of one step" % n}


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
    frac = (nx * ny * nz) / float(NX_LOC * NY_LOC * NZ)
    desc = "dycore.time_step, supercell %dx%dx%d fp64 (config 2's dx=dy=1000 m, dz, dt and 128 levels; %s of its per-GPU grid), %d timed steps after %d warm-up" % (
        nx, ny, nz, "all" if frac == 1.0 else "%.3g" % frac, steps, warmup)
    exe = ref_driver(omp=True)
    if exe is not None:
        env = dict(os.environ, OMP_NUM_THREADS=str(ncores), GATOR_INITIAL_MB="4096")
        out = subprocess.run([exe, "run", "nx=%d" % nx, "ny=%d" % ny, "nz=%d" % nz, "xlen=%g" % (nx * DX),
                              "ylen=%g" % (ny * DX), "zlen=%g" % ZLEN, "tracers=vapor", "steps=%d" % steps,
                              "warmup=%d" % warmup, "time=1"], env=env, capture_output=True, text=True, check=True).stdout
        sec = [json.loads(l) for l in out.splitlines() if l.startswith("{") and "seconds" in l][0]["seconds"]
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
    """The reference's own CPU implementation on the box's host cores, on the benchmark line's per-GPU grid (the whole
    512 x 512 x 128 block: at N = 1 that IS the configuration; at N > 1 the global grid is N such blocks and the CPU arm
    times one of them, throughput-normalised).  Timed steps are capped so the run ends within a couple of minutes."""
    if rank != 0:
        return
    steps = max(1, min(args.steps, REF_MAX_TIMED_STEPS))
    warm = 1 if args.warmup > 0 else 0
    t0 = time.time()
    v, cores, kind, desc = cpu_reference_run(steps, warm, sample=CPU_FULL if ref_driver() else CPU_SAMPLE)
    line = {"impl": "reference", "metric": "cell-updates/s per SSPRK3 step", "value": v, "unit": "cell-updates/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "steps_timed": steps, "warmup_run": warm,
            "ms_per_step": 1e3 * NX_LOC * NY_LOC * NZ / v if kind == "reference" else None,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": workload_config(max(args.gpus, 1)),
            "cpu_baseline": {"value": v, "unit": "cell-updates/s", "cores": cores, "kind": kind, "sample": desc},
            "e2e": {"value": v, "unit": "cell-updates/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0, "wall_s": time.time() - t0}
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
    nz = NZ
    npx, npy, px, py = decomposition(world, rank)
    if args.scaling == "strong":       # BASELINE configs[2] as written: ONE global 2048 x 2048 x 128 grid on 1/2/4/8 GPUs
        nglob = int(os.environ.get("MW_C3_GLOB", "2048"))
        assert nglob % npx == 0 and nglob % npy == 0
        nxl, nyl = nglob // npx, nglob // npy
    else:                              # weak form: 1024 x 1024 x 128 per GPU
        nxl = int(os.environ.get("MW_C3_NX", "1024")); nyl = int(os.environ.get("MW_C3_NY", "1024"))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    comm = None
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
        comm = mwd.create_comm(dist, rank, world, dev)
    nxg, nyg = nxl * npx, nyl * npy
    T = 3
    cfg = mw.make_config(nxl, nyl, nz, nxg * DX, nyg * DX, ZLEN, T, nx_glob=nxg, ny_glob=nyg, i_beg=px * nxl,
                         j_beg=py * nyl, nproc_x=npx, nproc_y=npy, px=px, py=py)
    wl = {"workload": "BASELINE configs[2] (%s form): supercell + Kessler + sponge + nudging, N=8, %dx%dx%d fp64 per GPU"
                      % (args.scaling, nxl, nyl, nz),
          "global_grid": [nxg, nyg, nz], "decomposition": "%dx%d (x,y)" % (npx, npy)}
    # device memory this needs per GPU: 3 RK registers of N haloed fields, 4 T tracer flux / FCT arrays, N coupler fields
    need_gb = (3 * 8 * nz * (nyl + 6) * (nxl + 6) + (4 * T + 8) * nz * nyl * nxl) * 8 / 1e9
    try:
        dy = mw.Dycore(cfg)
        if comm is not None:
            dy.attach_comm(comm)
        bg, col = supercell_column(nz, ZLEN)
        dy.set_background(bg)
        names = ["density_dry", "uvel", "vvel", "wvel", "temp", "water_vapor"]
        fields = [torch.tensor(col[n], device=dev)[:, None, None].expand(nz, nyl, nxl).contiguous() for n in names]
        fields += [torch.zeros((nz, nyl, nxl), device=dev, dtype=torch.float64) for _ in range(2)]
        precl = torch.zeros((nyl, nxl), device=dev, dtype=torch.float64)
    except (mw.MwError, torch.OutOfMemoryError) as e:
        # the configuration does not fit this GPU count: say so instead of measuring something smaller
        if rank == 0:
            print(json.dumps({"report": "config3", "n_gpus": world, "scaling": args.scaling, "config": wl, "value": None,
                              "oom": True, "needed_gb_per_gpu": need_gb,
                              "hbm_gb": torch.cuda.get_device_properties(dev).total_memory / 1e9,
                              "error": str(e)[:300]}), flush=True)
        if world > 1:
            dist.destroy_process_group()
        return
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
            "scaling": args.scaling, "dtype": "f64", "data": "synthetic",
            "config": dict(wl, dt=dt, state_finite=finite), "needed_gb_per_gpu": need_gb,
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
    ap.add_argument("--e2e-steps", type=int, default=5)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-counters", action="store_true", help="skip the ncu child that counts instructions / DRAM bytes")
    ap.add_argument("--no-parity", action="store_true", help="skip the fixture parity check after the measurement")
    ap.add_argument("--counters-child", action="store_true", help=argparse.SUPPRESS)
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="config3 only: weak = 1024x1024x128 per GPU, strong = global 2048x2048x128 (BASELINE configs[2])")
    ap.add_argument("--workload", default="config2", choices=["config2", "config3"],
                    help="config2 (default, the contract line): dry dycore 512x512x128 per GPU; config3: supercell + Kessler "
                         "+ sponge + nudging (N=8), 1024x1024x128 per GPU (SURVEY 8(d) weak-scaling size), extra report line")
    args = ap.parse_args()
    if args.counters_child:
        return counters_child()
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

    # ---- N-rank correctness against committed reference fixtures (untimed, every rank takes part) --------------------
    parity = None
    if not args.no_parity:
        from miniweatherml_b200.parity import fixture_parity
        parity = []
        for name in ("box3d_vapor_dycore5.npz", "building_city_loop6.npz"):
            try:
                parity.append(fixture_parity(dist, comm, rank, world, dev, name))
            except Exception as e:                                  # a failed check is reported, never hidden
                parity.append({"fixture": name, "ok": False, "error": repr(e)})

    if rank == 0:
        peak, peak_src = measured_peak()
        cells_loc = NX_LOC * NY_LOC * NZ
        # algorithmic bytes of one stage launch: stage 1 reads q, writes q1 (2N doubles/cell), stages 2,3 read q,q0 and
        # write (3N): average over the three launches of a step = 8N/3 doubles per cell
        alg_bytes_per_launch = cells_loc * 8.0 * NVAR * 8.0 / 3.0
        k_ms = s_ms / max(s_n, 1)
        achieved = alg_bytes_per_launch / (k_ms * 1e-3) / 1e9
        # instruction mix / DRAM traffic (ncu child) and the DFMA issue rate (probe kernel): both measured now, on this box
        counters = {"error": "skipped (--no-counters)"} if (args.no_counters or world > 1) else measure_counters()
        traffic = counters.get("dram_bytes_per_launch")
        fp64 = None
        try:
            pk = mw.probe_fp64_rate()
            fi = counters.get("fp64_thread_instr_per_cell_stage")
            fp64 = {"bound": "fp64-pipe", "peak": pk / 1e12, "unit": "T fp64-instr/s",
                    "peak_source": "mw_probe_fp64_rate: DFMA issue rate measured in this run"}
            if fi:
                ach = fi * cells_loc / (k_ms * 1e-3)
                fp64.update({"achieved": ach / 1e12, "frac": ach / pk, "fp64_instr_per_cell_stage": fi,
                             "other_instr_per_cell_stage": counters["thread_instr_per_cell_stage"] - fi,
                             "ncu_pipe_fp64_cycles_active_pct": counters.get("pipe_fp64_cycles_active_pct"),
                             "counters_source": counters.get("source")})
            else:
                fp64["counters"] = counters.get("error")
        except Exception as e:
            fp64 = {"error": repr(e)}
        line = {"metric": "cell-updates/s per SSPRK3 step", "value": value, "unit": "cell-updates/s", "n_gpus": world,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": dict(workload_config(world, dt), state_finite=finite),
                "clocks": clocks,
                "e2e": {"value": e2e_value, "unit": "cell-updates/s", "h2d_bytes_per_step": world * NVAR * field_bytes,
                        "d2h_bytes_per_step": world * NVAR * field_bytes, "bytes_are": "whole job (all ranks); per rank = /n_gpus",
                        "steps": args.e2e_steps, "api": "mw_dycore_time_step_host (pinned host buffers)"},
                "gpu_launches": int(launches),
                "parity": None if parity is None else {"ok": all(p.get("ok") is not False for p in parity) and any(p.get("ok") for p in parity),
                                                       "max_rel_err": max([p.get("max_rel_err", 0.0) for p in parity if p] or [None]),
                                                       "tol": 1e-9, "checks": parity},
                "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                             "traffic": traffic, "kernel": counters.get("kernel") or "k_stage_cell<1>", "kernel_ms": k_ms, "peak_source": peak_src,
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
