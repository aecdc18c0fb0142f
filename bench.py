#!/usr/bin/env python
"""bench.py -- MLUPS of the fused collide-stream sweep (BASELINE.json metric).

  python bench.py [--gpus N] [--steps K] [--warmup W]            our CUDA path
  python bench.py --impl reference [--gpus N] [--steps K] ...   the reference's CPU path on the host cores

Workload (config.workload): lid-driven cavity, D3Q19, fp64 BGK, tau 0.6, 512^3 interior cells per GPU
(BASELINE.json configs[1]; z-slab weak scaling for N > 1 = configs[4]).  A "step" is one
stream(); swap(); collide(); of the reference (src/main.cpp:50-52) = one launch of sweep_kernel.
MLUPS counts interior cell updates; the reference's own formula (ghost cells included,
src/main.cpp:64-65) is reported in config.mlups_reference_formula.

e2e (the headline against the reference arm): the reference's main loop (src/main.cpp:46-61) through the C ABI
with HOST buffers -- scenario boxes from the host, E2E_INTERVALS output intervals of K steps each, after each
interval density/velocity of all interior cells copied to page-locked host memory (io/vtk.hpp:62-73); the copy
of interval i overlaps with the steps of interval i+1 (lbm_b200_macroscopic_begin/_end), the last one is exposed.

Prints ONE JSON line on rank 0.
"""
import argparse
import json
import os
import statistics
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

TAU = 0.6
LID = (0.05, 0.0, 0.0)
E2E_INTERVALS = 5              # output intervals inside the e2e region (each: K steps + one density/velocity read-out)
FALLBACK_HBM_GBS = 6650.0      # /opt/skills/guides/B200_PROFILING.md fallback


def cavity_boxes(xl, yl, zl):
    """build/scenarios/cavity.xml:6-11: z0 noslip, zmax moving wall, x0, xmax, y0, ymax noslip."""
    from lbm_b200 import capi
    z = (0.0, 0.0, 0.0)
    return [
        (capi.NOSLIP, z, 1.0, (0, xl + 1, 0, yl + 1, 0, 0)),
        (capi.MOVINGWALL, LID, 1.0, (0, xl + 1, 0, yl + 1, zl + 1, zl + 1)),
        (capi.NOSLIP, z, 1.0, (0, 0, 0, yl + 1, 0, zl + 1)),
        (capi.NOSLIP, z, 1.0, (xl + 1, xl + 1, 0, yl + 1, 0, zl + 1)),
        (capi.NOSLIP, z, 1.0, (0, xl + 1, 0, 0, 0, zl + 1)),
        (capi.NOSLIP, z, 1.0, (0, xl + 1, yl + 1, yl + 1, 0, zl + 1)),
    ]


def measured_peak():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as fh:
            return float(json.load(fh)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return FALLBACK_HBM_GBS, "fallback (B200_PROFILING.md)"


def ncu_traffic(Q, n):
    """dram bytes per launch of sweep_kernel from the committed ncu capture of this size, with its provenance
    (profiles/sweep_traffic.json is rewritten by tools/profile_gpu.sh from a capture of the current build)."""
    path = os.path.join(ROOT, "profiles", "sweep_traffic.json")
    try:
        with open(path) as fh:
            d = json.load(fh)
        return d.get("D3Q%d_%d" % (Q, n)), d.get("_source")
    except Exception:
        return None, None


def workload_config(Q, n, world):
    """config of BOTH arms (ours and --impl reference): the same dict when they run the same workload"""
    return {"workload": "lid-driven cavity D3Q%d BGK fp64 tau=0.6, %d^3 interior cells per GPU (%dx%dx%d global), %s"
                        % (Q, n, n, n, n * world, "1 GPU" if world == 1 else "z-slabs over %d GPUs" % world),
            "mlups_definition": "interior cell updates / s / 1e6",
            "l2": "working set %.1f GB per GPU >> 126 MB L2, no flush needed" % (2 * Q * 8 * (n + 2) ** 3 / 1e9)}


class ClockSampler:
    """Samples SM clock and throttle reasons with NVML while the timed region runs."""

    def __init__(self, index):
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self._thread = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def _loop(self):
        nv = self.nv
        names = {
            getattr(nv, "nvmlClocksThrottleReasonHwSlowdown", 0x8): "hw_slowdown",
            getattr(nv, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40): "hw_thermal_slowdown",
            getattr(nv, "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20): "sw_thermal_slowdown",
            getattr(nv, "nvmlClocksThrottleReasonSwPowerCap", 0x4): "sw_power_cap",
            getattr(nv, "nvmlClocksThrottleReasonHwPowerBrakeSlowdown", 0x80): "hw_power_brake",
        }
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in names.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            self._stop.wait(0.1)

    def __enter__(self):
        if self.nv is not None:
            self._thread = threading.Thread(target=self._loop, daemon=True)
            self._thread.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        if self._thread is not None:
            self._thread.join()

    def summary(self):
        return {"sm_mhz": statistics.median(self.samples) if self.samples else None,
                "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": len(self.samples)}


# --------------------------------------------------------------------------------------
def _checkers():
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import _oracle as O
    return O


def cpu_reference_run(Q, n, steps, threads, untimed=0, which="ref", zl=None):
    """Times the reference's CPU path on an n x n x zl cavity: `steps` iterations of stream(); swap(); collide(); of
    ONE Domain, the first `untimed` of them warm-up (the clock only runs inside the three calls, src/main.cpp:49-53).
    which: "ref" = oracle/_ref (the reference's own headers, g++ -O2 -fopenmp), falling back to the oracle port;
    "fast" = the same headers with -O3 -mavx2 -mfma."""
    O = _checkers()
    zl = zl or n
    chk, kind = (O.ref_fast(), "reference") if which == "fast" else (O.ref(), "reference")
    if chk is None:
        if which == "fast":
            return None
        chk, kind = O.oracle(), "port"
    out = chk.run(Q, n, n, zl, TAU, O.cavity_boxes(n, n, zl, LID), steps, threads=threads, untimed=untimed, want=())
    timed = steps - untimed
    return n * n * zl * timed / out["seconds"] / 1e6, out["seconds"], kind


def host_memory_gb():
    try:
        with open("/proc/meminfo") as fh:
            for line in fh:
                if line.startswith("MemAvailable:"):
                    return int(line.split()[1]) / 1e6
    except Exception:
        pass
    return 0.0


def cpu_baseline(Q, budget_s=12.0):
    """Bounded CPU sample on this box's host cores: the reference's default (1 thread, io/configuration.h:24), all
    cores with the parity flags, and all cores with the generous flags."""
    threads = os.cpu_count() or 1
    n = 128
    mlups, sec, kind = cpu_reference_run(Q, n, 3, threads, untimed=1)          # calibrate
    steps = max(3, min(400, int(budget_s / max(sec / 2, 1e-4))))
    mlups, sec, kind = cpu_reference_run(Q, n, steps + 1, threads, untimed=1)
    out = {"value": round(mlups, 3), "unit": "MLUPS", "cores": threads, "kind": kind,
           "sample": "lid-driven cavity D3Q%d %d^3, %d steps, %.1f s, g++ -O2 -fopenmp, %d threads" % (Q, n, steps, sec, threads)}
    one = cpu_reference_run(Q, n, 4, 1, untimed=1)
    out["one_thread"] = {"value": round(one[0], 3), "cores": 1, "sample": "same scenario, 3 steps, 1 thread (the reference's default omp-threads)"}
    fast = cpu_reference_run(Q, n, max(4, steps // 2), threads, untimed=1, which="fast")
    if fast is not None:
        out["generous_flags"] = {"value": round(fast[0], 3), "cores": threads,
                                 "sample": "same scenario, g++ -O3 -mavx2 -mfma -fopenmp (FMA contraction: not bit-identical)"}
    return out


def run_reference_arm(args, rank):
    """The reference's own CPU implementation on this box's host cores, all threads, on OUR arm's workload: the
    n^3 cavity of one GPU (K timed steps after W warm-up steps of one Domain).  The 512^3 lattice needs 43.5 GB of
    host memory and ~2 s per step on 16 cores; if memory or the time budget (REF_BUDGET_S) is short the lattice is
    halved until it fits and the line says so."""
    if rank != 0:
        return
    Q = args.Q
    threads = os.cpu_count() or 1
    budget = float(os.environ.get("REF_BUDGET_S", 420.0))
    cal, _, kind = cpu_reference_run(Q, 96, 3, threads, untimed=1)               # MLUPS estimate for the size choice
    n = args.size
    cell_bytes = 2 * (Q + 1) * 8                                                  # two lattices of {double pdf[Q]; Collision*}
    while n > 64:
        need_gb = (n + 2) ** 3 * cell_bytes / 1e9 * 1.15
        est_s = (args.steps + args.warmup + 3) * n ** 3 / (cal * 1e6)             # + construction / first touch
        if need_gb <= host_memory_gb() and est_s <= budget:
            break
        n //= 2
    mlups, total_s, kind = cpu_reference_run(Q, n, args.steps + args.warmup, threads, untimed=args.warmup)
    cfg = workload_config(Q, args.size, args.gpus)
    if n != args.size:
        cfg["cpu_arm"] = "timed on a %d^3 sample of the per-GPU lattice (host memory / time budget)" % n
    elif args.gpus > 1:
        cfg["cpu_arm"] = "timed on the %d^3 lattice of ONE GPU of the job (MLUPS is a per-cell rate)" % n
    line = {
        "impl": "reference", "metric": "MLUPS", "value": round(mlups, 3), "unit": "MLUPS", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(1e3 * total_s / args.steps, 4),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": cfg,
        "cpu_baseline": {"value": round(mlups, 3), "unit": "MLUPS", "cores": threads, "kind": kind,
                         "sample": "cavity D3Q%d %d^3, %d timed steps after %d warm-up steps of one Domain, g++ -O2 -fopenmp, "
                                   "%d threads; time inside stream();swap();collide() only (src/main.cpp:49-53)"
                                   % (Q, n, args.steps, args.warmup, threads)},
        "e2e": {"value": round(mlups, 3), "unit": "MLUPS", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--size", type=int, default=512, help="interior cells per edge, per GPU")
    ap.add_argument("--Q", type=int, default=19, choices=[15, 19, 27])
    ap.add_argument("--transport", default="p2p", choices=["nccl", "p2p"],
                    help="p2p (default): the sweep stores leaving populations into the neighbour GPU's ghost plane (CUDA IPC "
                         "over NVLink, device-side hand-shake), falls back to nccl if IPC is unavailable; nccl: split-phase "
                         "grouped send/recv of the halo planes overlapped with the interior sweep")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-exact", action="store_true", help="skip the EXACT-arithmetic timing")
    ap.add_argument("--no-parity", action="store_true", help="N > 1: skip the slab-parity check against the CPU checker")
    ap.add_argument("--exact", action="store_true", help="bit-exact arithmetic (for curiosity; not the bench mode)")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local_rank = int(os.environ.get("LOCAL_RANK", 0))

    if args.impl == "reference":
        run_reference_arm(args, rank)
        return 0
    if args.warmup < 3:
        args.warmup = 3

    if not os.path.exists(os.path.join(ROOT, "lbm_b200", "liblbm_b200.so")) and local_rank == 0:
        import __graft_entry__            # a fresh checkout: compile the CUDA library (nvcc, sm_100a) first
        __graft_entry__.build()
    import torch
    import torch.distributed as dist
    for _ in range(600):                  # other ranks wait for rank 0's build
        if os.path.exists(os.path.join(ROOT, "lbm_b200", "liblbm_b200.so")):
            break
        time.sleep(1.0)
    from lbm_b200 import capi
    from lbm_b200.slabs import SlabRunner

    if world != args.gpus and world > 1:
        raise SystemExit("--gpus %d but WORLD_SIZE=%d" % (args.gpus, world))
    if not torch.cuda.is_available() or capi.device_count() < 1:
        raise SystemExit("bench.py needs a CUDA device: the product has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    Q, n = args.Q, args.size
    xl = yl = n
    zl_global = n * world                       # weak scaling: n^3 per GPU, slabs along z
    boxes = cavity_boxes(xl, yl, zl_global)
    transport = args.transport
    try:
        run = SlabRunner(Q, xl, yl, zl_global, TAU, boxes, rank=rank, world=world, device=local_rank,
                         transport=transport, exact=args.exact)
        ok = 1
    except capi.LbmError as ex:          # e.g. CUDA IPC not permitted in this container
        sys.stderr.write("rank %d: transport %s unavailable (%s)\n" % (rank, transport, ex))
        ok = 0
    if world > 1:
        flag = torch.tensor([ok], device="cuda")
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        if not int(flag.item()):
            if transport == "nccl":
                raise SystemExit("no usable halo transport")
            transport = "nccl"           # another GPU transport, not a CPU path
            run = SlabRunner(Q, xl, yl, zl_global, TAU, boxes, rank=rank, world=world, device=local_rank,
                             transport=transport, exact=args.exact)
    elif not ok:
        raise SystemExit("cannot create the domain")
    dom = run.dom
    cells_per_step = xl * yl * zl_global        # interior cell updates per step, all ranks

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    if world > 1:
        # set-up, not warm-up: the first epochs of the peer hand-shake touch the freshly mapped IPC regions of both
        # neighbours; let every rank get through them before the W warm-up steps the contract asks for
        run.step(4)
        barrier()
    run.step(args.warmup)
    barrier()
    launches0 = dom.launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local_rank) as clocks:
        ev0.record(run.stream)
        run.step(args.steps)
        ev1.record(run.stream)
        barrier()
    ms = ev0.elapsed_time(ev1)
    launches = dom.launch_count() - launches0
    if world > 1:
        t = torch.tensor([ms], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
        lt = torch.tensor([launches], device="cuda", dtype=torch.int64)
        dist.all_reduce(lt, op=dist.ReduceOp.SUM)
        launches = int(lt.item())
    ms_per_step = ms / args.steps
    mlups = cells_per_step / (ms_per_step * 1e-3) / 1e6

    # ---- roofline of the dominant kernel (sweep_kernel): algorithmic bytes = 2*Q*8 per cell update,
    # one launch updates this rank's xl*yl*zl_local cells; duration = CUDA-event time / launches on
    # the launching stream (single-GPU: exactly one sweep launch per step)
    peak, peak_src = measured_peak()
    bytes_per_cell = 2 * Q * 8
    local_cells = xl * yl * run.zl
    local_ms = ev0.elapsed_time(ev1) / args.steps
    achieved = bytes_per_cell * local_cells / (local_ms * 1e-3) / 1e9
    traffic, traffic_src = ncu_traffic(Q, n)
    roofline = {"bound": "hbm", "achieved": round(achieved, 1), "peak": peak, "unit": "GB/s",
                "frac": round(achieved / peak, 4), "traffic": traffic, "traffic_source": traffic_src,
                "peak_source": peak_src, "frac_of_nominal_8TBs": round(achieved / 8000.0, 4),
                "kernel": "sweep_kernel<%d,%s>" % (Q, "exact" if args.exact else "fast"),
                "algorithmic_bytes_per_launch": bytes_per_cell * local_cells}

    # ---- the price of bit-exactness: the same workload in EXACT arithmetic (the reference's association, no FMA
    #      contraction, correctly rounded quotients from reciprocals -- DESIGN.md section 2)
    exact_mlups = None
    if not args.exact and not args.no_exact:
        dom.set_arithmetic(capi.EXACT)
        run.step(3)
        barrier()
        ea, eb = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ea.record(run.stream)
        run.step(10)
        eb.record(run.stream)
        barrier()
        t = torch.tensor([ea.elapsed_time(eb)], device="cuda", dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        exact_mlups = round(cells_per_step * 10 / (float(t.item()) * 1e-3) / 1e6, 1)
        dom.set_arithmetic(capi.FAST)

    # ---- e2e through the C ABI with HOST buffers (see the module docstring)
    e2e = None
    if not args.no_e2e:
        e2e = measure_e2e(torch, dist, capi, run, args, world, cells_per_step)

    # ---- N > 1: the split must not change a single bit (SURVEY 8d config 5 proxy), checked in this very run
    parity = None
    if world > 1 and not args.no_parity:
        parity = slab_parity(torch, dist, capi, SlabRunner, args, rank, world, local_rank, transport)

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        try:
            cpu = cpu_baseline(Q)
        except Exception as ex:       # the checker is optional equipment, never the product
            cpu = {"value": None, "unit": "MLUPS", "cores": os.cpu_count(), "kind": "port", "sample": "failed: %s" % ex}

    if rank == 0:
        ghost_cells = (xl + 2) * (yl + 2) * (zl_global + 2)
        line = {
            "metric": "MLUPS", "value": round(mlups, 1), "unit": "MLUPS", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": round(ms_per_step, 4), "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": workload_config(Q, n, world),
            "arithmetic": "exact" if args.exact else "fast", "exact_mlups": exact_mlups,
            "transport": None if world == 1 else transport,
            "mlups_reference_formula": round(mlups * ghost_cells / cells_per_step, 1),
            "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": launches,
            "parity": parity, "clocks": clocks.summary(),
        }
        print(json.dumps(line), flush=True)
    run.close()
    if world > 1:
        dist.destroy_process_group()
    if parity is not None and not parity["bitwise"]:
        return 3
    return 0


def slab_parity(torch, dist, capi, SlabRunner, args, rank, world, local_rank, transport):
    """64 x 64 x (64*N) cavity, EXACT arithmetic, 40 steps on the same runner / transport as the timed run; every
    rank's populations, density and velocity against the CPU checker (rank 0 runs it), bit for bit."""
    import numpy as np
    Q, n, steps = args.Q, 64, 40
    zl = n * world
    boxes = cavity_boxes(n, n, zl)
    run = SlabRunner(Q, n, n, zl, TAU, boxes, rank=rank, world=world, device=local_rank, transport=transport, exact=True)
    run.step(steps)
    run.prepare_readback()
    f = run.dom.download().reshape(run.zl + 2, -1, Q)[1:-1]           # own planes
    rho, u = run.dom.macroscopic()
    ok = True
    want = [None]
    if rank == 0:
        O = _checkers()
        chk = O.ref() or O.oracle()
        ref = chk.run(Q, n, n, zl, TAU, O.cavity_boxes(n, n, zl, LID), steps)
        plane = (n + 2) * (n + 2)
        want = [(ref["f"].reshape(zl + 2, plane, Q), ref["rho"], ref["u"], chk.prefix)]
    dist.broadcast_object_list(want, src=0)
    wf, wrho, wu, who = want[0]
    z0 = run.z_first
    ok = (np.array_equal(f, wf[z0:z0 + run.zl]) and np.array_equal(rho, wrho[z0 - 1:z0 - 1 + run.zl])
          and np.array_equal(u, wu[z0 - 1:z0 - 1 + run.zl]))
    flag = torch.tensor([1 if ok else 0], device="cuda")
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    run.close()
    return {"ranks": world, "bitwise": bool(int(flag.item())), "checker": who,
            "case": "cavity D3Q%d 64x64x%d, EXACT arithmetic, %d steps, transport=%s: populations of every cell, density, "
                    "velocity of all ranks == CPU run" % (Q, zl, steps, transport)}


def measure_e2e(torch, dist, capi, run, args, world, cells_per_step):
    """The same metric through the reference-facing calls with HOST buffers, wall clock, per rank:
       Domain::setBoundaryCondition x 6 (io/scenario.h:91-128) -> lbm_b200_set_boxes(box list)           [H2D]
       E2E_INTERVALS x { K x { stream(); swap(); collide(); }  -> lbm_b200_step(K)
                         io::write_vtk_file's read-out loop     -> lbm_b200_macroscopic_begin / _end }   [D2H, pinned]
    i.e. src/main.cpp:46-61 with timesteps-per-plot = K.  The read-out of interval i crosses PCIe while interval
    i+1 is computed; the last one has nothing to hide behind."""
    dom = run.dom
    n_int = dom.xl * dom.yl * dom.zl
    boxes = cavity_boxes(dom.xl, dom.yl, dom.zl_global)
    ext, tab = capi._boxes_arrays(boxes)
    capi.lib.lbm_b200_bind_host_thread(run.device)                   # this rank's thread next to its GPU
    rho = capi.HostBuffer(n_int, device=run.device)                  # page-locked, on the GPU's NUMA node
    u = capi.HostBuffer(3 * n_int, device=run.device)
    K, M = args.steps, E2E_INTERVALS
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    import ctypes as C
    t0 = time.perf_counter()
    capi._check(capi.lib.lbm_b200_set_boxes(dom._h, ext.ctypes.data, C.cast(tab, C.c_void_p), len(boxes)))
    run.step(0)                       # commits the geometry (link mask) before the steps
    run.sync()
    t1 = time.perf_counter()
    for i in range(M):
        run.step(K)
        if i > 0:
            dom.macroscopic_end()     # output i-1 must be on the host before its buffers are reused
        dom.macroscopic_begin(rho.address, u.address)
    run.sync()                        # steps and the last reduction are done ...
    t_c = time.perf_counter()
    dom.macroscopic_end()             # ... what remains is the last read-out crossing PCIe, un-overlapped
    t2 = time.perf_counter()
    if world > 1:
        dist.barrier()
    dt = time.perf_counter() - t0
    checksum = float(rho.array[:: max(1, n_int // 4096)].sum())        # the host really holds the densities
    if world > 1:
        t = torch.tensor([dt], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dt = float(t.item())
    total_steps = K * M
    d2h = (rho.array.size + u.array.size) * 8 * M / total_steps
    h2d = (ext.nbytes + C.sizeof(tab)) / total_steps
    out = {"value": round(cells_per_step * total_steps / dt / 1e6, 1), "unit": "MLUPS",
           "h2d_bytes_per_step": int(round(h2d)), "d2h_bytes_per_step": int(d2h),
           "intervals": M, "steps_per_interval": K,
           "breakdown_ms": {"geometry": round(1e3 * (t1 - t0), 2), "intervals_overlapped": round(1e3 * (t_c - t1), 1),
                            "last_readout_exposed": round(1e3 * (t2 - t_c), 1)},
           "single_interval_mlups_estimate": round(cells_per_step * K / ((t1 - t0) + (t_c - t1) / M + (t2 - t_c)) / 1e6, 1),
           "density_checksum": checksum,
           "protocol": "per rank: scenario boxes from the host (lbm_b200_set_boxes), then %d output intervals of %d steps "
                       "(src/main.cpp:46-61 with timesteps-per-plot = %d); after each interval density + velocity of all "
                       "interior cells go to page-locked host memory (lbm_b200_macroscopic_begin/_end); the read-out of "
                       "interval i overlaps with the steps of interval i+1, the last one is exposed; wall clock, max over "
                       "ranks; single_interval_mlups_estimate = one interval with nothing to overlap" % (M, K, K)}
    rho.free()
    u.free()
    return out


if __name__ == "__main__":
    sys.exit(main())
