#!/usr/bin/env python
"""bench.py -- MLUPS of the fused collide-stream sweep (BASELINE.json metric).

  python bench.py [--gpus N] [--steps K] [--warmup W]            our CUDA path
  python bench.py --impl reference [--gpus N] [--steps K] ...   the reference's CPU path on the host cores

Workload (config.workload): lid-driven cavity, D3Q19, fp64 BGK, tau 0.6, 512^3 interior cells per GPU
(BASELINE.json configs[1]; z-slab weak scaling for N > 1 = configs[4]).  A "step" is one
stream(); swap(); collide(); of the reference (src/main.cpp:50-52) = one launch of sweep_kernel.
MLUPS counts interior cell updates; the reference's own formula (ghost cells included,
src/main.cpp:64-65) is reported in config.mlups_reference_formula.

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
    """dram bytes per launch of sweep_kernel from the committed ncu capture, if one exists for this size."""
    path = os.path.join(ROOT, "profiles", "sweep_traffic.json")
    try:
        with open(path) as fh:
            d = json.load(fh)
        return d.get("D3Q%d_%d" % (Q, n))
    except Exception:
        return None


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
def cpu_reference_run(Q, n, steps, threads):
    """Times the reference's CPU path (oracle/_ref if built, else the oracle port) on an n^3 cavity."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import _oracle as O
    chk = O.ref()
    kind = "reference"
    if chk is None:
        chk, kind = O.oracle(), "port"
    out = chk.run(Q, n, n, n, TAU, O.cavity_boxes(n, n, n, LID), steps, threads=threads, want=())
    return n ** 3 * steps / out["seconds"] / 1e6, out["seconds"], kind


def cpu_baseline(Q, budget_s=15.0):
    threads = os.cpu_count() or 1
    n = 128
    mlups, sec, kind = cpu_reference_run(Q, n, 2, threads)          # calibrate
    steps = max(3, min(400, int(budget_s / max(sec / 2, 1e-4))))
    mlups, sec, kind = cpu_reference_run(Q, n, steps, threads)
    return {"value": round(mlups, 3), "unit": "MLUPS", "cores": threads, "kind": kind,
            "sample": "lid-driven cavity D3Q%d %d^3, %d steps, %.1f s, g++ -O2 -fopenmp, %d threads" % (Q, n, steps, sec, threads)}


def run_reference_arm(args, rank):
    if rank != 0:
        return
    Q, n = args.Q, 128
    threads = os.cpu_count() or 1
    # W warm-up steps, then exactly K timed steps of the reference's own loop (one Domain, like src/main.cpp:48-61;
    # the clock only runs inside stream(); swap(); collide();)
    if args.warmup > 0:
        cpu_reference_run(Q, n, args.warmup, threads)
    _, total_s, kind = cpu_reference_run(Q, n, args.steps, threads)
    mlups = n ** 3 * args.steps / total_s / 1e6
    line = {
        "impl": "reference", "metric": "MLUPS", "value": round(mlups, 3), "unit": "MLUPS", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(1e3 * total_s / args.steps, 4),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "lid-driven cavity D3Q%d BGK fp64 tau=0.6, CPU sample %d^3 of the %d^3-per-GPU workload; "
                               "each step = one stream();swap();collide() (time inside those calls only, as src/main.cpp:49-53)" % (Q, n, args.size)},
        "cpu_baseline": {"value": round(mlups, 3), "unit": "MLUPS", "cores": threads, "kind": kind,
                         "sample": "cavity D3Q%d %d^3, %d steps, g++ -O2 -fopenmp, %d threads" % (Q, n, args.steps, threads)},
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
    traffic = ncu_traffic(Q, n)
    roofline = {"bound": "hbm", "achieved": round(achieved, 1), "peak": peak, "unit": "GB/s",
                "frac": round(achieved / peak, 4), "traffic": traffic, "peak_source": peak_src,
                "kernel": "sweep_kernel<%d,%s>" % (Q, "exact" if args.exact else "fast"),
                "algorithmic_bytes_per_launch": bytes_per_cell * local_cells}

    # ---- e2e through the C ABI with HOST buffers: geometry H2D (kind + bc-id maps from pinned
    # memory), K steps, density/velocity D2H into pinned memory -- what a caller of
    # Domain::setBoundaryCondition ... write_vtk_file pays per output interval
    e2e = None
    if not args.no_e2e:
        e2e = measure_e2e(torch, dist, capi, run, args, world, cells_per_step)

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
            "config": {
                "workload": "lid-driven cavity D3Q%d BGK fp64 tau=0.6, %d^3 interior cells per GPU (%dx%dx%d global), "
                            "z-slabs, %s" % (Q, n, xl, yl, zl_global, "1 GPU" if world == 1 else "transport=" + transport),
                "arithmetic": "exact" if args.exact else "fast",
                "l2": "working set %.1f GB per GPU >> 126 MB L2, no flush needed" % (2 * Q * 8 * (n + 2) ** 3 / 1e9),
                "mlups_definition": "interior cell updates / s / 1e6",
                "mlups_reference_formula": round(mlups * ghost_cells / cells_per_step, 1),
            },
            "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": launches,
            "clocks": clocks.summary(),
        }
        print(json.dumps(line), flush=True)
    run.close()
    if world > 1:
        dist.destroy_process_group()
    return 0


def measure_e2e(torch, dist, capi, run, args, world, cells_per_step):
    """The same metric through the reference-facing calls with HOST buffers, wall clock, per rank:
       Domain::setBoundaryCondition...  -> lbm_b200_set_geometry(kind map, handler-id map, table)   [H2D, pinned]
       K x { stream(); swap(); collide(); } -> lbm_b200_step(K)
       io::write_vtk_file's read-out      -> lbm_b200_macroscopic(rho, u)                            [D2H, pinned]
    i.e. what one output interval of src/main.cpp costs when the scenario is (re)applied from the host."""
    import ctypes as C
    dom = run.dom
    n_int = dom.xl * dom.yl * dom.zl
    kind_np, bcid_np, table = capi.paint_boxes(dom.xl, dom.yl, dom.zl, dom.z_first,
                                               cavity_boxes(dom.xl, dom.yl, dom.zl_global))
    kind = torch.from_numpy(kind_np).pin_memory()
    bcid = torch.from_numpy(bcid_np.view("int16")).pin_memory()      # same bits; torch has no uint16 pinning on all builds
    _, tab = capi._boxes_arrays([(k, v, rho, (0,) * 6) for (k, v, rho) in table])
    rho = torch.empty(n_int, dtype=torch.float64).pin_memory()
    u = torch.empty(3 * n_int, dtype=torch.float64).pin_memory()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    capi._check(capi.lib.lbm_b200_set_geometry(dom._h, kind.data_ptr(), bcid.data_ptr(), C.cast(tab, C.c_void_p), len(table)))
    run.step(0)                       # commits the geometry (scatter + link mask) before the steps
    run.sync()
    t1 = time.perf_counter()
    run.step(args.steps)
    run.sync()
    t2 = time.perf_counter()
    capi._check(capi.lib.lbm_b200_macroscopic(dom._h, rho.data_ptr(), u.data_ptr()))
    torch.cuda.synchronize()
    t3 = time.perf_counter()
    if world > 1:
        dist.barrier()
    dt = time.perf_counter() - t0
    if world > 1:
        t = torch.tensor([dt], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dt = float(t.item())
    h2d = (kind.numel() + 2 * bcid.numel()) / args.steps
    d2h = (rho.numel() + u.numel()) * 8 / args.steps
    return {"value": round(cells_per_step * args.steps / dt / 1e6, 1), "unit": "MLUPS",
            "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
            "breakdown_ms": {"geometry_h2d": round(1e3 * (t1 - t0), 1), "steps": round(1e3 * (t2 - t1), 1),
                             "macroscopic_d2h": round(1e3 * (t3 - t2), 1)},
            "protocol": "per rank: kind + handler-id maps H2D from pinned memory (lbm_b200_set_geometry), %d steps, "
                        "density/velocity D2H into pinned memory (lbm_b200_macroscopic); wall clock, max over ranks; "
                        "bytes are totals divided by the %d steps" % (args.steps, args.steps)}


if __name__ == "__main__":
    sys.exit(main())
