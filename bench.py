#!/usr/bin/env python
"""bench.py -- the reference's headline benchmark (Benchmark_dwf_fp32 shape) on the B200-native path.

  python bench.py --gpus N --steps K --warmup W [--impl reference] [--op Dhop|DhopEO] [--local L L L L] [--Ls 16]

A "step" is one full-lattice DomainWallFermionF::Dhop call (ref: benchmarks/Benchmark_dwf_fp32.cc:271-306):
fp32, 32^4 x Ls=16 per GPU (BASELINE.json configs[1]); random SU(3) links, random unit-norm source, mass 0.1, M5 1.8.
N > 1 (launched by torchrun, one rank per GPU) keeps the per-GPU volume fixed (weak scaling) and decomposes the
global lattice over z,t like the reference's --mpi 1.1.2.4: N=2 -> 1.1.1.2, N=4 -> 1.1.2.2, N=8 -> 1.1.2.4.

value      = whole-job GFlop/s (1320 flop per 5D site, ref Benchmark_dwf_fp32.cc:124), fields resident in HBM,
             CUDA events on the library's compute stream, max over ranks.
e2e        = same metric through the C ABI with HOST buffers (gb_op_dhop_host): every step moves the source from pinned
             host memory to the device, runs Dhop and moves the result back; the three are pipelined over t-slices so
             H2D and D2H overlap (PCIe full duplex) -- at N > 1 too (faces first, one halo exchange, slices streamed).
roofline   = algorithmic bytes (228 B per 5D site fp32) / measured kernel time vs MEASURED_PEAKS.json.
cpu_baseline = the reference's own CPU code (oracle/_ref/libgridref.so: unmodified paboyle/Grid compiled by
             oracle/Makefile.ref, AVX2 + OpenMP) timed on the host cores on a bounded sample; kind "reference".
             Falls back to the oracle port (kind "port") only where that library was not built.
--impl reference runs only that CPU leg (rank 0 only) and prints the same JSON line with "impl": "reference".
Beside the headline the same line carries (each measured in this run, each skippable with --no-...):
cg         = BASELINE configs[2]: even-odd Schur Moebius mixed-precision CG to 1e-8 at the headline volume, timed AFTER a warm-up
             solve, with ms per fp32 iteration (from a fixed 50-iteration fp32 CG) and its fraction of the 1872-B/site yardstick.
e2e_cg     = the call HMC makes: gauge field + source from pinned host memory in, solution back to the host, copies timed.
config4    = BASELINE configs[3]: Dhop (+ halo-exchange bandwidth + mixed CG) at the local volume 64.64.32.16 x Ls16 per GPU
             (global 64^4 on 8 GPUs as 1.1.2.4): the weak-scaling series whose 8-vs-1 ratio the north star asks for; its "strong" entry is
             the Dhop at FIXED global 64^4 x Ls16 on the N GPUs of the run (the strong-scaling series of the same config).
config5    = BASELINE configs[4]: improved staggered Dhop fp32 at 48^4 (global; split 1.1.2.4 on 8 GPUs).
At N > 1 the decomposed path is first compared with the CPU oracle on a small global lattice (parity_check): the hop per site (fp32, +-dag,
tolerance 1e-6) and a Schur conjugate-gradient solve (fp64: iteration count, true residual, solution) -- the driver-side parity of the
halo exchange and of the reductions summed over ranks -- and the host-buffer entry point that `e2e` times (pipelined on z / t splits: faces
first, one exchange, slices streamed); should that one disagree, `e2e` falls back to import + hop + export (parity_check.host_entry_form)
and e2e.parity_ok says whether the timed form matched.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

FLOPS_PER_SITE = 1320.0           # ref: benchmarks/Benchmark_dwf_fp32.cc:124
METRIC = "DWF Dhop fp32 GFlop/s (1320 flop/site), whole job"
UNIT = "GFlop/s"
MPI_FOR_N = {1: (1, 1, 1, 1), 2: (1, 1, 1, 2), 4: (1, 1, 2, 2), 8: (1, 1, 2, 4)}
PUBLISHED_A100_GFLOPS_PER_GPU = 2417.0  # Booster 4-node log, other hardware -- context only


def alg_bytes_per_site(Ls, w=4):
    """one spinor read + one written + the 8 double-stored links amortised over Ls."""
    return 2 * 24 * w + 8 * 18 * w / Ls


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def ncu_traffic(op):
    """dram bytes per launch of the dominant kernel from the committed ncu --set full capture (profiles/)."""
    p = os.path.join(ROOT, "profiles", "dhop_traffic.json")
    if os.path.exists(p):
        try:
            return json.load(open(p)).get(op)
        except Exception:
            pass
    return None


class ClockSampler:
    """nvidia-smi clocks / throttle reasons while the GPU is under load (recipe: B200_PROFILING.md)."""
    Q = "index,clocks.sm,clocks.max.sm,power.draw,utilization.gpu,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100", "-i", str(self.index)],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, smmax, reasons, power = [], [], set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                util = float(f[4])
                if util < 50:   # keep samples taken under load
                    continue
                sm.append(float(f[1])); smmax.append(float(f[2])); power.append(float(f[3]))
                for n, v in zip(names, f[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
            except ValueError:
                continue
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(smmax) if smmax else None,
                "power_w_max": max(power) if power else None, "samples_under_load": len(sm), "reasons": sorted(reasons)}


_CPU_SETUP = {}


class _StdoutToStderr:
    """The compiled reference logs on stdout (Grid : Message ...); this script's stdout carries ONE JSON line."""

    def __enter__(self):
        sys.stdout.flush()
        self.saved = os.dup(1)
        os.dup2(2, 1)

    def __exit__(self, *exc):
        sys.stdout.flush()
        os.dup2(self.saved, 1)
        os.close(self.saved)


def host_cores():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def use_all_host_cores(affinity=None):
    """torchrun exports OMP_NUM_THREADS=1 to its children and bench.py pins ranks to their GPU's NUMA node: the CPU legs undo
    both (they run on rank 0 alone while the other ranks wait)."""
    if affinity:
        try:
            os.sched_setaffinity(0, affinity)
        except Exception:
            pass
    n = host_cores()
    os.environ["OMP_NUM_THREADS"] = str(n)
    os.environ["GRIDREF_THREADS"] = str(n)
    return n


def cpu_leg(Ls, local, target_s=12.0, op="Dhop", fields=None):
    """Times the reference's CPU hopping term (fp32) on the host cores on a bounded sample of the SAME workload: the local volume
    of one rank, same operator, `ncall` calls (~target_s seconds).  kind "reference" = the unmodified reference compiled by
    oracle/Makefile.ref (DomainWallFermionF::Dhop, AVX2 SIMD, OpenMP on all host cores, comms none; the faster of its generic and
    hand-unrolled kernels); kind "port" = the oracle restatement, used only where oracle/_ref/libgridref.so is absent.
    fields = (U, x) host arrays exported from the device (same synthetic fields as the GPU run); drawn on the host if None."""
    import numpy as np
    from grid_b200 import synthetic as syn
    from oracle import pyoracle as po
    from oracle import pyref as pr
    dims = tuple(local)
    key = (Ls, dims, op)
    use_ref = pr.available()
    if key not in _CPU_SETUP:
        if fields is None:
            U = syn.hot_gauge(dims, seed=1, dtype=np.complex64)
            x = syn.random_fermion(dims, Ls, seed=2, dtype=np.complex64, normalise=True)
        else:
            U, x = fields
        which, vol = po.OP_DHOP, int(np.prod(dims)) * Ls
        if op == "DhopEO":
            x, which, vol = po.pick_checkerboard(dims, Ls, 1, x), po.OP_DHOP_EO, vol // 2
        if use_ref:
            with _StdoutToStderr():
                orc = pr.RefOp(1, dims, Ls, 0.1, 1.8, prec=0)
                orc.import_gauge(U)
                t_opt = {}
                for opt in (pr.OPT_GENERIC, pr.OPT_HAND_UNROLL):
                    pr.set_kernel_opt(opt)
                    orc.time_apply(which, x, 1)
                    t_opt[opt] = orc.time_apply(which, x, 2) / 2
                best = min(t_opt, key=t_opt.get)
                pr.set_kernel_opt(best)
            _CPU_SETUP[key] = (orc, x, which, vol, t_opt[best], "hand-unrolled" if best == pr.OPT_HAND_UNROLL else "generic")
        else:
            orc = po.OracleOp(1, dims, Ls, 0.1, 1.8, prec=0)
            orc.import_gauge(U)
            _CPU_SETUP[key] = (orc, x, which, vol, orc.time_apply(which, x, 1), "oracle")
    orc, x, which, vol, t1, variant = _CPU_SETUP[key]
    ncall = max(2, int(target_s / max(t1, 1e-3)))
    shape = "x".join(map(str, dims))
    if use_ref:
        with _StdoutToStderr():
            t = orc.time_apply(which, x, ncall)
        cores, kind = pr.num_threads(), "reference"
        what = f"the reference's DomainWallFermionF::{op} (paboyle/Grid CPU build: AVX2, OpenMP {cores} threads, comms none, {variant} kernel)"
    else:
        t = orc.time_apply(which, x, ncall)
        cores, kind = po.num_threads(), "port"
        what = f"the oracle {op} fp32 (port of the reference's generic kernel)"
    return {"value": FLOPS_PER_SITE * vol * ncall / t / 1e9, "unit": UNIT, "cores": cores, "kind": kind,
            "sample": f"{ncall} calls of {what} on {shape} x Ls{Ls} (the local volume of one GPU of the workload), {t:.1f} s",
            "seconds": t, "calls": ncall, "ms_per_call": 1e3 * t / ncall}


def run_reference(args, rank):
    if rank != 0:
        return
    use_all_host_cores()
    steps = max(1, args.steps)
    per_step_target = min(20.0, 120.0 / (steps + args.warmup))
    if os.environ.get("GB_BENCH_REF_SECONDS"):      # tests/test_bench_reference_arm.py shortens the sample; the driver never sets it
        per_step_target = float(os.environ["GB_BENCH_REF_SECONDS"])
    legs = [cpu_leg(args.Ls, args.local, target_s=per_step_target, op=args.op) for _ in range(args.warmup + steps)][args.warmup:]
    value = statistics.mean(l["value"] for l in legs)
    ms = statistics.mean(l["seconds"] for l in legs) * 1e3
    cb = dict(legs[-1]); cb["value"] = value
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": steps, "warmup": args.warmup,
            "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(args), "cpu_baseline": cb,
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "note": "the reference's own CPU implementation on the host cores, one process, all cores (see cpu_baseline.kind / sample); each step is a "
                    "bounded sample: calls at the per-GPU local volume of the workload (the CPU rate does not depend on the global extent)"}
    print(json.dumps(line), flush=True)


def workload_config(args):
    mpi = MPI_FOR_N[args.gpus]
    g = [l * m for l, m in zip(args.local, mpi)]
    return {"workload": f"DomainWallFermionF::{args.op} fp32, local {'x'.join(map(str, args.local))} x Ls{args.Ls} per GPU "
                        f"(BASELINE configs[1] shape), global {'x'.join(map(str, g))}, mpi {'.'.join(map(str, mpi))}",
            "op": args.op, "local_lattice": list(args.local), "global_lattice": g, "Ls": args.Ls, "mpi": list(mpi),
            "mass": 0.1, "M5": 1.8, "l2": "inputs (1.6 GB/field) larger than the 126 MB L2; no flush needed"}


def pin_to_gpu_numa(local_rank):
    """Pins this rank to the cores of its GPU's NUMA node (pinned staging buffers are then first-touched next to the GPU's PCIe
    root).  Returns (the affinity before pinning, a note)."""
    before = None
    try:
        before = os.sched_getaffinity(0)
        import pynvml
        pynvml.nvmlInit()
        bus = pynvml.nvmlDeviceGetPciInfo(pynvml.nvmlDeviceGetHandleByIndex(local_rank)).busId
        bus = bus.decode() if isinstance(bus, bytes) else bus
        bus = bus.lower()
        if len(bus.split(":")[0]) == 8:       # nvml prints an 8-digit domain, sysfs a 4-digit one
            bus = bus[4:]
        node = int(open(f"/sys/bus/pci/devices/{bus}/numa_node").read())
        if node < 0:
            return before, "numa node unknown"
        cpus = set()
        for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
            a, _, b = part.partition("-")
            cpus.update(range(int(a), int(b or a) + 1))
        cpus &= before
        if cpus:
            os.sched_setaffinity(0, cpus)
            return before, f"numa node {node}, {len(cpus)} cores"
        return before, f"numa node {node} has no allowed cores"
    except Exception as e:                     # no nvml / no sysfs: run unpinned
        return before, f"unpinned ({type(e).__name__})"


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=300)   # ref: ncall=300, Benchmark_dwf_fp32.cc:274
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--op", default="Dhop", choices=["Dhop", "DhopEO"])
    ap.add_argument("--local", type=int, nargs=4, default=[32, 32, 32, 32])
    ap.add_argument("--Ls", type=int, default=16)
    ap.add_argument("--e2e-steps", type=int, default=8)
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-overlap", action="store_true")
    ap.add_argument("--no-cg", action="store_true", help="skip the mixed-precision CG blocks (cg, e2e_cg)")
    ap.add_argument("--no-config4", action="store_true", help="skip the 64.64.32.16 x Ls16 per-GPU block (BASELINE configs[3])")
    ap.add_argument("--no-config5", action="store_true", help="skip the improved staggered 48^4 block (BASELINE configs[4])")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    assert args.gpus in MPI_FOR_N, "--gpus must be 1, 2, 4 or 8"
    if args.impl == "reference":
        run_reference(args, rank)
        return
    assert world == args.gpus, f"WORLD_SIZE={world} but --gpus {args.gpus}: launch with torch.distributed.run"

    import numpy as np
    import torch
    import grid_b200 as gb
    from grid_b200 import decomp

    affinity0, numa_note = pin_to_gpu_numa(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    ctx = gb.Context(local_rank)
    if world > 1:
        uid = [gb.Context.unique_id() if rank == 0 else None]
        dist.broadcast_object_list(uid, src=0)
        ctx.comm_init(rank, world, uid[0])
    else:
        ctx.comm_init(0, 1, None)
    mpi = MPI_FOR_N[args.gpus]
    nsplit = sum(1 for m in mpi if m > 1)
    peak, peak_src = measured_peak()

    def barrier():
        ctx.synchronize()
        if dist is not None:
            dist.barrier()
        ctx.synchronize()

    def max_over_ranks(x):
        if dist is None:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=f"cuda:{local_rank}")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def timed(step, steps, warmup):
        """W untimed calls, then exactly `steps` calls between barriers; CUDA events on the library's stream, max over ranks.
        Returns (ms per step, launches in the timed region on this rank)."""
        for _ in range(warmup):
            step()
        barrier()
        l0 = ctx.launch_count()
        ctx.timer_start()
        for _ in range(steps):
            step()
        ms = ctx.timer_stop()
        launches = ctx.launch_count() - l0
        barrier()
        return max_over_ranks(ms) / steps, launches

    def pinned(shape, dtype):
        return torch.empty(shape, dtype=dtype).pin_memory().numpy()

    # ---- N > 1: per-site parity of the decomposed hop (the form that is timed below) against the CPU oracle on a small GLOBAL
    #      lattice, before anything is timed
    parity_check = None
    if world > 1:
        from grid_b200 import synthetic as syn
        from oracle import pyoracle as po
        pl = [8, 8, 8, 8]
        pg = tuple(l * m for l, m in zip(pl, mpi))
        pLs = 16
        Ug = syn.hot_gauge(pg, seed=11)
        xg = syn.random_fermion(pg, pLs, seed=12)
        orc = po.OracleOp(1, pg, pLs, mass=0.1, M5=1.8, prec=1)
        orc.import_gauge(Ug)
        pgrid = gb.GridCartesian(ctx, pg, mpi)
        pD = gb.DomainWallFermion(gb.LatticeGaugeField(pgrid, gb.F32).import_lex(decomp.scatter(Ug, pg, mpi, rank)), pgrid, pLs, 0.1, 1.8)
        pin = gb.LatticeFermion(pgrid, pLs, gb.F32).import_lex(decomp.scatter(xg, pg, mpi, rank, inner=pLs).astype(np.complex64))
        pout = gb.LatticeFermion(pgrid, pLs, gb.F32)
        errs = {}
        for dag in (0, 1):
            pD.Dhop(pin, pout, dag)
            ref = decomp.scatter(orc.apply(po.OP_DHOP, xg, dag=dag), pg, mpi, rank, inner=pLs)
            a = pout.export_lex().reshape(ref.shape[0], -1).astype(np.complex128); r = ref.reshape(ref.shape[0], -1)
            errs[f"Dhop dag{dag}"] = max_over_ranks(float(np.max(np.linalg.norm(a - r, axis=1) / np.linalg.norm(r, axis=1))))
        # the host-buffer entry point that the e2e figure times (faces first, one exchange, slices streamed: dhop_host.cu); should it
        # ever disagree, the e2e leg falls back to import + hop + export (GB_HOST_PIPE_DECOMP=0) and says so
        host_form = "pipelined"
        hx = decomp.scatter(xg, pg, mpi, rank, inner=pLs).astype(np.complex64)
        refh = decomp.scatter(orc.apply(po.OP_DHOP, xg, dag=0), pg, mpi, rank, inner=pLs)
        for attempt in (0, 1):
            a = pD.Dhop_host(hx, np.empty_like(hx), 0).reshape(refh.shape[0], -1).astype(np.complex128); r = refh.reshape(refh.shape[0], -1)
            eh = max_over_ranks(float(np.max(np.linalg.norm(a - r, axis=1) / np.linalg.norm(r, axis=1))))
            if eh < 1e-6 or attempt == 1:
                break
            os.environ["GB_HOST_PIPE_DECOMP"] = "0"
            host_form = f"import + hop + export (the pipelined form differed from the oracle: {eh:.3e})"
        errs["Dhop_host dag0"] = eh
        # ... and of the decomposed Schur CG (hops with halos, s-space passes, reductions summed over ranks by ncclAllReduce in-stream):
        # iteration count, true residual and the solution itself against the oracle's ConjugateGradient on the global lattice
        cg_par, cg_ok = {}, True
        try:
            cLs = 8
            Mf = gb.MobiusFermion(gb.LatticeGaugeField(pgrid, gb.F64).import_lex(decomp.scatter(Ug, pg, mpi, rank)), pgrid, cLs, 0.1, 1.8, 1.5, 0.5)
            sg = syn.random_fermion(pg, cLs, seed=13)
            sfull = gb.LatticeFermion(pgrid, cLs, gb.F64).import_lex(decomp.scatter(sg, pg, mpi, rank, inner=cLs))
            so_ = gb.LatticeFermion(pgrid, cLs, gb.F64, gb.HALF)
            gb.pickCheckerboard(gb.Odd, so_, sfull)
            xo_ = gb.LatticeFermion(pgrid, cLs, gb.F64, gb.HALF).zero()
            pcg = gb.ConjugateGradient(1e-8, 10000)
            pcg(gb.SchurDiagMooeeOperator(Mf), so_, xo_)
            xfull = gb.LatticeFermion(pgrid, cLs, gb.F64).zero()
            gb.setCheckerboard(xfull, xo_)
            cg_par = {"iterations": pcg.IterationsToComplete, "true_residual": pcg.TrueResidual}
            ref_info = [None]
            if rank == 0:
                try:                                                        # whatever happens here, the broadcast below is reached
                    orc2 = po.OracleOp(1, pg, cLs, mass=0.1, M5=1.8, b=1.5, c=0.5, prec=1)
                    orc2.import_gauge(Ug)
                    x_ref, info = orc2.cg(1, po.pick_checkerboard(pg, cLs, 1, sg), 1e-8, 10000)
                    xr = np.zeros_like(sg)
                    po.set_checkerboard(pg, cLs, 1, xr, x_ref)
                    ref_info = [(info, xr)]
                    del orc2
                except Exception as e:
                    ref_info = [(f"{type(e).__name__}: {e}", None)]
            dist.broadcast_object_list(ref_info, src=0)
            info, xref_full = ref_info[0]
            if xref_full is None:
                raise RuntimeError(f"oracle side of the check failed on rank 0: {info}")
            mine = xfull.export_lex().reshape(-1); want = decomp.scatter(xref_full, pg, mpi, rank, inner=cLs).reshape(-1)
            num = max_over_ranks(float(np.linalg.norm(mine - want))); den = float(np.linalg.norm(xref_full.ravel()))
            cg_par.update(oracle_iterations=info["iterations"], oracle_true_residual=info["true_residual"], solution_rel_err=num / den * np.sqrt(world))
            cg_ok = abs(cg_par["iterations"] - info["iterations"]) <= max(1, 0.02 * info["iterations"]) and cg_par["solution_rel_err"] < 1e-6 and cg_par["true_residual"] < 1.5e-8
        except Exception as e:                                          # a fault of the check itself must not cost the bench line
            cg_par = {"error": f"{type(e).__name__}: {e}"}
        parity_check = {"against": "CPU oracle (fp64) on the global lattice", "global_lattice": list(pg), "Ls": pLs, "mpi": list(mpi),
                        "max_site_rel_err": errs, "tolerance": 1e-6, "host_entry_form": host_form,
                        "schur_cg": dict(cg_par, what="ConjugateGradient on SchurDiagMooeeOperator(MobiusFermion fp64, Ls 8) to 1e-8, decomposed over the ranks"),
                        "host_entry_ok": bool(errs["Dhop_host dag0"] < 1e-6),
                        # the line's `value` stands on the device-resident hop and the CG; a host-entry disagreement is reported (and voids
                        # `e2e`), it does not cost the whole line
                        "ok": all(e < 1e-6 for k_, e in errs.items() if not k_.startswith("Dhop_host")) and cg_ok}
        del pD, pin, pout, pgrid, orc, Ug, xg
        if not parity_check["ok"]:
            if rank == 0:
                print(json.dumps({"metric": METRIC, "error": "decomposed hop differs from the oracle", "parity_check": parity_check}), flush=True)
            sys.exit(1)

    # =========================================================== headline: BASELINE configs[1] per GPU
    gdims = [l * m for l, m in zip(args.local, mpi)]
    grid = gb.GridCartesian(ctx, gdims, mpi)
    Ls = args.Ls
    U = gb.LatticeGaugeField(grid, gb.F32).random(1)
    Dw = gb.DomainWallFermion(U, grid, Ls, 0.1, 1.8)
    if args.no_overlap:
        Dw.set_overlap(False)
    hop_form = "single rank" if world == 1 else ("serial comms" if args.no_overlap else "overlapped: one hop launch that sends its own t faces and acquires the halos in its surface CTAs (z faces by a pack kernel in front of it)")
    src = gb.LatticeFermion(grid, Ls, gb.F32).random(2)
    n2 = gb.norm2(src)
    gb.scale(src, 1.0 / np.sqrt(n2), src)     # ref: Benchmark_dwf_fp32.cc:175-176
    if args.op == "Dhop":
        fin, fout = src, gb.LatticeFermion(grid, Ls, gb.F32)
        step = lambda: Dw.Dhop(fin, fout, 0)
        sites_local = grid.lsites * Ls
    else:
        fin, fout = gb.LatticeFermion(grid, Ls, gb.F32, gb.HALF), gb.LatticeFermion(grid, Ls, gb.F32, gb.HALF)
        gb.pickCheckerboard(gb.Odd, fin, src)
        step = lambda: Dw.DhopEO(fin, fout, 0)
        sites_local = grid.lsites * Ls // 2
    sites_total = sites_local * world

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    ms_step, launches = timed(step, args.steps, args.warmup)
    value = FLOPS_PER_SITE * sites_total / (ms_step * 1e-3) / 1e9

    # ---- end to end through the C ABI with host buffers (pinned), copies inside the timed region
    host_in = pinned((fin.local_sites, 4, 3), torch.complex64)
    host_out = pinned((fin.local_sites, 4, 3), torch.complex64)
    host_in[...] = fin.export_lex()
    e2e_in = fin.like()

    def e2e_step():
        if args.op == "Dhop":
            # the host-buffer entry point of the C ABI: H2D / hop / D2H pipelined over t-slices
            Dw.Dhop_host(host_in, host_out, 0)
            return
        e2e_in.import_lex(host_in)
        e2e_in.set_checkerboard(gb.Odd)
        Dw.DhopEO(e2e_in, fout, 0)
        gb._chk(gb.lib().gb_fermion_export(fout.h, host_out.ctypes.data, gb.F32))
    e2e_step(); e2e_step()                      # untimed: staging buffers, first touch of the pinned host pages
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.e2e_steps):
        e2e_step()
    ctx.synchronize()
    e2e_s = max_over_ranks((time.perf_counter() - t0) / args.e2e_steps)
    barrier()
    e2e_value = FLOPS_PER_SITE * sites_total / e2e_s / 1e9
    # keep the device busy a little longer so the clock sampler sees the kernel under load
    t_end = time.time() + 1.5
    while time.time() < t_end:
        for _ in range(50):
            step()
        ctx.synchronize()
    clocks = sampler.stop() if rank == 0 else None
    barrier()
    # the same synthetic fields for the CPU leg (exported before they are released)
    cpu_fields = None
    if rank == 0 and not args.no_cpu and args.op == "Dhop":
        try:
            cpu_fields = (U.export_lex(np.complex64), host_in.copy())
        except Exception:
            cpu_fields = None
    del e2e_in, host_out

    # =========================================================== cg: BASELINE configs[2] at the headline volume
    def mixed_cg_block(grid_, U32, tag):
        """Even-odd Schur Moebius mixed-precision CG to 1e-8 (Test_dwf_mixedcg_prec shape): warm-up solve, timed solve, and a fixed
        50-iteration fp32 CG for the steady-state cost of one inner iteration."""
        vol_cb = grid_.lsites * Ls // 2
        Ud = gb.LatticeGaugeField(grid_, gb.F64).random(1)
        Dd = gb.MobiusFermion(Ud, grid_, Ls, 0.1, 1.8, 1.5, 0.5)
        Df = gb.MobiusFermion(U32, grid_, Ls, 0.1, 1.8, 1.5, 0.5)
        del Ud
        srcd = gb.LatticeFermion(grid_, Ls, gb.F64).random(2)
        so = gb.LatticeFermion(grid_, Ls, gb.F64, gb.HALF)
        gb.pickCheckerboard(gb.Odd, so, srcd)
        del srcd
        sol = gb.LatticeFermion(grid_, Ls, gb.F64, gb.HALF).zero()
        Lf, Ld = gb.SchurDiagMooeeOperator(Df), gb.SchurDiagMooeeOperator(Dd)
        warm = gb.MixedPrecisionConjugateGradient(1e-3, 10000, 50, Lf, Ld)
        warm(so, sol)                                                     # pools, s-space matrices, first-touch: not timed
        sol.zero()
        mcg = gb.MixedPrecisionConjugateGradient(1e-8, 10000, 50, Lf, Ld)
        barrier()
        t0 = time.perf_counter()
        mcg(so, sol)
        ctx.synchronize()
        cg_s = max_over_ranks(time.perf_counter() - t0)
        its = mcg.TotalInnerIterations + mcg.TotalFinalStepIterations
        # steady-state fp32 iteration: a ConjugateGradient on the fp32 operator stopped after 50 iterations
        sf, xf = gb.LatticeFermion(grid_, Ls, gb.F32, gb.HALF), gb.LatticeFermion(grid_, Ls, gb.F32, gb.HALF).zero()
        gb.precisionChange(sf, so)
        cg50 = gb.ConjugateGradient(1e-30, 50, err_on_no_conv=False)
        cg50(Lf, sf, xf)
        xf.zero()
        barrier()
        l0 = ctx.launch_count()
        ctx.timer_start()
        cg50(Lf, sf, xf)
        ms_it = max_over_ranks(ctx.timer_stop()) / max(cg50.IterationsToComplete, 1)
        l_it = (ctx.launch_count() - l0) / max(cg50.IterationsToComplete, 1)
        yard = 1872.0 * vol_cb                                            # ideal-fused bytes per Schur CG iteration and cb site
        out = {"solver": "MixedPrecisionConjugateGradient on SchurDiagMooeeOperator(MobiusFermion b=1.5 c=0.5, m=0.1, M5=1.8), tol 1e-8",
               "lattice": tag, "time_to_solution_s": cg_s, "timed_after_warmup_solve": True,
               "inner_iterations": mcg.TotalInnerIterations, "outer_iterations": mcg.TotalOuterIterations,
               "final_iterations": mcg.TotalFinalStepIterations, "true_residual": mcg.TrueResidual,
               "gflops": (1452.0 * 4 + 336.0) * vol_cb * world * its / cg_s / 1e9,
               "ms_per_iteration": ms_it, "launches_per_iteration": l_it,
               "roofline_frac": yard / (ms_it * 1e-3) / 1e9 / peak, "yardstick_bytes_per_cb_site": 1872.0}
        return out, (Dd, Df, so, sol)

    cg = e2e_cg = None
    if not args.no_cg:
        del fout
        cg, (Dd, Df, so, sol) = mixed_cg_block(grid, U, "headline volume")
        # ---- e2e_cg: what HMC calls -- gauge field and source arrive in host memory, the solution goes back to it
        try:
            hU = pinned((grid.lsites, 4, 3, 3), torch.complex128)
            hs = pinned((so.local_sites, 4, 3), torch.complex128)
            hx = pinned((so.local_sites, 4, 3), torch.complex128)
            Ud2 = gb.LatticeGaugeField(grid, gb.F64).random(1)
            hU[...] = Ud2.export_lex()
            hs[...] = so.export_lex()
            Uf2 = gb.LatticeGaugeField(grid, gb.F32)
            hU32 = pinned((grid.lsites, 4, 3, 3), torch.complex64)
            barrier()
            t0 = time.perf_counter()
            Ud2.import_lex(hU)
            hU32[...] = hU                                               # the fp32 copy of the links HMC's inner solver uses
            Uf2.import_lex(hU32)
            Dd.ImportGauge(Ud2); Df.ImportGauge(Uf2)
            so.import_lex(hs); so.set_checkerboard(gb.Odd)
            sol.zero()
            m2 = gb.MixedPrecisionConjugateGradient(1e-8, 10000, 50, gb.SchurDiagMooeeOperator(Df), gb.SchurDiagMooeeOperator(Dd))
            m2(so, sol)
            gb._chk(gb.lib().gb_fermion_export(sol.h, hx.ctypes.data, gb.F64))
            ctx.synchronize()
            e2e_cg_s = max_over_ranks(time.perf_counter() - t0)
            e2e_cg = {"time_to_solution_s": e2e_cg_s, "h2d_bytes": int(hU.nbytes + hU32.nbytes + hs.nbytes) * world, "d2h_bytes": int(hx.nbytes) * world,
                      "true_residual": m2.TrueResidual, "iterations": m2.TotalInnerIterations + m2.TotalFinalStepIterations,
                      "what": "LatticeGaugeField (fp64 + fp32) and odd-checkerboard source from pinned host memory, ImportGauge on both operators, "
                              "mixed CG to 1e-8, solution exported to host memory; wall clock, max over ranks"}
            del hU, hs, hx, hU32, Ud2, Uf2
        except Exception as e:                                          # never lose the headline line to a secondary block
            e2e_cg = {"error": f"{type(e).__name__}: {e}"}
        del Dd, Df, so, sol
    del Dw, src, fin, U, grid

    # =========================================================== config4: BASELINE configs[3], 64.64.32.16 x Ls16 per GPU
    config4 = None
    if not args.no_config4:
        try:
            l4 = [64, 64, 32, 16]
            g4 = [l * m for l, m in zip(l4, mpi)]
            grid4 = gb.GridCartesian(ctx, g4, mpi)
            U4 = gb.LatticeGaugeField(grid4, gb.F32).random(1)
            D4 = gb.DomainWallFermion(U4, grid4, 16, 0.1, 1.8)
            s4 = gb.LatticeFermion(grid4, 16, gb.F32).random(2)
            gb.scale(s4, 1.0 / np.sqrt(gb.norm2(s4)), s4)
            o4 = gb.LatticeFermion(grid4, 16, gb.F32)
            ms4, l4n = timed(lambda: D4.Dhop(s4, o4, 0), 100, 5)
            sites4 = grid4.lsites * 16
            config4 = {"workload": f"DomainWallFermionF::Dhop fp32, local 64x64x32x16 x Ls16 per GPU, global {'x'.join(map(str, g4))}, mpi {'.'.join(map(str, mpi))} "
                                   "(BASELINE configs[3]: global 64^4 x Ls16 at 8 GPUs; weak-scaling series, efficiency = per-GPU GFlop/s at N over N=1)",
                       "ms_per_step": ms4, "per_gpu_gflops": FLOPS_PER_SITE * sites4 / (ms4 * 1e-3) / 1e9, "gflops": FLOPS_PER_SITE * sites4 * world / (ms4 * 1e-3) / 1e9,
                       "roofline_frac": alg_bytes_per_site(16) * sites4 / (ms4 * 1e-3) / 1e9 / peak, "gpu_launches_per_step": l4n / 100}
            if world > 1:
                nbytes = 0
                for _ in range(5):
                    nbytes = D4.halo_exchange(s4)
                barrier()
                ctx.timer_start()
                for _ in range(50):
                    D4.halo_exchange(s4)
                msh = max_over_ranks(ctx.timer_stop()) / 50
                config4["halo"] = {"bytes_sent_per_gpu_per_hop": int(nbytes), "ms_exchange_alone": msh, "nvlink_GBs_out_per_gpu": nbytes / msh / 1e6,
                                   "nvlink_GBs_per_direction": nbytes / msh / 1e6 / (2 * nsplit),
                                   "what": "pack (project) + peer stores of every face of one Dhop + arrival, nothing else running; "
                                           "bytes / time, so this is a latency-inclusive lower bound of the link rate"}
            del D4, s4, o4
            if not args.no_cg:
                c4, keep = mixed_cg_block(grid4, U4, "local 64x64x32x16 x Ls16 per GPU")
                del keep
                config4["cg"] = c4
            del U4, grid4
            # ---- the strong-scaling series of the same config: global 64^4 x Ls16 on N GPUs (N = 8 is the line above)
            try:
                if world == 8:
                    config4["strong"] = {"global_lattice": [64, 64, 64, 64], "ms_per_step": config4["ms_per_step"], "gflops": config4["gflops"],
                                         "note": "identical to the weak-scaling line at 8 GPUs"}
                else:
                    gS = [64, 64, 64, 64]
                    gridS = gb.GridCartesian(ctx, gS, mpi)
                    US = gb.LatticeGaugeField(gridS, gb.F32).random(1)
                    DS = gb.DomainWallFermion(US, gridS, 16, 0.1, 1.8)
                    del US
                    sS = gb.LatticeFermion(gridS, 16, gb.F32).random(2)
                    oS = gb.LatticeFermion(gridS, 16, gb.F32)
                    msS, _ = timed(lambda: DS.Dhop(sS, oS, 0), 10, 3)
                    config4["strong"] = {"global_lattice": gS, "local_lattice": [g // m for g, m in zip(gS, mpi)], "ms_per_step": msS,
                                         "gflops": FLOPS_PER_SITE * 64 ** 4 * 16 / (msS * 1e-3) / 1e9,
                                         "roofline_frac": alg_bytes_per_site(16) * 64 ** 4 * 16 / world / (msS * 1e-3) / 1e9 / peak,
                                         "what": "DomainWallFermionF::Dhop fp32 at FIXED global 64^4 x Ls16 (strong scaling: ms at N over ms at 1)"}
                    del DS, sS, oS, gridS
            except Exception as e:
                config4["strong"] = {"error": f"{type(e).__name__}: {e}"}
        except Exception as e:
            config4 = {"error": f"{type(e).__name__}: {e}"}

    # =========================================================== config5: BASELINE configs[4], improved staggered 48^4 (global)
    config5 = None
    if not args.no_config5:
        try:
            g5 = [48, 48, 48, 48]
            grid5 = gb.GridCartesian(ctx, g5, mpi)
            U5 = gb.LatticeGaugeField(grid5, gb.F32).random(1)
            D5 = gb.ImprovedStaggeredFermion(U5, U5, grid5, 0.1)
            del U5
            s5, o5 = gb.LatticeStaggeredFermion(grid5, 1, gb.F32).random(2), gb.LatticeStaggeredFermion(grid5, 1, gb.F32)
            ms5, l5n = timed(lambda: D5.Dhop(s5, o5, 0), 100, 5)
            sites5 = 48 ** 4
            config5 = {"workload": f"ImprovedStaggeredFermionF::Dhop fp32 (Naik 3-hop), global 48^4, mpi {'.'.join(map(str, mpi))} (BASELINE configs[4]; strong scaling)",
                       "ms_per_step": ms5, "gflops": 1146.0 * sites5 / (ms5 * 1e-3) / 1e9,      # ref: benchmarks/Benchmark_staggered.cc:105
                       "roofline_frac": 1200.0 * sites5 / world / (ms5 * 1e-3) / 1e9 / peak, "algorithmic_bytes_per_site": 1200.0, "gpu_launches_per_step": l5n / 100}
            del D5, s5, o5, grid5
        except Exception as e:
            config5 = {"error": f"{type(e).__name__}: {e}"}

    # =========================================================== the reference's CPU code beside it (rank 0, every N)
    cpu = None
    if rank == 0 and not args.no_cpu:
        use_all_host_cores(affinity0)
        try:
            cpu = cpu_leg(Ls, args.local, op=args.op, fields=cpu_fields)
        except Exception as e:
            cpu = {"error": f"{type(e).__name__}: {e}"}
    if rank == 0:
        bps = alg_bytes_per_site(Ls)
        achieved = bps * sites_local / (ms_step * 1e-3) / 1e9       # per GPU
        kernel = "gb::dhop_col2_kernel<16,0,0>" if world == 1 else "gb::dhop_col2_kernel<16,0,1> (projects and sends its own t faces; + pack_send_kernel for the z faces where z is split)"
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
                "data": "synthetic", "config": workload_config(args), "hop_form": hop_form,
                "per_gpu_gflops": value / world, "vs_published_a100_per_gpu": value / world / PUBLISHED_A100_GFLOPS_PER_GPU,
                "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(host_in.nbytes) * world, "d2h_bytes_per_step": int(host_in.nbytes) * world,
                        "ms_per_step": e2e_s * 1e3, "steps": args.e2e_steps, "GBs_per_direction_per_gpu": host_in.nbytes / e2e_s / 1e9, "host_pinning": numa_note,
                        "parity_ok": True if parity_check is None else parity_check.get("host_entry_ok", True)},
                "gpu_launches": int(launches),
                "roofline": {"bound": "hbm", "kernel": kernel, "achieved": achieved, "peak": peak, "unit": "GB/s",
                             "frac": achieved / peak, "traffic": ncu_traffic(args.op) if world == 1 else None, "peak_source": peak_src,
                             "algorithmic_bytes_per_site": bps, "sites_per_launch": sites_local},
                "clocks": clocks, "cg": cg, "e2e_cg": e2e_cg, "config4": config4, "config5": config5, "parity_check": parity_check}
        if cpu is not None:
            line["cpu_baseline"] = cpu
        print(json.dumps(line), flush=True)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
