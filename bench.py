#!/usr/bin/env python
"""bench.py — LLG steps/s of the B200 hot path (base_projection + prepareElements + solve) on a
BASELINE.json configuration, with the SpMV roofline, an end-to-end leg through the public API with
host buffers, and the CPU baseline.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload NAME] [--impl reference]

One JSON line on stdout (rank 0).  For N > 1 launch with torch.distributed.run, one rank per GPU:
the unknowns are row-block (slab) partitioned over the ranks, same mesh at every N (strong
scaling).  `--impl reference` times the CPU implementation (oracle port driving the reference's
algorithm with all host threads) on a bounded sample of the same workload.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

METRIC = "LLG steps/s (assembly+BiCGStab)"
UNIT = "steps/s"


def log(*a):
    print(*a, file=sys.stderr, flush=True)


def measured_peak_gbs():
    """HBM copy bandwidth measured by the driver on this pool (MEASURED_PEAKS.json), else the
    fallback of /opt/skills/guides/B200_PROFILING.md."""
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            d = json.load(open(p))
            for k in ("hbm_gbs", "hbm_gbs_sustained"):
                if k in d and d[k]:
                    return float(d[k]), "measured (MEASURED_PEAKS.json %s)" % k
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md, 6.65 TB/s)"


class ClockSampler:
    """SM clock, power and throttle reasons sampled DURING the timed region: NVML polled every 10 ms
    from a thread (starts at once; the timed region of the default run is only ~0.2 s), nvidia-smi
    `-lms` as a fallback, and a one-shot query right after the region if neither produced a sample."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
    NAMES = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
    # nvmlClocksEventReason* bits
    BITS = dict(hw_slowdown=0x8, hw_thermal_slowdown=0x40, sw_thermal_slowdown=0x20, sw_power_cap=0x4)

    def __init__(self, device):
        self.device = device
        self.f = None
        self.p = None
        self.th = None
        self.stop_flag = False
        self.samples = []      # (sm_mhz, max_mhz, power_w, reasons bitmask)
        self.nvml = None

    def _poll(self):
        nv, h = self.nvml, self.h
        while not self.stop_flag:
            try:
                sm = nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)
                mx = nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM)
                pw = nv.nvmlDeviceGetPowerUsage(h) / 1000.0
                try:
                    rs = nv.nvmlDeviceGetCurrentClocksEventReasons(h)
                except Exception:
                    rs = nv.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                self.samples.append((float(sm), float(mx), float(pw), int(rs)))
            except Exception:
                pass
            time.sleep(0.01)

    def start(self):
        try:
            import pynvml as nv
            nv.nvmlInit()
            self.h = nv.nvmlDeviceGetHandleByIndex(self.device)
            self.nvml = nv
            import threading
            self.th = threading.Thread(target=self._poll, daemon=True)
            self.th.start()
            return
        except Exception:
            self.nvml = None
        try:
            self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
            self.p = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.device), "--query-gpu=" + self.Q,
                 "--format=csv,noheader,nounits", "-lms", "100"], stdout=self.f,
                stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def _parse_smi(self, text):
        sm, mx, power, reasons = [], [], [], set()
        for ln in text.splitlines():
            c = [x.strip() for x in ln.split(",")]
            if len(c) < 9:
                continue
            try:
                sm.append(float(c[1]))
                mx.append(float(c[2]))
                power.append(float(c[3]))
            except ValueError:
                continue
            for k, nm in enumerate(self.NAMES):
                if c[5 + k].lower().startswith("active"):
                    reasons.add(nm)
        return sm, mx, power, reasons

    def stop(self):
        out = dict(sm_mhz=None, sm_max_mhz=None, reasons=[])
        sm, mx, power, reasons, how = [], [], [], set(), None
        if self.th is not None:
            self.stop_flag = True
            self.th.join(timeout=2)
            for s_, m_, p_, r_ in self.samples:
                sm.append(s_)
                mx.append(m_)
                power.append(p_)
                for nm, bit in self.BITS.items():
                    if r_ & bit:
                        reasons.add(nm)
            how = "nvml 10 ms poll"
        elif self.p is not None:
            time.sleep(0.12)
            self.p.terminate()
            try:
                self.p.wait(timeout=5)
            except Exception:
                self.p.kill()
            self.f.flush()
            self.f.seek(0)
            sm, mx, power, reasons = self._parse_smi(self.f.read())
            how = "nvidia-smi -lms 100"
            try:
                os.unlink(self.f.name)
            except OSError:
                pass
        if not sm:
            try:  # nothing caught inside the region: one query right after it (GPU still warm)
                txt = subprocess.run(["nvidia-smi", "-i", str(self.device), "--query-gpu=" + self.Q,
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True,
                                     timeout=20).stdout
                sm, mx, power, reasons = self._parse_smi(txt)
                how = "nvidia-smi once, right after the timed region"
            except Exception:
                pass
        if sm:
            out = dict(sm_mhz=float(np.median(sm)), sm_max_mhz=float(max(mx)),
                       reasons=sorted(reasons), power_w=float(np.median(power)), samples=len(sm),
                       how=how)
        return out


# ---------------------------------------------------------------------------------------------
# CPU legs: the oracle port (reference algorithm, plain C + OpenMP) on the host cores
# ---------------------------------------------------------------------------------------------
FULL_NT = dict(ellipsoid=499, sp4=186000, disk1m=942000, tube5m=8 * 152 * 690 * 6, film20m=19969200,
               film20m_k=19969200)
SAMPLE_SCALE = dict(ellipsoid=1.0, sp4=1.0, disk1m=0.45, tube5m=0.2, film20m=0.25, film20m_k=0.25)


def _cpu_run(w, steps, warmup, threads, budget_s, min_steps=2):
    """W warm-up + K timed steps of base_projection + prepareElements + solve (+ evolution) of the CPU
    implementation on workload `w`.  `threads` > 1: OpenMP on the loops the reference runs under
    std::execution::par (basis, element integrals, matrix scatter, SpMV), BLAS-1 serial as in the
    reference; 1: everything serial (what the reference's pSTL does here without TBB).  Stops early once
    `budget_s` seconds of timed work are spent (never before `min_steps`); the counts really done are
    returned."""
    from oracle import fg_oracle_py as fo
    from feellgood_b200.linear_algebra import M_2_PI, mt19937_uniform01
    fo.build()
    pt = [fo.tet_prm(**r) for r in w.tet_regions]
    pf = [fo.tri_prm(**r) for r in w.tri_regions]
    oc = fo.OracleCtx(w.mesh, pt, pf, npi=w.npi, npi_tri=4 if w.npi == 5 else 1, tol=w.tol,
                      maxiter=w.maxiter)
    oc.set_num_threads(threads)
    oc.set_state(w.u)
    t = w.timing()

    def one(k):
        oc.base_projection(M_2_PI * mt19937_uniform01(1000 + k))
        oc.prepare_elements(w.Hext, t.get_dt(), t.prefactor)
        failed = oc.solve(t.get_dt())
        oc.evolution()
        return failed, oc.iter_info()["nit"]

    tw0 = time.perf_counter()
    wdone = 0
    for k in range(warmup):
        one(k)
        wdone += 1
        # a warm-up that would eat the whole budget is cut short (and reported)
        if time.perf_counter() - tw0 > 0.35 * budget_s:
            break
    its, t0, done = [], time.perf_counter(), 0
    for k in range(steps):
        f, nit = one(warmup + k)
        its.append(nit)
        done += 1
        if time.perf_counter() - t0 > budget_s and done >= min_steps:
            break
    el = time.perf_counter() - t0
    m = [float(x) for x in oc.avg(0, -1)]
    oc.close()
    return dict(steps=done, warmup=wdone, seconds=el, sps=done / el, mean_iters=float(np.mean(its)),
                avg_u=m)


def cpu_leg(workload_name, steps, warmup, budget_s=25.0, full=False):
    """The CPU baseline of BASELINE.md §3, both figures: `value` = threaded stand-in for TBB (all host
    cores), `serial` = as shipped in this container (1 core).  full=False: a reduced-extent sample of
    the workload (same cell size, materials, dt), scaled to the full workload by the tetrahedron count
    (work per step is linear in the mesh size at fixed cell size: same sparsity per row, same BiCGStab
    iteration count).  full=True: the threaded figure runs the WHOLE workload, no scaling."""
    from feellgood_b200 import workloads
    ncores = os.cpu_count() or 1
    NT_full = FULL_NT[workload_name]
    ws = workloads.build(workload_name, scale=SAMPLE_SCALE[workload_name])
    ratio = ws.mesh.NT / float(NT_full)
    if full:
        wf = workloads.build(workload_name) if ratio < 0.999 else ws
        r = _cpu_run(wf, steps, warmup, ncores, budget_s)
        value = r["sps"]
        sample = ("%s, the whole workload: %d tets / %d nodes, %d warm-up + %d timed steps in %.1f s; mean "
                  "%.1f BiCGStab iterations; oracle port with OpenMP (%d threads) on the reference's "
                  "parallel loops, serial BLAS-1 as in the reference"
                  % (wf.name, wf.mesh.NT, wf.mesh.NOD, r["warmup"], r["steps"], r["seconds"],
                     r["mean_iters"], ncores))
        del wf
    else:
        r = _cpu_run(ws, steps, warmup, ncores, budget_s)
        value = r["sps"] * ratio
        sample = ("%s: %d tets / %d nodes (%.4g of the %d-tet workload, same cell size and dt), %d "
                  "steps in %.1f s = %.3f steps/s on the sample, scaled by the tet ratio; mean %.1f "
                  "BiCGStab iterations; oracle port with OpenMP (%d threads) on the reference's parallel "
                  "loops, serial BLAS-1 as in the reference"
                  % (ws.name, ws.mesh.NT, ws.mesh.NOD, ratio, NT_full, r["steps"], r["seconds"], r["sps"],
                     r["mean_iters"], ncores))
    # as shipped here (TBB absent => serial pSTL backend): one core, always on the sample
    rs = _cpu_run(ws, 3, 1, 1, min(8.0, 0.4 * budget_s), min_steps=1)
    serial = dict(value=rs["sps"] * ratio, cores=1,
                  sample="%s (%.4g of the workload by tets), %d steps in %.1f s, scaled by the tet ratio"
                         % (ws.name, ratio, rs["steps"], rs["seconds"]))
    return dict(value=value, unit=UNIT, cores=ncores, kind="port", sample=sample,
                ms_per_step=1e3 / value, steps=r["steps"], warmup=r["warmup"], full_workload=bool(full),
                serial=serial)


def measure_traffic(args, kernel_regex, timeout_s=240):
    """dram__bytes_read.sum + dram__bytes_write.sum of one launch of the dominant kernel, from a fresh ncu
    capture of this very command (2 steps after the warm-up) in a child process.  None when ncu is not
    usable on this box or the capture does not finish in time."""
    import shutil
    ncu = shutil.which("ncu") or "/usr/local/cuda/bin/ncu"
    if not os.path.exists(ncu):
        return None
    out = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
    out.close()
    cmd = [ncu, "--metrics", "dram__bytes_read.sum,dram__bytes_write.sum", "--clock-control", "none",
           "-k", "regex:" + kernel_regex, "-s", "3", "-c", "1", "--csv", "--log-file", out.name,
           sys.executable, os.path.abspath(__file__), "--workload", args.workload, "--scale", str(args.scale),
           "--solver", args.solver, "--steps", "2", "--warmup", "3", "--no-e2e", "--no-cpu-baseline",
           "--traffic", "off", "--no-named-meshes"]
    try:
        subprocess.run(cmd, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL, timeout=timeout_s,
                       env=dict(os.environ, CUDA_VISIBLE_DEVICES=os.environ.get("CUDA_VISIBLE_DEVICES", "0")))
        tot, unit_mul = 0.0, dict(byte=1.0, Kbyte=1e3, Mbyte=1e6, Gbyte=1e9)
        import csv
        rows = [r for r in csv.reader(open(out.name)) if len(r) > 5]
        hdr = next(r for r in rows if "Metric Name" in r)
        iname, iunit, ival = hdr.index("Metric Name"), hdr.index("Metric Unit"), hdr.index("Metric Value")
        n = 0
        for r in rows:
            if r is hdr or len(r) <= ival or not r[iname].startswith("dram__bytes_"):
                continue
            tot += float(r[ival].replace(",", "")) * unit_mul.get(r[iunit], 1.0)
            n += 1
        return (tot, "ncu, this run (1 launch after 3 of the same kernel)") if n == 2 and tot > 0 else None
    except Exception:
        return None
    finally:
        try:
            os.unlink(out.name)
        except OSError:
            pass


# ---------------------------------------------------------------------------------------------
def main():
    # stdout carries exactly ONE JSON line: libraries that print to fd 1 (NCCL's version banner)
    # are sent to stderr for the whole run, and the line is written to the saved descriptor
    sys.stdout.flush()
    saved_stdout = os.dup(1)
    os.dup2(2, 1)

    def emit(line):
        os.write(saved_stdout, (json.dumps(line) + "\n").encode())

    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="film20m")
    ap.add_argument("--scale", type=float, default=1.0, help="shrink the workload (debug only)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--solver", default="persistent", choices=["persistent", "multi"],
                    help="persistent: one cooperative kernel per solve (default); multi: one kernel per phase")
    ap.add_argument("--partition", default="slab", choices=["slab", "rcb"],
                    help="N > 1: contiguous row blocks of the reference's node order (default) or a geometric "
                         "k-way partition (recursive coordinate bisection)")
    ap.add_argument("--traffic", default="ncu", choices=["ncu", "static", "off"],
                    help="roofline.traffic: ncu = measure DRAM bytes of the dominant kernel with a fresh ncu "
                         "capture in a child process (adds ~1 min); static = the committed capture in profiles/")
    ap.add_argument("--no-named-meshes", dest="named_meshes", action="store_false",
                    help="skip the short runs of BASELINE configs 1-4 that the default 1-GPU line appends")
    ap.add_argument("--kernel-times", action="store_true",
                    help="bracket every kernel with CUDA events and print the per-class table (stderr)")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.warmup < 3:
        args.warmup = 3

    if args.impl == "reference":
        if rank != 0:
            return 0
        # the WHOLE workload the product arm runs, W warm-up + K timed steps, cut only if the run would
        # exceed ~4 minutes (the counts really done are what the line reports)
        cb = cpu_leg(args.workload, args.steps, args.warmup, budget_s=float(os.environ.get("FG_REF_BUDGET_S", "120")),
                     full=True)
        line = dict(metric=METRIC, value=cb["value"], unit=UNIT, n_gpus=args.gpus, steps=cb["steps"],
                    warmup=cb["warmup"], ms_per_step=1e3 / cb["value"], higher_is_better=True,
                    scaling="strong", vs_baseline=None, dtype="f64", data="synthetic",
                    impl="reference", config=dict(workload=args.workload),
                    cpu_baseline=cb,
                    e2e=dict(value=cb["value"], unit=UNIT, h2d_bytes_per_step=0,
                             d2h_bytes_per_step=0))
        emit(line)
        return 0

    import torch
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (this framework has no CPU fallback; "
                         "use --impl reference for the CPU baseline)")
    torch.cuda.set_device(local_rank)
    try:  # run (and first-touch the pinned buffers) on the cores next to this rank's GPU: the end-to-end leg
        # moves 64 B per node and step over PCIe, and a buffer on the other socket halves that rate
        import pynvml as _nv
        _nv.nvmlInit()
        _nv.nvmlDeviceSetCpuAffinity(_nv.nvmlDeviceGetHandleByIndex(local_rank))
    except Exception:
        pass
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    from feellgood_b200 import LinAlgebra, workloads
    from feellgood_b200.linear_algebra import M_2_PI, mt19937_uniform01

    t_build = time.perf_counter()
    w = workloads.build(args.workload, scale=args.scale)
    if world > 1:
        from feellgood_b200.dist import DistLinAlgebra
        la = DistLinAlgebra(w.settings(), w.mesh, rank=rank, world=world, device=local_rank,
                            partition=args.partition)
    else:
        la = LinAlgebra(w.settings(), w.mesh, device=local_rank)
    la.set_state(w.u)
    la.set_solver(args.solver)
    tm = w.timing()
    if rank == 0:
        log("bench: %s NOD=%d NT=%d n=%d nnz=%d  (setup %.1f s)"
            % (w.name, w.mesh.NOD, w.mesh.NT, la.n, la.nnz, time.perf_counter() - t_build))
    stream = torch.cuda.ExternalStream(la.stream(), device=torch.device("cuda", local_rank))

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    iters = []
    # the basis angle of step k (same stream on every rank), drawn before the timed loops: input data
    n_ang = 2 * (args.warmup + args.steps) + 8
    angles = [M_2_PI * mt19937_uniform01(1000 + k) for k in range(n_ang)]

    def one_step(k):
        failed = la.step(w.Hext, tm, angle=angles[k])
        la.evolution()
        iters.append(la.iter["nit"])
        return failed

    def timed(fn, nsteps, k0):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        nfail = 0
        for k in range(nsteps):
            nfail += bool(fn(k0 + k))
        e1.record(stream)
        barrier()
        ms = e0.elapsed_time(e1)
        if dist is not None:
            tt = torch.tensor([ms], device="cuda", dtype=torch.float64)
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            ms = float(tt.item())
        return ms, nfail

    # ---- device-resident leg --------------------------------------------------------------
    for k in range(args.warmup):
        one_step(k)
    iters.clear()
    la.set_profiling(3 if args.kernel_times else 2)   # CUDA-event pair around every SpMV launch
    l0 = la.kernel_launches()
    clocks = ClockSampler(local_rank)
    if rank == 0:
        clocks.start()
    ms, nfail = timed(one_step, args.steps, args.warmup)
    clk = clocks.stop() if rank == 0 else {}
    launches = la.kernel_launches() - l0
    if args.kernel_times:
        kt = la.kernel_times()
        tot = sum(v[0] for v in kt.values())
        log("rank %d kernel classes over %d steps (%.3f ms/step timed, %.3f ms/step in kernels+gaps):"
            % (rank, args.steps, ms / args.steps, tot / args.steps))
        for k, (t, cnt) in kt.items():
            if cnt:
                log("  rank %d %-10s %5d launches  %8.3f ms/step  %8.1f us/launch"
                    % (rank, k, cnt, t / args.steps, 1e3 * t / cnt))
    solve_t = la.solve_times()
    if args.kernel_times and solve_t["kernel"][1]:
        for k, (t, cnt) in solve_t.items():
            if cnt:
                bd = la.solve_breakdown.get(k)
                log("  rank %d solve.%-10s %5d x  %8.3f ms/step  %8.1f us each%s"
                    % (rank, k, cnt, t / args.steps, 1e3 * t / cnt,
                       "   (work %.1f us, cross-GPU %.1f us)" % (1e3 * bd["work_ms"] / cnt, 1e3 * bd["cross_gpu_ms"] / cnt)
                       if bd else ""))
    spmv_ms, spmv_n = la.spmv_times()
    la.set_profiling(0)
    mean_it = float(np.mean(iters)) if iters else 0.0
    value = args.steps / (ms * 1e-3)
    # ---- parity block: the state after warm-up + timed steps, comparable across N (same mesh, same angle
    # stream, same dt): <m> over the magnetic volume (mesh::avg, collective on a partitioned mesh) and
    # the summed BiCGStab iterations; when profiles/parity_ref.json holds the 1-GPU value for this
    # workload and step count, the difference to it
    avg_u = [float(x) for x in la.avg("u")]
    parity = dict(after_steps=args.warmup + args.steps, avg_u=avg_u, sum_iters=int(np.sum(iters)),
                  failed_steps=int(nfail), solver=args.solver)
    try:
        pr = json.load(open(os.path.join(ROOT, "profiles", "parity_ref.json")))
        key = "%s:%d" % (w.name, args.warmup + args.steps)
        if key in pr:
            parity["ref_n1"] = pr[key]
            parity["max_abs_diff_vs_n1"] = float(np.max(np.abs(np.array(avg_u) - np.array(pr[key]["avg_u"]))))
            parity["sum_iters_diff_vs_n1"] = int(np.sum(iters)) - int(pr[key]["sum_iters"])
    except Exception:
        pass

    # ---- end-to-end leg: host buffers in and out every step -----------------------------------
    e2e = None
    if not args.no_e2e:
        NODl = la.NOD_local if hasattr(la, "NOD_local") else la.NOD
        pin = lambda *s: torch.empty(*s, dtype=torch.float64, pin_memory=True).numpy()
        h_phi, h_phiv = pin(NODl), pin(NODl)
        h_phi[:] = 0.0
        h_phiv[:] = 0.0
        h_u, h_v = pin(NODl, 3), pin(NODl, 3)

        def e2e_step(k):
            # what the reference's loop moves around one hot-path call: the demag solver's
            # potentials in (Nodes::set_phi/set_phiv), the new u and v out (its next input)
            la.set_potentials(h_phi, h_phiv)
            la.evolution()
            failed = la.step(w.Hext, tm, angle=angles[k])
            la.get_state_into(1, u=h_u, v=h_v)
            return failed

        for k in range(2):
            e2e_step(args.warmup + args.steps + k)
        ms_e, _ = timed(e2e_step, args.steps, args.warmup + args.steps + 2)
        e2e = dict(value=args.steps / (ms_e * 1e-3), unit=UNIT,
                   h2d_bytes_per_step=int(2 * 8 * w.mesh.NOD), d2h_bytes_per_step=int(48 * w.mesh.NOD),
                   note="per rank: potentials of its owned+ghost nodes in, u and v of its nodes out" if world > 1 else None,
                   ms_per_step=ms_e / args.steps,
                   api="LinAlgebra.set_potentials + evolution + step + get_state (pinned host)")

    # ---- roofline of the dominant kernel ---------------------------------------------------------
    # persistent solver: the solve kernel itself (setup product + all iterations + node update in one
    # launch, CUDA events around it), with the time its CTA 0 spent in each phase from in-kernel
    # %globaltimer stamps; multi-kernel solver: the SpMV launches
    peak, peak_src = measured_peak_gbs()
    n_loc = la.n_local if hasattr(la, "n_local") else la.n
    nnz_loc = la.nnz_local if hasattr(la, "nnz_local") else la.nnz
    col_bytes = getattr(la, "col_bytes", 4)
    b_spmv = workloads.spmv_bytes(n_loc, nnz_loc, col_bytes)
    roof = None
    if solve_t["kernel"][1] > 0:
        kms, kn = solve_t["kernel"]
        b_solve = workloads.solve_bytes(n_loc, nnz_loc, mean_it, col_bytes)
        ach = b_solve / (kms * 1e-3 / kn) / 1e9
        ph = {k: dict(us=1e3 * t / c, count=c) for k, (t, c) in solve_t.items() if c and k != "kernel"}
        sp = [solve_t[k] for k in ("B_spmv_v", "D_spmv_t") if solve_t[k][1]]
        sp_us = 1e3 * sum(t for t, _ in sp) / max(1, sum(c for _, c in sp))
        roof = dict(bound="hbm", kernel="k_llg_solve<BS,IDX16> (persistent: r = b - K x0, %.2f BiCGStab "
                    "iterations, node update)" % mean_it, achieved=ach, peak=peak, unit="GB/s",
                    frac=ach / peak, traffic=None, traffic_source=None, peak_source=peak_src,
                    bytes_per_launch=b_solve, col_index_bytes=col_bytes, launches=kn,
                    us_per_launch=1e3 * kms / kn, share_of_step=kms / ms, phases=ph,
                    spmv_phase=dict(us=sp_us, bytes=b_spmv, achieved=b_spmv / (sp_us * 1e-6) / 1e9 if sp_us else None,
                                    frac=b_spmv / (sp_us * 1e-6) / 1e9 / peak if sp_us else None,
                                    note="SpMV phase incl. its grid-wide reduction, timed inside the kernel",
                                    csr_equiv_gbs=workloads.spmv_bytes_csr(n_loc, nnz_loc) / (sp_us * 1e-6) / 1e9
                                    if sp_us else None))
    elif spmv_n > 0:
        ach = b_spmv / (spmv_ms * 1e-3 / spmv_n) / 1e9
        roof = dict(bound="hbm", kernel="k_spmv_node3<STAGE>", achieved=ach, peak=peak, unit="GB/s",
                    frac=ach / peak, traffic=None, traffic_source=None, peak_source=peak_src,
                    bytes_per_launch=b_spmv, col_index_bytes=col_bytes, launches=spmv_n, us_per_launch=1e3 * spmv_ms / spmv_n,
                    share_of_step=spmv_ms / ms,
                    csr_equiv_gbs=workloads.spmv_bytes_csr(n_loc, nnz_loc)
                    / (spmv_ms * 1e-3 / spmv_n) / 1e9)
    if roof is not None and world == 1 and rank == 0:
        tr = measure_traffic(args, roof["kernel"].split("<")[0]) if args.traffic == "ncu" else None
        if tr is not None:
            roof["traffic"], roof["traffic_source"] = tr
        else:
            tfile = os.path.join(ROOT, "profiles", "solve_traffic.json" if args.solver == "persistent" else "spmv_traffic.json")
            if os.path.exists(tfile):
                try:
                    tj = json.load(open(tfile))
                    if tj.get("workload") == w.name:
                        roof["traffic"] = tj.get("dram_bytes_per_launch")
                        roof["traffic_source"] = "static: %s (%s)" % (os.path.relpath(tfile, ROOT), tj.get("source", "ncu capture"))
                except Exception:
                    pass
    b_step = workloads.step_bytes(w.mesh.NOD, w.mesh.NT, la.n, la.nnz, mean_it, getattr(la, "col_bytes", 4))
    b_survey = workloads.step_bytes_survey(w.mesh.NOD, w.mesh.NT, la.n, la.nnz, mean_it)
    step_roof = dict(bytes_per_step=b_step, achieved=b_step / (ms * 1e-3 / args.steps) / 1e9 / world,
                     unit="GB/s per GPU", frac=b_step / (ms * 1e-3 / args.steps) / 1e9 / world / peak,
                     note="bytes this library's layout has to move (K never materialised)",
                     survey_bytes_per_step=b_survey,
                     survey_equiv_gbs=b_survey / (ms * 1e-3 / args.steps) / 1e9 / world)

    # ---- the other named meshes of BASELINE.json (configs 1-4), one short run each on this GPU: throughput
    # and whole-step roofline fraction, so that the line carries every named mesh (the headline stays config 5)
    named = None
    if rank == 0 and world == 1 and args.named_meshes and args.workload == "film20m" and args.scale == 1.0:
        named = {}
        for nm in ("tube5m", "disk1m", "sp4", "ellipsoid"):
            try:
                w2 = workloads.build(nm)
                la2 = LinAlgebra(w2.settings(), w2.mesh, device=local_rank)
                la2.set_state(w2.u)
                la2.set_solver(args.solver)
                tm2 = w2.timing()
                st2 = torch.cuda.ExternalStream(la2.stream(), device=torch.device("cuda", local_rank))
                it2 = []
                for k in range(5):
                    la2.step(w2.Hext, tm2, angle=angles[k])
                    la2.evolution()
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(st2)
                for k in range(40):
                    la2.step(w2.Hext, tm2, angle=angles[(5 + k) % n_ang])
                    la2.evolution()
                    it2.append(la2.iter["nit"])
                e1.record(st2)
                torch.cuda.synchronize()
                ms2 = e0.elapsed_time(e1) / 40
                b2 = workloads.step_bytes(w2.mesh.NOD, w2.mesh.NT, la2.n, la2.nnz, float(np.mean(it2)),
                                          getattr(la2, "col_bytes", 4))
                named[nm] = dict(config=w2.config, NT=w2.mesh.NT, NOD=w2.mesh.NOD, value=1e3 / ms2, unit=UNIT,
                                 ms_per_step=ms2, steps=40, warmup=5, mean_bicgstab_iters=float(np.mean(it2)),
                                 step_roofline_frac=b2 / (ms2 * 1e-3) / 1e9 / peak)
                la2.close()
                del w2, la2
            except Exception as e:
                named[nm] = dict(error=str(e))

    cb = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        try:
            cb = cpu_leg(args.workload, 400, 1, budget_s=12.0)   # ~12 + ~5 s of CPU work on the sample
        except Exception as e:  # the checker is optional for the measurement itself
            cb = dict(error=str(e))

    if rank == 0:
        line = dict(metric=METRIC, value=value, unit=UNIT, n_gpus=world, steps=args.steps,
                    warmup=args.warmup, ms_per_step=ms / args.steps, higher_is_better=True,
                    scaling="strong", vs_baseline=None, dtype="f64", data="synthetic",
                    config=dict(workload=w.name, baseline_config=w.config, NOD=w.mesh.NOD,
                                NT=w.mesh.NT, n=la.n, nnz=la.nnz, dt=w.dt, tol=w.tol, npi=w.npi,
                                mean_bicgstab_iters=mean_it, failed_steps=nfail,
                                l2="inputs exceed L2 (matrix %.0f MB >> 126 MB), no flush"
                                   % (8e-6 * la.nnz),
                                parallelism=("%s x%d" % ("row-block slabs" if args.partition == "slab" else "rcb boxes", world))
                                if world > 1 else "1 GPU"),
                    roofline=roof, step_roofline=step_roof, cpu_baseline=cb, e2e=e2e, parity=parity,
                    named_meshes=named,
                    gpu_launches=int(launches), clocks=clk)
        emit(line)
    la.close()
    if dist is not None:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
