#!/usr/bin/env python
"""bench.py -- collocation defect + FD-Jacobian evals/sec (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--batch B] [--workload NAME]
    python bench.py --impl reference ...        # the CPU path (oracle port) on the host cores

One *step* = one pass of the hot path over one batch of synthetic instances: for every
instance the stacked vector c = [c_eq; c_ineq; cost] and its dense forward-difference
Jacobian (K1 D.X tensor-core GEMM + K2 fused sweep, the two launches ogb_eval_fd enqueues).
One *eval* = one instance of one step.

Default workload: the 4096-instance Goddard 50-node batch (north_star's headline; the
batch=1024 case of BASELINE.json configs[1] is `--batch 1024`).  Weak scaling: every
rank evaluates its own B instances (shards of one seeded batch of N*B), no data-path
collective; rank 0 prints ONE JSON line.

Timing: W >= 3 warm-up steps, then exactly K timed steps, each bracketed by CUDA events
on the launching stream; a 256 MiB memset flushes L2 between steps outside the event
pairs (each step also writes ~3 GB, 24x L2).  Time = sum of the K event durations, MAX
over ranks.  `e2e` repeats the measurement through the host-buffer C-ABI call
(ogb_host_eval_fd): pinned host p -> H2D -> K1 -> K2a (sweep kernel, packed output) -> D2H of c and the
packed non-zeros -> host threads write the dense J into host memory, all inside the timed region (host
clock).  Also in the line (set-up or after the timed region, never inside it): `parity` (3 instances
against the numpy oracle), `tensor` (K1 against a measured FP64 DGEMM peak), `e2e_sqp` (the transport the
SQP driver uses: packed values scattered into per-instance SLSQP buffers), `extra` (single-instance
latency, the sparse / exact evaluation rates, a batched multi-start solve next to the CPU reference).
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
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "collocation defect+FD-Jacobian evals/sec"
UNIT = "evals/s"
WORKLOADS = {                      # name -> (workloads.CONFIGS key, default batch)
    "goddard50": ("cfg2_goddard50", 4096),
    "goddard_knot30x2": ("cfg3_goddard_knot30x2", 4096),
    "polar3x40": ("cfg4_polar3x40", 512),
    "lowthrust128": ("cfg5_lowthrust128", 1024),
    "brachistochrone20": ("cfg1_brachistochrone20", 4096),
}


# ------------------------------------------------------------------ CPU baseline (reference / oracle port)
_CPU_STATE = {}


def cpu_kind():
    """"reference": the unmodified OpenGoddard module ($OPENGODDARD_REF, /root/reference or the copy
    pip-installed into baseline/_ref) + SciPy's own approx_derivative; "port": the numpy oracle."""
    if os.environ.get("OGB200_CPU_KIND") in ("reference", "port"):
        return os.environ["OGB200_CPU_KIND"]
    from oracle import ref_loader
    return "reference" if ref_loader.reference_available() else "port"


def _cpu_setup(cfg, kind):
    key = (cfg, kind)
    if key in _CPU_STATE:
        return _CPU_STATE[key]
    import numpy as np
    from opengoddard_b200 import workloads
    if kind == "reference":
        from scipy.optimize._numdiff import approx_derivative
        from oracle import ref_loader
        mod = ref_loader.load_reference()
        wl = workloads.build(cfg, mod)
        cap = ref_loader.capture_solve(mod, wl.prob, wl.obj)      # the closures Problem.solve hands to SciPy
        lb = np.array([-np.inf if b[0] is None else float(b[0]) for b in cap.bounds])
        ub = np.array([np.inf if b[1] is None else float(b[1]) for b in cap.bounds])
        h = float(np.sqrt(np.finfo(np.float64).eps))
        funs = [(cap.constraints[0]["fun"], cap.constraints[0]["args"]), (cap.constraints[1]["fun"], cap.constraints[1]["args"])]
        if cap.jac is None:
            funs.append((cap.fun, cap.args))                      # no cost_derivative: SciPy differences the cost too

        def one(p):                                               # what SLSQP asks for in mode -1 (_slsqp_py.py:533-534)
            x = np.clip(p, lb, ub)
            for f, a in funs:
                approx_derivative(f, x, method="2-point", abs_step=h, args=a, bounds=(lb, ub))
    else:
        from oracle import og_numpy
        wl = workloads.build(cfg, og_numpy)
        lb, ub = og_numpy.bounds_arrays(wl.prob)

        def one(p):
            og_numpy.eval_fd(wl.prob, wl.obj, p, lb, ub)
    _CPU_STATE[key] = (wl, one)
    return _CPU_STATE[key]


def _cpu_worker(args):
    """Evaluate a slice of the sample on one host core; returns (n_done, seconds)."""
    cfg, first, count, kind = args
    os.environ["OMP_NUM_THREADS"] = "1"
    import contextlib
    import io
    from opengoddard_b200 import workloads
    with contextlib.redirect_stdout(io.StringIO()):
        wl, one = _cpu_setup(cfg, kind)
    P = workloads.make_batch(wl, count, first=first)
    t0 = time.perf_counter()
    for p in P:
        one(p)
    return count, time.perf_counter() - t0


class CpuPool:
    """Host worker processes (spawned once) that run the CPU path on seeded instances."""

    def __init__(self, cfg, procs, kind=None):
        import multiprocessing as mp
        self.cfg, self.procs, self.kind = cfg, procs, kind or cpu_kind()
        self.pool = None
        if procs > 1:
            self.pool = mp.get_context("spawn").Pool(procs)
            self.pool.map(_cpu_worker, [(cfg, 0, 1, self.kind)] * procs, chunksize=1)   # warm the workers (imports)
        else:
            _cpu_worker((cfg, 0, 1, self.kind))

    def run(self, sample, first=0):
        """-> (evals/s over the wall clock, wall seconds, summed CPU seconds)"""
        per = [sample // self.procs + (1 if r < sample % self.procs else 0) for r in range(self.procs)]
        jobs = []
        for k in per:
            if k:
                jobs.append((self.cfg, first, k, self.kind))
                first += k
        t0 = time.perf_counter()
        res = self.pool.map(_cpu_worker, jobs, chunksize=1) if self.pool else [_cpu_worker(j) for j in jobs]
        wall = time.perf_counter() - t0
        return sum(r[0] for r in res) / wall, wall, sum(r[1] for r in res)

    def close(self):
        if self.pool:
            self.pool.close()
            self.pool.join()


def cpu_info():
    model = "unknown"
    try:
        for line in open("/proc/cpuinfo"):
            if line.startswith("model name"):
                model = line.split(":", 1)[1].strip()
                break
    except OSError:
        pass
    import numpy
    import scipy
    return {"cpu_model": model, "os_cpu_count": os.cpu_count(), "numpy": numpy.__version__,
            "scipy": scipy.__version__, "omp_num_threads_per_worker": 1}


def host_procs():
    try:
        n = len(os.sched_getaffinity(0))
    except AttributeError:
        n = os.cpu_count() or 1
    return max(1, min(n, 64))


# ------------------------------------------------------------------ clocks
class ClockSampler:
    """SM clock and throttle reasons sampled DURING the timed region (NVML, every ~2 ms; falls
    back to one `nvidia-smi` query if the NVML binding is unavailable)."""
    BITS = {"hw_slowdown": 0x8, "sw_power_cap": 0x4, "sw_thermal_slowdown": 0x20,
            "hw_thermal_slowdown": 0x40, "hw_power_brake_slowdown": 0x80}

    def __init__(self, index):
        self.index, self.sm, self.reasons, self.smax = index, [], set(), None
        self._stop = threading.Event()
        self.thread = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.smax = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
        except Exception:
            self.nv = None

    def _sample(self):
        nv = self.nv
        self.sm.append(float(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)))
        try:
            mask = int(nv.nvmlDeviceGetCurrentClocksEventReasons(self.h))
        except Exception:
            mask = int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h))
        for name, bit in self.BITS.items():
            if mask & bit:
                self.reasons.add(name)

    def _loop(self):
        while not self._stop.is_set():
            try:
                self._sample()
            except Exception:
                break
            time.sleep(0.002)

    def start(self):
        if self.nv is not None:
            self.thread = threading.Thread(target=self._loop, daemon=True)
            self.thread.start()

    def stop(self):
        if self.nv is None:
            try:
                q = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=clocks.sm,clocks.max.sm",
                                    "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=20)
                a, b = [float(x) for x in q.stdout.strip().split(",")]
                return {"sm_mhz": a, "sm_max_mhz": b, "reasons": [], "samples": 1,
                        "note": "single nvidia-smi query after the timed region (no NVML binding)"}
            except Exception:
                return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["clock query unavailable"], "samples": 0}
        self._stop.set()
        if self.thread is not None:
            self.thread.join(timeout=2)
        if not self.sm:
            try:
                self._sample()
            except Exception:
                pass
        return {"sm_mhz": statistics.median(self.sm) if self.sm else None, "sm_max_mhz": self.smax,
                "reasons": sorted(self.reasons), "samples": len(self.sm)}


# ------------------------------------------------------------------ reference arm
CPU_WHAT = {"reference": "the unmodified reference module (its Problem.solve closures) + SciPy approx_derivative for eq, "
                         "ineq and cost",
            "port": "oracle port: numpy callbacks + restated SciPy 2-point FD for eq, ineq and cost"}


def run_reference(args, rank, world):
    """The reference's own CPU implementation of the path on every host core: each step evaluates a
    bounded sample (8 seeded instances per worker) of the workload; value = best of the timed steps'
    rates is NOT used -- the line reports total instances / total wall time like the CUDA arm."""
    if rank != 0:
        return
    cfg, _ = WORKLOADS[args.workload]
    procs = host_procs()
    kind = cpu_kind()
    per_step = procs * 8                          # >= 8 instances per worker per step: pool dispatch stays < 1 %
    vals = []
    pool = CpuPool(cfg, procs, kind)
    for _ in range(args.warmup):
        pool.run(per_step)
    for k in range(args.steps):
        v, wall, _ = pool.run(per_step, first=k * per_step)
        vals.append((v, wall))
    pool.close()
    total_wall = sum(w for _, w in vals)
    value = per_step * len(vals) / total_wall
    sample = "%d seeded instances of %s per step (%s), %d processes, 1 BLAS thread each" % (
        per_step, cfg, CPU_WHAT[kind], procs)
    out = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world,
           "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total_wall / len(vals),
           "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
           "data": "synthetic",
           "config": {"workload": "%s_b%d" % (args.workload, args.batch), "sample_per_step": per_step},
           "cpu_baseline": dict(value=value, unit=UNIT, cores=procs, kind=kind, sample=sample,
                                best_step_value=max(v for v, _ in vals), **cpu_info()),
           "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
           "gpu_launches": 0}
    print(json.dumps(out), flush=True)


# ------------------------------------------------------------------ checks and extras of the CUDA arm
def parity_check(eng, cfg, workloads, P3):
    """c and J of the first instances against the numpy oracle, outside the timed region (the tolerances are
    the tests': |dc| <= 1e-6 max(|c|, 1e-6 terms), |dJ| <= 1e-6 rowmax, zero pattern up to FD dust)."""
    import numpy as np
    from oracle import og_numpy
    wo = workloads.build(cfg, og_numpy)
    lb, ub = og_numpy.bounds_arrays(wo.prob)
    c, J = eng.eval_fd(P3)
    c, J = c.cpu().numpy(), J.cpu().numpy()
    out = {"instances": int(len(P3)), "c_max_rel": 0.0, "J_rowscaled_max": 0.0, "zero_mismatch": 0, "ok": True}
    for b in range(len(P3)):
        c_ref, J_ref = og_numpy.eval_fd(wo.prob, wo.obj, P3[b], lb, ub)
        x = np.clip(P3[b], lb, ub)
        terms = (np.abs(J_ref) * np.abs(x)[None, :]).sum(axis=1)
        scale = np.maximum(np.abs(c_ref), 1e-6 * terms)
        out["c_max_rel"] = max(out["c_max_rel"], float((np.abs(c[b] - c_ref) / np.maximum(scale, 1e-300)).max()))
        rowmax = np.abs(J_ref).max(axis=1, keepdims=True)
        Jb = J[b].T
        out["J_rowscaled_max"] = max(out["J_rowscaled_max"], float((np.abs(Jb - J_ref) / np.maximum(rowmax, 1e-300)).max()))
        mism = (Jb == 0) != (J_ref == 0)
        big = np.maximum(np.abs(Jb), np.abs(J_ref)) > 1e-7 * rowmax
        out["zero_mismatch"] += int((mism & big).sum())
        out["zero_mismatch_dust"] = out.get("zero_mismatch_dust", 0) + int((mism & ~big).sum())
    out["ok"] = bool(out["c_max_rel"] <= 1e-6 and out["J_rowscaled_max"] <= 1e-6 and out["zero_mismatch"] == 0)
    out["against"] = "oracle/og_numpy.py (bit-identical to the reference-generated goldens, tests/test_oracle.py)"
    assert out["ok"], "bench.py parity check failed: %r" % (out,)
    return out


def tensor_check(eng, wl, P, torch):
    """K1 (the batched FP64 tensor-core GEMM D.X) timed alone against the FP64 DGEMM peak measured here with
    cuBLAS (torch.matmul f64 4096^3, best of 5) -- MEASURED_PEAKS.json has no FP64 entry."""
    dev = P.device
    N = 4096
    a = torch.randn((N, N), dtype=torch.float64, device=dev)
    b = torch.randn((N, N), dtype=torch.float64, device=dev)
    for _ in range(2):
        torch.matmul(a, b)
    best = float("inf")
    for _ in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        torch.matmul(a, b)
        e1.record()
        e1.synchronize()
        best = min(best, e0.elapsed_time(e1))
    peak = 2.0 * N ** 3 / (best * 1e-3) / 1e12
    del a, b
    DX = eng.dx_gemm(P, clip=True)
    for _ in range(3):
        eng.dx_gemm(P, out=DX, clip=True)
    k1 = float("inf")
    for _ in range(10):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        eng.dx_gemm(P, out=DX, clip=True)
        e1.record()
        e1.synchronize()
        k1 = min(k1, e0.elapsed_time(e1))
    prob = wl.prob
    flop = 2.0 * sum(N_ * N_ * ns for N_, ns in zip(prob.nodes, prob.number_of_states)) * P.shape[0]
    pipe = None
    try:
        pipe = json.load(open(os.path.join(ROOT, "profiles", "traffic.json"))).get("k1_tensor_pipe_pct")
    except (OSError, ValueError):
        pass
    return {"kernel": "ogb_dx_gemm2_kernel (K1, mma.sync m8n8k4 f64 = DMMA; tcgen05 has no f64 kind)",
            "flop_per_launch": flop, "k1_ms_alone": k1, "k1_tflops": flop / (k1 * 1e-3) / 1e12,
            "k1_ms_alone_how": "one launch between two CUDA events on an idle stream, best of 10: includes the launch "
                               "latency of an idle GPU; k1_ms_in_step (ms_per_step - K2 alone, cold L2) is what a step pays",
            "dgemm_peak_tflops": peak, "dgemm_peak_how": "torch.matmul float64 4096^3 (cuBLAS), best of 5, CUDA events",
            "frac": flop / (k1 * 1e-3) / 1e12 / peak, "pipe_pct_from_ncu": pipe,
            "note": "K1 is 2 N^2 nstates FLOP per instance (15 kFLOP at Goddard-50): one wave of warps, latency bound "
                    "by construction, ~3 % of a step"}


def extras(eng, wl, cfg, workloads, P, P_host, torch, args, api):
    import numpy as np
    dev = P.device
    B = P.shape[0]
    n, M = eng.nvars, eng.nrows
    out = {}

    def best_ms(fn, reps=5):
        for _ in range(2):
            fn()
        best = float("inf")
        for _ in range(reps):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            fn()
            e1.record()
            e1.synchronize()
            best = min(best, e0.elapsed_time(e1))
        return best
    # ---- sparse FD and exact Jacobians (packed output), device-resident
    try:
        nnz = eng.nnz
        c = torch.empty((B, M), dtype=torch.float64, device=dev)
        vals = torch.empty((B, nnz), dtype=torch.float64, device=dev)
        sparse_bytes = 8 * n + 8 * M + 8 * nnz
        peak = 6437.9
        try:
            peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))).get("hbm_gbs", peak))
        except (OSError, ValueError):
            pass
        for name, fn in (("sparse_fd", lambda: eng.eval_sparse(P, out_c=c, out_vals=vals)),
                         ("exact", lambda: eng.eval_exact(P, out_c=c, out_vals=vals))):
            ms = best_ms(fn)
            out[name] = {"evals_per_s": B / (ms * 1e-3), "ms_per_step": ms, "nnz": int(nnz),
                         "algorithmic_bytes_per_eval": sparse_bytes,
                         "roofline": {"bound": "hbm", "achieved": B * sparse_bytes / (ms * 1e-3) / 1e9, "peak": peak,
                                      "unit": "GB/s", "frac": B * sparse_bytes / (ms * 1e-3) / 1e9 / peak,
                                      "note": "K1 + the sweep kernel with packed output; latency / issue bound, "
                                              "not HBM bound (8 n + 8 M + 8 nnz bytes per eval)"}}
        # how far the two Jacobians are apart (FD truncation + rounding noise, row-scaled)
        _, v_fd = eng.eval_sparse(P[:64])
        _, v_ex = eng.eval_exact(P[:64])
        Jf, Je = eng.densify(v_fd), eng.densify(v_ex)
        rowmax = Je.abs().amax(dim=1, keepdim=True).clamp_min(1e-300)
        out["exact"]["fd_vs_exact_rowscaled_max"] = float(((Jf - Je).abs() / rowmax).max())
        del Jf, Je, c, vals
    except Exception as ex:
        out["sparse_error"] = str(ex)[:200]
    # ---- a single instance per call (what Problem.solve does at every SLSQP iteration): latency
    try:
        import time as _t
        eng1 = wl.prob.compile(wl.obj, device=dev, jit=False)
        x = P_host[0].numpy().copy()
        for _ in range(5):
            eng1.eval_fd_host(x)
        t0 = _t.perf_counter()
        for _ in range(50):
            eng1.eval_fd_host(x)
        fd_us = (_t.perf_counter() - t0) / 50 * 1e6
        t0 = _t.perf_counter()
        for _ in range(50):
            eng1.eval_host(x)
        out["latency_b1"] = {"eval_fd_host_us": fd_us, "eval_host_us": (_t.perf_counter() - t0) / 50 * 1e6,
                             "what": "one instance, host vector in, c and the dense J back in host memory "
                                     "(pinned staging, interpreter kernel as Problem.solve uses it)"}
    except Exception as ex:
        out["latency_error"] = str(ex)[:200]
    # ---- the product's real end to end: a batched multi-start solve (SLSQP cores in worker processes fed by
    #      the scatter transport), next to the reference's way of doing the same iterations on this host
    try:
        import time as _t
        from opengoddard_b200 import sqp
        S = max(1, int(args.solve_starts))
        iters = 6
        procs = host_procs()
        X0 = P_host[:S].numpy().copy()
        wl.prob._engine, wl.prob._engine_key = eng, wl.prob._fingerprint(wl.obj, True)
        with sqp.WorkerPool(min(procs, S)) as pool:          # started (and warmed by one short solve) before the clock
            wl.prob.solve_batch(X0[:max(2, min(S, procs))], wl.obj, ftol=1e-10, maxiter=1, max_outer=1, processes=pool)
            t0 = _t.perf_counter()
            res = wl.prob.solve_batch(X0, wl.obj, ftol=1e-10, maxiter=iters, max_outer=1, processes=pool)
            wall = _t.perf_counter() - t0
        done = int(np.asarray(res["nit"]).sum())
        ev = eng.host_evaluator()
        t0 = _t.perf_counter()
        reps = 3
        Cs = np.zeros((S, n, M - 1))
        Gs = np.zeros((S, n))
        for _ in range(reps):
            ev.eval_fd_scatter(X0, [Cs.ctypes.data + b * Cs.strides[0] for b in range(S)], M - 1, M - 1,
                               [Gs.ctypes.data + b * Gs.strides[0] for b in range(S)])
        dev_s = (_t.perf_counter() - t0) / reps
        out["solve_batch"] = {"starts": S, "slsqp_iterations_each": iters, "processes": procs,
                              "instance_iterations_per_s": done / wall, "wall_s": wall, "instance_iterations": done,
                              "device_eval_s_per_iteration": dev_s,
                              "device_share": min(1.0, dev_s * (iters + 1) / wall),
                              "what": "Problem.solve_batch: lock-step SLSQP (SciPy's C core, one state per instance "
                                      "in worker processes, already running) over one batched device evaluation per round"}
        out["solve_batch"]["reference"] = reference_solve_rate(cfg, iters)
    except Exception as ex:
        out["solve_error"] = str(ex)[:300]
    # ---- the same multi-start with the whole SLSQP iteration on the device (opt-in, qp="device"): the packed sweep
    #      and one thread block per instance for the QP / line search / BFGS; only a mode word per instance returns
    try:
        import time as _t
        Sd = max(1, min(B, 4 * max(1, int(args.solve_starts))))
        iters = 6
        X0 = P_host[:Sd].numpy().copy()
        with eng.device_sqp(Sd, 1e-10, iters) as dq:
            dq.solve(X0[:min(Sd, 8)])
            torch.cuda.synchronize()
            t0 = _t.perf_counter()
            res = dq.solve(X0)
            torch.cuda.synchronize()
            wall = _t.perf_counter() - t0
            sc = dq.k.scalars(Sd)
            clk = {k[4:]: float(sc[k].sum()) for k in sc if k.startswith("clk_") and k != "clk_lsei_phase1"}
            tot = sum(clk.values()) or 1.0
            done = int(res["nit"].sum())
            out["solve_batch_device_qp"] = {
                "starts": Sd, "slsqp_iterations_each": iters, "instance_iterations_per_s": done / wall, "wall_s": wall,
                "instance_iterations": done, "rounds": int(res["rounds"]), "device_bytes": dq.bytes,
                "qp_phase_share": {k: round(v / tot, 3) for k, v in clk.items()},
                "what": "Problem.solve_batch(qp='device'): Kraft's SLSQP restated as one thread block per instance "
                        "(ogb_sqp_step) on the packed sweep's output; host wall clock, synchronised"}
    except Exception as ex:
        out["solve_device_error"] = str(ex)[:300]
    return out


def reference_solve_rate(cfg, iters):
    """SLSQP iterations per second the CPU path achieves on one core for the same problem: scipy.minimize on
    the reference's (or the port's) closures, FD Jacobians by SciPy -- what Problem.solve of the reference does."""
    import contextlib
    import io
    import numpy as np
    from scipy import optimize
    from opengoddard_b200 import workloads
    kind = cpu_kind()
    if kind == "reference":
        from oracle import ref_loader
        mod = ref_loader.load_reference()
        wl = workloads.build(cfg, mod)
        cap = ref_loader.capture_solve(mod, wl.prob, wl.obj)
        fun, args_, cons, jac, bounds = cap.fun, cap.args, cap.constraints, cap.jac, cap.bounds
    else:
        from oracle import og_numpy
        wl = workloads.build(cfg, og_numpy)
        prob, obj = wl.prob, wl.obj
        fun, args_, jac, bounds = (lambda x: prob.eval_cost(x, obj)), (), None, prob.bounds
        cons = ({"type": "eq", "fun": lambda x: prob.eval_equality(x, obj)},
                {"type": "ineq", "fun": lambda x: prob.eval_inequality(x, obj)})
    X0 = workloads.make_batch(wl, 2)
    t0 = time.perf_counter()
    nit = 0
    for x0 in X0:
        with contextlib.redirect_stdout(io.StringIO()):
            opt = optimize.minimize(fun, x0, args=args_, bounds=bounds, constraints=cons, jac=jac, method="SLSQP",
                                    options={"disp": False, "maxiter": iters, "ftol": 1e-10})
        nit += int(opt.nit)
    wall = time.perf_counter() - t0
    return {"kind": kind, "cores": 1, "instance_iterations_per_s_per_core": nit / wall,
            "ms_per_iteration": 1e3 * wall / max(1, nit), "instances": len(X0)}


# ------------------------------------------------------------------ the CUDA arm
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="goddard50", choices=sorted(WORKLOADS))
    ap.add_argument("--batch", type=int, default=0, help="instances per GPU (default: per workload)")
    ap.add_argument("--e2e-steps", type=int, default=0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e-variants", action="store_true")
    ap.add_argument("--no-autotune", action="store_true",
                    help="keep the sweep kernel's default CTA size (256 threads) instead of timing 256 / 384 / 128 first")
    ap.add_argument("--autotune", action="store_true", help="(default; kept for older command lines)")
    ap.add_argument("--host-threads", type=int, default=0, help="host threads of the e2e session (default: cores / ranks)")
    ap.add_argument("--no-extras", action="store_true", help="skip parity / tensor / latency / sparse / solve extras")
    ap.add_argument("--solve-starts", type=int, default=256, help="multi-start instances of the solve_batch extra")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    cfg, default_batch = WORKLOADS[args.workload]
    if args.batch <= 0:
        args.batch = default_batch
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    # CPU baseline first (rank 0 at N=1 only), before CUDA is touched in this process
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        procs = host_procs()
        kind = cpu_kind()
        sample = max(256, procs * 8)
        pool = CpuPool(cfg, procs, kind)
        v_all, wall_all, cpu_s = max(pool.run(sample) for _ in range(3))     # best of 3 (noisy shared hosts)
        pool.close()
        v_one, _, _ = CpuPool(cfg, 1, kind).run(48)
        cpu = dict(value=v_all, unit=UNIT, cores=procs, kind=kind,
                   sample="%d seeded instances of %s (same generator as the GPU batch), %s; best of 3 passes, "
                          "%.1f s of CPU work each" % (sample, cfg, CPU_WHAT[kind], cpu_s),
                   single_core_value=v_one, **cpu_info())

    import numpy as np
    import torch
    import OpenGoddard.optimize as api
    from opengoddard_b200 import workloads

    assert torch.cuda.is_available(), "bench.py needs a CUDA device (there is no CPU fallback)"
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        if os.environ.get("NCCL_DEBUG", "").upper() in ("", "VERSION"):
            os.environ["NCCL_DEBUG"] = "WARN"          # keep NCCL's version banner off stdout (one JSON line)
        dist.init_process_group("nccl", device_id=dev)

    wl = workloads.build(cfg, api)
    eng = wl.prob.compile(wl.obj, device=dev)
    B = args.batch
    tuned = None
    # this rank's shard of the seeded global batch (instances rank*B .. rank*B + B - 1)
    P_host = torch.from_numpy(workloads.make_batch(wl, B, first=rank * B)).pin_memory()
    P = P_host.to(dev)
    if not args.no_autotune:                            # set-up, before anything is timed: the engine times
        try:                                            # its 256- / 384- / 128-thread sweep kernels on this batch
            tuned = eng.autotune(P)
        except Exception as ex:                         # never fatal: the default CTA size stays
            tuned = {"error": str(ex)[:200]}
            try:
                eng.set_option(1, 256)
            except Exception:
                pass
        torch.cuda.empty_cache()
    n, M = eng.nvars, eng.nrows
    parity = tensor = None
    if rank == 0 and not args.no_extras:
        parity = parity_check(eng, cfg, workloads, P_host.numpy()[:3])
        tensor = tensor_check(eng, wl, P, torch)
    c = torch.empty((B, M), dtype=torch.float64, device=dev)
    J = torch.empty((B, n, M), dtype=torch.float64, device=dev)
    DX = torch.empty((B, eng.ndx), dtype=torch.float64, device=dev)
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
    bytes_per_eval = 8 * n + 8 * M * (n + 1)

    def barrier():
        torch.cuda.synchronize(dev)
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def step(timed):
        """One pass of the hot path = what ogb_eval_fd enqueues: K1 (D.X tensor-core GEMM into the
        scratch) then K2 (fused sweep), on the current stream; events around the step and K2."""
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)] if timed else None
        if timed:
            ev[0].record()
        eng.dx_gemm(P, out=DX, clip=True)
        if timed:
            ev[1].record()
        eng.sweep_fd(P, DX, c, J)
        if timed:
            ev[2].record()
        return ev

    for _ in range(args.warmup):
        step(False)
        flush.zero_()
    launches0 = eng.launches
    sampler = ClockSampler(local_rank)
    barrier()
    sampler.start()
    events = []
    for _ in range(args.steps):
        events.append(step(True))
        flush.zero_()                                   # L2 flush, outside the event pairs
    barrier()
    clocks = sampler.stop()
    launches = eng.launches - launches0
    step_ms = [e[0].elapsed_time(e[2]) for e in events]
    sweep_ms = [e[1].elapsed_time(e[2]) for e in events]
    gemm_ms = [e[0].elapsed_time(e[1]) for e in events]
    total_ms = torch.tensor([sum(step_ms)], dtype=torch.float64, device=dev)
    if dist is not None:
        dist.all_reduce(total_ms, op=dist.ReduceOp.MAX)
    total_ms = float(total_ms.item())
    value = world * B * args.steps / (total_ms * 1e-3)

    # ---- the one exchange step of the path: all-gather of the per-instance
    #      decision vectors and costs over NCCL, outside the timed steps, reported on its own
    gather = None
    if dist is not None:
        from opengoddard_b200 import batch as ogb_batch
        cost = c[:, M - 1].contiguous()
        for _ in range(2):
            ogb_batch.gather_rows(P, world * B)
            ogb_batch.gather_rows(cost, world * B)
        barrier()
        g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        g0.record()
        allP = ogb_batch.gather_rows(P, world * B)
        allcost = ogb_batch.gather_rows(cost, world * B)
        g1.record()
        barrier()
        gms = torch.tensor([g0.elapsed_time(g1)], dtype=torch.float64, device=dev)
        dist.all_reduce(gms, op=dist.ReduceOp.MAX)
        assert allP.shape[0] == world * B and torch.equal(allP[rank * B:(rank + 1) * B], P)
        gather = {"what": "all-gather of p (B x nvars) and cost (B) from every rank (NCCL), after the timed steps",
                  "ms": float(gms.item()), "bytes_per_rank": int(P.numel() * 8 + cost.numel() * 8)}
        del allP, allcost

    # ---- end to end through the host-buffer C-ABI entry point (ogb_host_eval_fd): p in pinned HOST
    #      memory in, c and the dense J in HOST memory out, every copy inside the timed region.
    #      The call is synchronous and its last stage runs on host threads, so the clock is the
    #      host's (perf_counter between barriers), MAX over ranks.
    del J, flush
    torch.cuda.empty_cache()
    e2e_steps = args.e2e_steps or max(3, min(args.steps, 8))
    local_world = int(os.environ.get("LOCAL_WORLD_SIZE", str(world)))
    threads = args.host_threads or max(1, host_procs() // max(1, local_world))
    sess = eng.host_session(B, threads=threads)
    hc = np.empty((B, M), dtype=np.float64)
    hJ = np.empty((B, n, M), dtype=np.float64)

    def time_mode(mode, Jbuf, steps):
        sess.eval_fd(P_host, hc, Jbuf, mode=mode)           # warm-up (also faults the pages in)
        barrier()
        t0 = time.perf_counter()
        for _ in range(steps):
            sess.eval_fd(P_host, hc, Jbuf, mode=mode)
        dt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
        barrier()
        if dist is not None:
            dist.all_reduce(dt, op=dist.ReduceOp.MAX)
        st = sess.stats()
        return world * B * steps / float(dt.item()), st

    e2e_value, st = time_mode("dense", hJ, e2e_steps)
    e2e = {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(st.h2d_bytes),
           "d2h_bytes_per_step": int(st.d2h_bytes), "steps": e2e_steps,
           "api": "ogb_host_eval_fd(mode=OGB_HOST_J_DENSE): pinned host p -> H2D -> K1 -> K2a (packed sweep) -> D2H of c "
                  "and the packed non-zeros -> %d host threads rewrite the whole dense J (zeros included) in "
                  "pageable host memory; %d chunks of %d instances; host wall clock" % (st.threads, st.nchunks, st.chunk),
           "nnz_per_instance": int(st.nnz), "host_threads": int(st.threads)}
    if not args.no_e2e_variants:
        # the same call with the other transports, for context (not the headline)
        var = {}
        var["keep_zeros"], _ = time_mode("keep_zeros", hJ, max(2, e2e_steps // 2))
        hv = np.empty((B, int(st.nnz)), dtype=np.float64)
        var["packed"], _ = time_mode("packed", hv, max(2, e2e_steps // 2))
        del hv
        try:
            hJp = torch.empty((B, n, M), dtype=torch.float64).pin_memory()
            var["dma_dense_pinned"], st2 = time_mode("dma", hJp, 2)
            e2e["dma_d2h_bytes_per_step"] = int(st2.d2h_bytes)
            del hJp
        except Exception as ex:                              # pinned allocation can fail on a small host
            var["dma_dense_pinned"] = None
            e2e["dma_error"] = str(ex)[:120]
        e2e["variants"] = var
    launches_e2e = int(st.launches)
    # ---- the transport the SQP driver really uses (ogb_host_eval_fd_scatter): the packed non-zeros of every
    #      instance scattered straight into that instance's persistent SLSQP buffers (C: Fortran (m, n), g: (n,)),
    #      which keep their zero background -- no dense J is rebuilt anywhere.  Bound: PCIe (8 nnz + 8 M bytes
    #      per eval over the device->host link).
    e2e_sqp = None
    try:
        m = M - 1
        del hJ
        Cbuf = np.zeros((B, n, max(1, m)), dtype=np.float64)          # per instance an (m, n) Fortran matrix
        Gbuf = np.zeros((B, n), dtype=np.float64)
        Cp = [Cbuf.ctypes.data + b * Cbuf.strides[0] for b in range(B)]
        Gp = [Gbuf.ctypes.data + b * Gbuf.strides[0] for b in range(B)]
        sess.eval_fd_scatter(P_host, hc, Cp, max(1, m), m, Gp)        # warm-up (faults the pages in)
        barrier()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            sess.eval_fd_scatter(P_host, hc, Cp, max(1, m), m, Gp)
        dt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
        barrier()
        if dist is not None:
            dist.all_reduce(dt, op=dist.ReduceOp.MAX)
        st3 = sess.stats()
        e2e_sqp = {"value": world * B * e2e_steps / float(dt.item()), "unit": UNIT,
                   "h2d_bytes_per_step": int(st3.h2d_bytes), "d2h_bytes_per_step": int(st3.d2h_bytes),
                   "api": "ogb_host_eval_fd_scatter: pinned host p -> H2D -> K1 -> K2a -> D2H of c and the packed "
                          "non-zeros -> host threads scatter them into B persistent per-instance SLSQP buffers "
                          "(C (m, n) Fortran + g), zero background kept; host wall clock",
                   "pcie_bound_evals_per_s_at_55GBs": 55e9 / (8.0 * (st3.nnz + M))}
        del Cbuf, Gbuf
    except Exception as ex:
        e2e_sqp = {"error": str(ex)[:200]}
    sess.close()
    extra = None
    if rank == 0 and world == 1 and not args.no_extras:      # (N = 1 only, like the CPU baseline: the other ranks would idle)
        extra = extras(eng, wl, cfg, workloads, P, P_host, torch, args, api)

    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except (OSError, ValueError):
            pass
        peak = float(peaks.get("hbm_gbs", 6650.0))
        peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback 6.65 TB/s"
        sweep_avg_ms = sum(sweep_ms) / len(sweep_ms)
        achieved = B * bytes_per_eval / (sweep_avg_ms * 1e-3) / 1e9
        traffic = None
        try:
            tr = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
            key = "%s_b%d" % (args.workload, B)
            traffic = tr.get(key, {}).get("dram_bytes_per_launch")
        except (OSError, ValueError):
            pass
        out = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": total_ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": "%s_b%d" % (args.workload, B), "instances_per_gpu": B, "nvars": n,
                       "rows": M, "meq": eng.meq, "mineq": eng.mineq, "bytes_per_eval": bytes_per_eval,
                       "l2": "256 MiB memset between timed steps (outside the event pairs); every step "
                             "also writes %.2f GB of Jacobian (L2 is 126 MB)" % (B * bytes_per_eval / 1e9),
                       "parallelism": "instance batch sharded over %d GPU(s), no data-path collective" % world,
                       "sweep_cta_threads": int(getattr(eng, "tuned_threads", 256)),
                       "autotune_ms": tuned},
            "clocks": clocks,
            "e2e": e2e,
            "gpu_launches": launches, "gpu_launches_per_e2e_step": launches_e2e,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": traffic, "kernel": "ogb_sweep_kernel (K2)",
                         "kernel_ms": sweep_avg_ms, "bytes_per_launch": B * bytes_per_eval,
                         "peak_source": peak_src},
            "cpu_baseline": cpu,
            "e2e_sqp": e2e_sqp,
            "parity": parity,
            "tensor": tensor,
            "extra": extra,
        }
        if tensor is not None:
            k1 = sum(gemm_ms) / len(gemm_ms)
            tensor["k1_ms_in_step"] = k1
            tensor["k1_tflops_in_step"] = tensor["flop_per_launch"] / (k1 * 1e-3) / 1e12
            tensor["frac_in_step"] = tensor["k1_tflops_in_step"] / tensor["dgemm_peak_tflops"]
        if gather is not None:
            out["gather"] = gather
        print(json.dumps(out), flush=True)
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
