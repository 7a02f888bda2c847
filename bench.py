#!/usr/bin/env python
"""bench.py -- headline benchmark of the BRDF-fitting hot path.

    python bench.py --gpus N --steps K --warmup W            (N>1: launched under torchrun)
    python bench.py --impl reference ...                     (the reference's CPU levmar, same metric)

Workload (BASELINE.json configs[1]): one global Blinn-Phong fit of 10^6 synthetic fp64 samples per
GPU with the reference's global preset (brdfdata.cpp:1002,1046-1058: p0=0, bounds [0,100]^3,
itmax 2000, opts {1e-3,1e-15,1e-10,1e-50, delta=1}).  One "step" = one complete fit.

    metric  LM sample-evals/sec = info[7] * n_total / time, with levmar's own accounting of
            function evaluations (a difference Jacobian counts m+1, lmbc_core.c:1119-1124)
    value   samples already resident in HBM when the timed region starts
    e2e     the same fit through the levmar-signature C-ABI call brdfgpu_dlevmar_bc_dif with HOST
            buffers (pinned), host<->device copies inside the timed region
    seconds_per_fit / time_to_solution  both arms solve the same problem to the same answer but count
            different numbers of evaluations (summation order), so seconds per fit is the
            like-for-like comparison with the CPU; it sits at the top level next to the metric
    roofline      the dominant kernel of the step (k_persistent_fit): 24 B x samples x PHYSICAL sweeps / its
                  CUDA-event duration against the measured HBM peak.  At 10^6 samples the shard lives in shared
                  memory, so this is a small fraction by construction; the binding limits (fp64 pipe, exchange
                  latency) are named next to it
    roofline_hbm  K2 (fused residual + difference Jacobian + J^T J) at 10^8 samples, inputs far larger than
                  L2 -- the regime BASELINE.json's 60 %-of-HBM target names -- plus K3 and a complete fit
    stages        the other two stages of the path, each with a GPU number, a CPU number timed on this box
                  and a roofline fraction: batched per-face fits (configs[3]) and the gather (cup, bunny x 13)
    cpu_baseline  the reference's single-threaded levmar timed on this box's host cores on the same workload

Multi-GPU (N > 1, weak scaling): every rank holds its own 10^6-sample shard of ONE fit over N x 10^6 samples, each
evaluation all-reduces 11 doubles inside the persistent kernel (peer stores over NVLink).  A fit over N shards is a
different problem for every N and levmar walks a different path on each (33 / 95 / 39 / 890 iterations at 1 / 2 / 4 /
8 GPUs; replicating one shard on every rank does not help: doubling every sum changes the projected-gradient step
lengths and the trajectory with them -- 93 iterations at N = 2), so value(N) / value(1) mixes the cost of the exchange
with the trajectory.  `scaling_invariant` is the figure without that: a scripted sequence of Jacobian and cost sweeps,
each with its exchange, timed in the same kernel at every N (us per sweep).  Before any timing a small ragged sharded
fit is checked against one GPU and against the CPU oracle (`parity`; the run fails outside 1e-4 / 1e-6), and the line
carries `scale_hbm` (10^7 and 10^8 samples in total, strong-scaled, and 10^8 per GPU, weak), `bunny` (configs[2]:
13-view gather sharded by view + global fit against the golden values) and `batched` (configs[3] strong and weak).
"""
import argparse
import ctypes as C
import hashlib
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import numpy as np  # noqa: E402

N_PER_GPU = 1_000_000
SEED = 88172645463325252
METRIC = "LM sample-evals/sec (global BRDF fit, levmar nfev accounting)"
UNIT = "sample-evals/s"
WORKLOAD = "synthetic single global BRDF fit, 10^6 samples, fp64 (BASELINE configs[1])"
# fp64-pipe utilisation of the two fit kernels from the committed ncu --set full captures (profiles/r02_ncu_tables.md)
NCU = {"k_persistent_fit": {"fp64_pipe_pct": 38.0, "issue_slots_pct": 32.0, "dram_bytes": 40.98e6, "source": "profiles/r01_ncu_tables.md (r01_persist_spec)"},
       "k_batched_fit<32,2>": {"fp64_pipe_pct": 29.0, "dram_bytes": 101.3e6, "source": "profiles/r01_ncu_tables.md (batch_walk)"}}
try:
    with open(os.path.join(ROOT, "profiles", "r02_ncu.json")) as _f:
        NCU.update(json.load(_f))
except Exception:
    pass


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks and throttle reasons during the timed region."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"], stdout=subprocess.PIPE, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


# ----------------------------------------------------------------------------------------------
# reference arm / cpu baselines: the reference's levmar (and the oracle's gather) on the host
# ----------------------------------------------------------------------------------------------
def cpu_solver():
    import oracle_lib as O
    ref = O.ref()
    O.oracle()
    return (ref, "", "reference") if ref is not None else (O.oracle(), "oracle_", "port")


def cpu_reference_fit(n, itmax=None, seed=SEED):
    """One dlevmar_bc_dif BRDF fit on the CPU.  Returns (seconds, nfev, kind, p, info)."""
    import oracle_lib as O
    import synth
    c, td, th, x = synth.samples(n, seed=seed)
    lib, prefix, kind = cpu_solver()
    preset = dict(O.REF_GLOBAL)
    if itmax is not None:
        preset["itmax"] = itmax
    t0 = time.perf_counter()
    ret, p, info = O.brdf_fit(lib, prefix, c, td, th, x, 1, preset)
    dt = time.perf_counter() - t0
    return dt, float(info[7]), kind, p, info


def cpu_batched_fits(count, nper=64, seed=2026):
    """The first `count` fits of BASELINE configs[3] through the reference's per-face call (brdfdata.cpp:1119), one
    thread.  Returns (seconds, kind, mean nfev)."""
    import oracle_lib as O
    import synth
    c, td, th, x, _ = synth.batched(count, nper, seed=seed)
    lib, prefix, kind = cpu_solver()
    O.brdf_fit(lib, prefix, c[0], td[0], None, x[0], 1, O.REF_PERFACE)
    nfev = 0.0
    t0 = time.perf_counter()
    for f in range(count):
        _, _, info = O.brdf_fit(lib, prefix, c[f], td[f], None, x[f], 1, O.REF_PERFACE)
        nfev += info[7]
    return time.perf_counter() - t0, kind, nfev / count


def run_reference_arm(args):
    """The reference's own CPU path (levmar 2.6 compiled unmodified into oracle/_ref, else the oracle port) on
    the same workload: COMPLETE fits, so the mix of Jacobian and line-search evaluations is the one the GPU arm
    sees.  One fit takes ~5.5 s on one host core (the reference is single-threaded by construction); when K
    fits would not finish within ~2.5 minutes, fewer fits are timed and the line says how many."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    n = N_PER_GPU
    for _ in range(max(0, args.warmup)):
        cpu_reference_fit(n // 10, 3)      # page in the library and the data path
    total_t, total_ev, kind, fits = 0.0, 0.0, "port", 0
    budget = 150.0
    while fits < args.steps:
        dt, nfev, kind, _, info = cpu_reference_fit(n)
        total_t += dt
        total_ev += nfev * n
        fits += 1
        if total_t + dt > budget:
            break
    value = total_ev / total_t
    sample = "%d complete fit(s) of n=%d samples, REF_GLOBAL preset (%d iterations, %d function evaluations each)" % (
        fits, n, int(info[5]), int(info[7]))
    out = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
           "warmup": args.warmup, "ms_per_step": 1e3 * total_t / fits, "seconds_per_fit": total_t / fits,
           "higher_is_better": True, "scaling": "weak",
           "vs_baseline": None, "dtype": "f64", "data": "synthetic", "timed_fits": fits,
           "config": {"workload": WORKLOAD, "samples_per_gpu": n, "preset": "REF_GLOBAL", "model": "blinn-phong"},
           "cpu_baseline": {"value": value, "unit": UNIT, "cores": 1, "kind": kind, "sample": sample,
                            "host_cores_available": os.cpu_count()},
           "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
           "gpu_launches": 0}
    print(json.dumps(out))


# ----------------------------------------------------------------------------------------------
# GPU arm
# ----------------------------------------------------------------------------------------------
class Rig:
    """Context, stream, distributed plumbing and the timing helpers shared by every leg."""

    def __init__(self, args):
        import torch
        from brdf_b200 import api as A
        self.torch, self.A, self.args = torch, A, args
        self.rank = int(os.environ.get("RANK", "0"))
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        if self.world != args.gpus and self.world > 1:
            raise SystemExit("--gpus %d but WORLD_SIZE=%d" % (args.gpus, self.world))
        torch.cuda.set_device(self.local)
        self.ctx = A.Context(self.local)
        self.dist = None
        if self.world > 1:
            import torch.distributed as dist
            self.dist = dist
            dist.init_process_group("nccl", device_id=torch.device("cuda", self.local))
            ids = [A.comm_unique_id() if self.rank == 0 else None]
            dist.broadcast_object_list(ids, src=0)
            self.ctx.comm_init(ids[0], self.rank, self.world)
            # peer buffers for the fused in-kernel exchange (the export records travel through torch.distributed)
            handles = [None] * self.world
            dist.all_gather_object(handles, self.ctx.peer_export())
            self.ctx.peer_attach(handles, self.rank, self.world)
        self.stream = torch.cuda.ExternalStream(self.ctx.stream)
        self.peak, self.peak_src = peaks()

    def barrier(self):
        self.ctx.synchronize()
        self.torch.cuda.synchronize()
        if self.dist is not None:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def max_over_ranks(self, v):
        if self.dist is None:
            return float(v)
        t = self.torch.tensor([float(v)], device="cuda", dtype=self.torch.float64)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def timed(self, fn, reps=1):
        """CUDA events on the context's stream around `reps` calls, barrier on both sides, max over ranks (ms per call)."""
        torch = self.torch
        self.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(self.stream)
        out = None
        for _ in range(reps):
            out = fn()
        e1.record(self.stream)
        self.barrier()
        return self.max_over_ranks(e0.elapsed_time(e1) / reps), out

    def gather_bytes(self, blob):
        if self.dist is None:
            return [blob]
        out = [None] * self.world
        self.dist.all_gather_object(out, blob)
        return out


def fit_blob(ret, p, info):
    return np.concatenate([[float(ret)], np.asarray(p, dtype=np.float64), np.asarray(info, dtype=np.float64)]).tobytes()


def multi_gpu_parity(rig, n_total=200_001):
    """Before any timing: one small global fit over RAGGED shards through the fused in-kernel exchange, checked for
    (a) bit-identical results on every rank, (b) agreement with ONE GPU fitting the whole set, (c) agreement with the
    CPU oracle.  Outside the parity bars (parameters 1e-4, cost 1e-6 relative) the run fails."""
    import oracle_lib as O
    import synth
    A, ctx, rank, world = rig.A, rig.ctx, rig.rank, rig.world
    c, td, th, x = synth.samples(n_total, seed=77)
    lo, hi = rank * n_total // world, (rank + 1) * n_total // world
    s = ctx.upload(c[lo:hi], td[lo:hi], x[lo:hi], A.BLINN_PHONG)
    out = {"samples_total": n_total, "shards": "ragged: rank r holds samples [r*n/N, (r+1)*n/N)", "presets": {}}
    ok = True
    for name, preset, opreset in (("REF_GLOBAL", A.REF_GLOBAL, O.REF_GLOBAL), ("REF_PERFACE", A.REF_PERFACE, O.REF_PERFACE)):
        ret, p, info = ctx.fit_global(s, preset, drive=A.DRIVE_PERSISTENT)
        blobs = rig.gather_bytes(fit_blob(ret, p, info))
        rec = {"ranks_bit_identical": all(b == blobs[0] for b in blobs), "hash": hashlib.sha256(b"".join(blobs)).hexdigest()[:16]}
        if rank == 0:
            single = A.Context(rig.local)
            s1 = single.upload(c, td, x, A.BLINN_PHONG)
            r1, p1, i1 = single.fit_global(s1, preset)
            s1.free()
            single.close()
            wret, wp, winfo = O.brdf_fit(O.oracle(), "oracle_", c, td, th, x, 1, opreset)
            rec.update(p=[float(v) for v in p], iterations=int(info[5]), stop_reason=int(info[6]),
                       p_rel_err_vs_1gpu=float(np.max(np.abs(p - p1) / np.abs(p1))),
                       cost_rel_err_vs_1gpu=float(abs(info[1] - i1[1]) / i1[1]),
                       p_rel_err_vs_oracle=float(np.max(np.abs(p - wp) / np.abs(wp))),
                       cost_rel_err_vs_oracle=float(abs(info[1] - winfo[1]) / winfo[1]),
                       oracle_stop_reason=int(winfo[6]), return_sign_equal=bool((ret >= 0) == (wret >= 0) == (r1 >= 0)))
            rec["pass"] = bool(rec["ranks_bit_identical"] and rec["return_sign_equal"] and rec["p_rel_err_vs_1gpu"] <= 1e-4 and
                               rec["p_rel_err_vs_oracle"] <= 1e-4 and rec["cost_rel_err_vs_1gpu"] <= 1e-6 and
                               rec["cost_rel_err_vs_oracle"] <= 1e-6)
            ok = ok and rec["pass"]
        out["presets"][name] = rec
    s.free()
    flag = [ok]
    if rig.dist is not None:
        rig.dist.broadcast_object_list(flag, src=0)
    out["pass"] = bool(flag[0])
    if not flag[0]:
        if rank == 0:
            sys.stderr.write("bench.py: multi-GPU parity FAILED: %s\n" % json.dumps(out))
        raise SystemExit(3)
    return out


def headline(rig, replicated):
    """K complete REF_GLOBAL fits of N x 10^6 samples, resident.  replicated: every rank holds the SAME shard."""
    A, ctx, args = rig.A, rig.ctx, rig.args
    n = N_PER_GPU
    s = ctx.synth(n, SEED, start=0 if replicated else rig.rank * n)
    drive = A.DRIVE_HOST if os.environ.get("BRDF_BENCH_DRIVE") == "host" else A.DRIVE_PERSISTENT
    for _ in range(max(3, args.warmup)):
        ret, p, info = ctx.fit_global(s, A.REF_GLOBAL, drive=drive)
    acc = {"evals": 0.0, "sweeps": 0.0, "levmar_passes": 0.0}
    launches0 = ctx.launches

    def step():
        ret, p, info = ctx.fit_global(s, A.REF_GLOBAL, drive=drive)
        st = ctx.fit_stats()
        acc["evals"] += info[7]
        acc["sweeps"] += st["jac_passes"] + st["cost_passes"]   # physical sweeps over the samples
        # passes the REFERENCE algorithm counts for the same trajectory: one per Jacobian (fused here; levmar makes m+1)
        # and one per counted cost evaluation
        acc["levmar_passes"] += info[8] + (info[7] - 4.0 * info[8])
        return ret, p, info, st
    ms, (ret, p, info, st) = rig.timed(step, args.steps)
    blobs = rig.gather_bytes(fit_blob(ret, p, info))
    return dict(samples=s, drive=drive, ms=ms, ret=ret, p=p, info=info, st=st, launches=ctx.launches - launches0,
                evals=acc["evals"] / args.steps, sweeps=acc["sweeps"] / args.steps, levmar_passes=acc["levmar_passes"] / args.steps,
                ranks_bit_identical=all(b == blobs[0] for b in blobs))


def scripted_sweeps(rig, s, rounds=300):
    """Trajectory-invariant scaling figure: `rounds` x (one Jacobian sweep + one cost sweep) over the resident shard inside
    the persistent kernel, every sweep followed by the grid-wide and cross-GPU exchange, NO LM control code: the same
    sequence at every rank count whatever the data, so us per sweep at 1 / 2 / 4 / 8 GPUs isolates what the exchange costs."""
    A, ctx = rig.A, rig.ctx
    preset = dict(A.REF_GLOBAL, itmax=rounds, p0=(0.6, 0.35, 12.0))
    os.environ["BRDFGPU_SPEC_JAC"] = "16"
    try:
        ctx.fit_global(s, preset)
        ms, _ = rig.timed(lambda: ctx.fit_global(s, preset), 3)
        st = ctx.fit_stats()
    finally:
        del os.environ["BRDFGPU_SPEC_JAC"]
    sweeps = st["jac_passes"] + st["cost_passes"]
    return {"what": "scripted: %d x (Jacobian sweep + cost sweep), each with its exchange, no LM control code" % rounds,
            "sweeps": sweeps, "ms": ms, "us_per_sweep": 1e3 * ms / sweeps,
            "cycles_per_sweep_cta0": {"sweep": st["cyc_sweep"] / sweeps, "exchange": st["cyc_exchange"] / sweeps,
                                      "exchange_phases": [v / sweeps for v in st["cyc_exchange_phases"]]}}


def e2e_single(rig, s):
    """The levmar-signature call with pinned host buffers: H2D of 24 MB + fit + D2H of the result, per step."""
    A, torch, args = rig.A, rig.torch, rig.args
    n = N_PER_GPU
    c_h, t_h, x_h = s.download()
    angles = torch.empty(3 * n, dtype=torch.float64).pin_memory()
    xs = torch.empty(n, dtype=torch.float64).pin_memory()
    angles[:n] = torch.from_numpy(c_h); angles[n:2 * n] = torch.from_numpy(t_h); angles[2 * n:] = 0.0
    xs[:] = torch.from_numpy(x_h)
    extra_data = A.ExtraData(C.cast(angles.data_ptr(), A.dptr), 1)
    g = A.REF_GLOBAL
    lb, ub, opts = (np.array(g[k], dtype=np.float64) for k in ("lb", "ub", "opts"))
    fn = A.func_address("brdfgpu_BRDFFunc")

    def call():
        pp = np.array(g["p0"], dtype=np.float64)
        inf = np.zeros(10)
        r = A.lib().brdfgpu_dlevmar_bc_dif(fn, A._d(pp), C.cast(xs.data_ptr(), A.dptr), 3, n, A._d(lb), A._d(ub), None,
                                           g["itmax"], A._d(opts), A._d(inf), None, None,
                                           C.cast(C.pointer(extra_data), C.c_void_p))
        return r, pp, inf
    for _ in range(2):
        call()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    ev = 0.0
    for _ in range(args.steps):
        r, pp, inf = call()
        ev += inf[7]
    dt = time.perf_counter() - t0
    return {"value": ev * n / dt, "unit": UNIT, "h2d_bytes_per_step": 3 * 8 * n, "d2h_bytes_per_step": 8 * (3 + 10) + 4,
            "ms_per_step": 1e3 * dt / args.steps, "seconds_per_fit": dt / args.steps,
            "call": "brdfgpu_dlevmar_bc_dif(brdfgpu_BRDFFunc, ...) with pinned host buffers", "p": [float(v) for v in pp]}


def e2e_multi(rig, s, drive):
    """Multi-GPU: every step moves this rank's shard host -> device (pinned), fits over all ranks' shards, reads the result."""
    A, torch, ctx, args = rig.A, rig.torch, rig.ctx, rig.args
    n = N_PER_GPU
    c_h, t_h, x_h = s.download()
    pin = [torch.from_numpy(a).pin_memory() for a in (c_h, t_h, x_h)]
    h = C.c_void_p()
    ctx._ok(A.lib().brdfgpu_samples_upload(ctx.handle, n, C.cast(pin[0].data_ptr(), A.dptr), C.cast(pin[1].data_ptr(), A.dptr),
                                           C.cast(pin[2].data_ptr(), A.dptr), 1, C.byref(h)))
    ss = A.Samples(ctx, h)
    rig.barrier()
    t0 = time.perf_counter()
    ev = 0.0
    for _ in range(args.steps):
        ctx._ok(A.lib().brdfgpu_samples_reload(ctx.handle, ss.handle, C.cast(pin[0].data_ptr(), A.dptr),
                                               C.cast(pin[1].data_ptr(), A.dptr), C.cast(pin[2].data_ptr(), A.dptr)))
        r, pp, inf = ctx.fit_global(ss, A.REF_GLOBAL, drive=drive)
        ev += inf[7]
    rig.barrier()
    dt = rig.max_over_ranks(time.perf_counter() - t0)
    ss.free()
    return {"value": ev * n * rig.world / dt, "unit": UNIT, "h2d_bytes_per_step": 3 * 8 * n, "d2h_bytes_per_step": 8 * 13 + 4,
            "ms_per_step": 1e3 * dt / args.steps, "seconds_per_fit": dt / args.steps,
            "call": "brdfgpu_samples_reload (pinned host shard -> device) + brdfgpu_fit_global per rank"}


# ---- the HBM regime: BASELINE configs[4] -----------------------------------------------------
def hbm_kernels(rig, n=100_000_000):
    """K2 / K3 alone at 10^8 samples and a complete fit (every evaluation streams 2.4 GB from HBM), one GPU."""
    A, ctx, peak = rig.A, rig.ctx, rig.peak
    s = ctx.synth(n, SEED)
    p = [0.6, 0.35, 12.0]
    res = {}
    for kind, name in ((0, "k_normal_eq_tma<forward>"), (1, "k_cost")):
        ctx.repeat(s, p, 1.0, kind, 3)
        ms, _ = rig.timed(lambda: ctx.repeat(s, p, 1.0, kind, 10))
        ms /= 10
        gbs = 24.0 * n / (ms * 1e-3) / 1e9
        res[name] = {"launch_ms": ms, "achieved": gbs, "frac": gbs / peak, "sample_visits_per_s": n / (ms * 1e-3)}
    k2 = res["k_normal_eq_tma<forward>"]
    out = {"bound": "hbm", "kernel": "k_normal_eq_tma<forward>", "samples": n, "achieved": k2["achieved"], "peak": peak,
           "unit": "GB/s", "frac": k2["frac"], "traffic": NCU.get("k_normal_eq_tma<forward>", {}).get("dram_bytes", 2.4048e9),
           "traffic_source": "ncu --set full, " + NCU.get("k_normal_eq_tma<forward>", {}).get("source", "profiles/r01_ncu_tables.md"),
           "fp64_pipe_pct": NCU.get("k_normal_eq_tma<forward>", {}).get("fp64_pipe_pct"),
           "algorithmic_bytes_per_launch": 24.0 * n, "launch_ms": k2["launch_ms"], "k_cost": res["k_cost"],
           "note": "inputs 2.4 GB >> 126 MB L2, 10 back-to-back launches after 3 warm-ups; BASELINE.json north_star's "
                   ">= 60 % of HBM target is stated for this regime"}
    ctx.fit_global(s, A.REF_GLOBAL)
    ms8, (ret8, p8, info8) = rig.timed(lambda: ctx.fit_global(s, A.REF_GLOBAL))
    st8 = ctx.fit_stats()
    sweeps8 = st8["jac_passes"] + st8["cost_passes"]
    gbs = 24.0 * n * sweeps8 / (ms8 * 1e-3) / 1e9
    out["complete_fit"] = {"samples": n, "ms": ms8, "iterations": float(info8[5]), "nfev": float(info8[7]), "sweeps": sweeps8,
                           "driver": "persistent kernel" if st8["ctas"] else "one kernel per evaluation (K2 TMA-staged, K3)",
                           "achieved": gbs, "unit": "GB/s", "frac": gbs / peak,
                           "sample_evals_per_s": float(info8[7]) * n / (ms8 * 1e-3), "p": [float(v) for v in p8]}
    s.free()
    return out


def secant_leg(rig, sizes=(10_000_000, 100_000_000)):
    """brdfgpu_dlevmar_dif (levmar's secant LM, lm_core.c:438-842) on resident samples: the n x 3 Jacobian lives in HBM."""
    A, ctx, peak = rig.A, rig.ctx, rig.peak
    out = []
    for n in sizes:
        s = ctx.synth(n, SEED)
        p0, opts = (0.5, 0.3, 8.0), (1e-3, 1e-15, 1e-15, 1e-20, 1e-6)
        ctx.fit_global_unc(s, p0, 60, opts)
        ms, (ret, p, info) = rig.timed(lambda: ctx.fit_global_unc(s, p0, 60, opts))
        st = ctx.fit_stats()
        njac, ncost, nupd = st["jac_passes"], st["cost_passes"], st["secant_updates"]
        # bytes the passes move per sample: difference rebuild 24 read + 32 written, trial cost 24 read, Broyden update
        # 56 read + 24 (32 when the step was accepted) written (secant_fit.cu header)
        moved = float(n) * (56.0 * njac + 24.0 * ncost + 80.0 * nupd + 8.0 * st["secant_updates_accepted"])
        gbs = moved / (ms * 1e-3) / 1e9
        out.append({"samples": n, "ms": ms, "iterations": int(info[5]), "nfev": int(info[7]), "stop_reason": int(info[6]),
                    "difference_rebuilds": njac, "cost_passes": ncost, "broyden_updates": nupd,
                    "bytes_moved": moved, "achieved": gbs, "unit": "GB/s", "frac": gbs / peak, "bound": "hbm", "p": [float(v) for v in p]})
        s.free()
    return out


def scale_hbm(rig):
    """BASELINE configs[4] at N GPUs: 10^7 and 10^8 samples IN TOTAL (strong scaling: n/N per GPU through the persistent
    kernel's TMA-streamed path + the in-kernel exchange) and 10^8 PER GPU (weak: K2/K3 per evaluation + NCCL all-reduce)."""
    A, ctx, world, rank, peak = rig.A, rig.ctx, rig.world, rig.rank, rig.peak
    rows = []
    for label, n_total in (("strong 1e7 total", 10_000_000), ("strong 1e8 total", 100_000_000), ("weak 1e8 per GPU", 100_000_000 * world)):
        lo, hi = rank * n_total // world, (rank + 1) * n_total // world
        s = ctx.synth(hi - lo, SEED, start=lo)
        ctx.fit_global(s, A.REF_GLOBAL)
        ms, (ret, p, info) = rig.timed(lambda: ctx.fit_global(s, A.REF_GLOBAL))
        st = ctx.fit_stats()
        sweeps = st["jac_passes"] + st["cost_passes"]
        per_gpu = 24.0 * (hi - lo) * sweeps / (ms * 1e-3) / 1e9
        blobs = rig.gather_bytes(fit_blob(ret, p, info))
        rows.append({"case": label, "samples_total": n_total, "samples_per_gpu": hi - lo, "ms_per_fit": ms, "iterations": int(info[5]),
                     "nfev": int(info[7]), "sweeps": sweeps, "trial_points": st["cost_points"], "us_per_sweep": 1e3 * ms / max(sweeps, 1),
                     "driver": "persistent kernel, %d%% of the shard on chip" % round(100.0 * min(1.0, st["resident_samples"] / max(hi - lo, 1)))
                     if st["ctas"] else "kernel per evaluation + NCCL all-reduce",
                     "sample_evals_per_s": float(info[7]) * n_total / (ms * 1e-3), "gbs_per_gpu": per_gpu, "frac_of_hbm_peak": per_gpu / peak,
                     "ranks_bit_identical": all(b == blobs[0] for b in blobs), "p": [float(v) for v in p], "stop_reason": int(info[6])})
        s.free()
    return {"peak_gbs": peak, "rows": rows,
            "note": "GB/s per GPU = 24 B x samples per GPU x physical sweeps / time; a projected-gradient sweep evaluates up to 8 trial "
                    "points per pass over the samples, so a sweep can be fp64-bound at a low HBM fraction (trial_points / sweeps says how many)"}


def bunny_sharded(rig):
    """BASELINE configs[2]: img/bunny through all 13 camera calibrations; views sharded over the ranks (rank r takes views
    r, r+N, ...: independent units, no collective), every rank's samples form ONE global fit per colour channel with the
    in-kernel exchange.  Checked against tests/golden/real_scenes.json."""
    import real_scenes as R
    A, ctx, world, rank = rig.A, rig.ctx, rig.world, rig.rank
    sc = R.load("bunny")
    gold_path = os.path.join(ROOT, "tests", "golden", "real_scenes.json")
    if sc is None or not os.path.exists(gold_path):
        return {"unavailable": "tests/_scenes/bunny.npz (decoded from the reference's img/bunny by tests/real_scenes.py) is not on this box"}
    gold = json.load(open(gold_path))["bunny"]
    scene = ctx.scene(sc["V"], sc["F"], sc["imgs"], dark=sc["dark"])
    mine = list(range(rank, len(sc["cams"]), world))
    cams = sc["cams"][mine] if mine else sc["cams"][:0]

    def gather_all(ch):
        if len(mine) == 0:
            return ctx.upload(np.zeros(0), np.zeros(0), np.zeros(0), A.BLINN_PHONG), 0
        s, _, nfit = scene.gather_resident(cams, model=A.BLINN_PHONG, channel=ch, want_global=True, want_batch=False)
        return s, nfit
    s, nfit = gather_all(0)
    s.free()
    t_gather, _ = rig.timed(lambda: gather_all(0)[0].free(), 3)
    fits = []
    ok = True
    t_fit = 0.0
    for ch in range(3):
        s, nfit = gather_all(ch)
        ctx.fit_global(s, A.REF_GLOBAL)
        ms, (ret, p, info) = rig.timed(lambda: ctx.fit_global(s, A.REF_GLOBAL))
        t_fit += ms
        want = gold["global"][ch]
        rec = {"channel": "BGR"[ch], "ret": int(ret), "stop_reason": int(info[6]), "iterations": int(info[5]), "ms": ms,
               "golden_stop_reason": int(want["info"][6]), "p": [float(v) for v in p]}
        # levmar's outcome: the same success / failure, the same parameters (1e-4) and cost (1e-6).  The stop reason has
        # to agree when the reference FAILED (7: non-finite residuals) or both converged; a fit that crawls along an active
        # bound may run into itmax on one summation order and meet the step test on another -- reported
        both_failed = want["ret"] < 0 and ret < 0
        good = (ret >= 0) == (want["ret"] >= 0) and (int(info[6]) == int(want["info"][6]) if both_failed else int(info[6]) in (1, 2, 3, 5, 6))
        wp = np.array(want["p"])
        rec["p_rel_err_vs_golden"] = float(np.max(np.abs(p - wp) / np.maximum(np.abs(wp), 1e-7)))
        good = good and rec["p_rel_err_vs_golden"] <= 1e-4
        if want["ret"] >= 0:
            rec["cost_rel_err_vs_golden"] = float(abs(info[1] - want["info"][1]) / want["info"][1])
            good = good and rec["cost_rel_err_vs_golden"] <= 1e-6
        else:
            good = good and int(info[5]) == int(want["info"][5])   # died on the same evaluation
        rec["pass"] = bool(good)
        ok = ok and good
        fits.append(rec)
        s.free()
    counts = rig.gather_bytes(int(nfit))
    scene.free()
    total = int(sum(counts))
    return {"views": len(sc["cams"]), "views_per_rank": [len(range(r, len(sc["cams"]), world)) for r in range(world)],
            "fits_total": total, "golden_fits_total": int(gold["total_fits"]), "samples_per_channel": total * 16,
            "gather_ms": t_gather, "face_views_per_s": len(sc["cams"]) * sc["F"].shape[0] / (t_gather * 1e-3),
            "global_fits_ms": t_fit, "fits": fits, "pass": bool(ok and total == int(gold["total_fits"]))}


# ---- the other two stages on one GPU -----------------------------------------------------------
def batched_leg(rig, cpu=True):
    """BASELINE configs[3]: 65 536 independent per-face fits x 64 samples, one launch.  Two kernels: the levmar-exact one
    (what the reference's drivers run: every fit EQUALS dlevmar_bc_dif's result) is the stage's number, the fast one
    (exp(n ln t), butterfly sums: converged fits within the parity bars) is reported beside it."""
    A, ctx, peak = rig.A, rig.ctx, rig.peak
    nfit, nper = 65536, 64
    b = ctx.batch_synth(nfit, nper, seed=2026)
    modes = {}
    for name, mode, kernel in (("levmar_exact", A.JAC_FD_EXACT, "k_batched_fit_exact<32,2>"), ("fast", A.JAC_FD, "k_batched_fit<32,2>")):
        for _ in range(2):
            b.fit(A.REF_PERFACE, jac_mode=mode)
        ms, _ = rig.timed(lambda: b.fit(A.REF_PERFACE, jac_mode=mode), 3)
        pp, info, ret = b.results()
        gbs = 24.0 * nper * nfit / (ms * 1e-3) / 1e9
        ncu = NCU.get(kernel, {})
        modes[name] = {"value": nfit / (ms * 1e-3), "unit": "fits/s", "ms_per_launch": ms, "seconds_per_fit": ms * 1e-3 / nfit,
                       "mean_iterations": float(info[:, 5].mean()), "mean_nfev": float(info[:, 7].mean()),
                       "sample_evals_per_s": float(info[:, 7].sum() * nper / (ms * 1e-3)),
                       "converged_fraction": float(np.isin(info[:, 6].astype(int), (1, 2, 6)).mean()),
                       "roofline": {"bound": "hbm", "kernel": kernel, "achieved": gbs, "peak": peak, "unit": "GB/s", "frac": gbs / peak,
                                    "algorithmic_bytes_per_launch": 24.0 * nper * nfit, "launch_ms": ms, "traffic": ncu.get("dram_bytes"),
                                    "binding_limit": "latency of per-fit control flow; fp64 pipe %s %% busy (ncu, %s)" % (ncu.get("fp64_pipe_pct", "?"), ncu.get("source", "?")),
                                    "fp64_pipe_pct": ncu.get("fp64_pipe_pct"),
                                    "note": "each fit reads its 1 536 B once and then iterates on chip: HBM cannot be the bound (SURVEY.md 8d)"}}
    b.free()
    out = {"metric": "batched BRDF fits/sec", "nfit": nfit, "samples_per_fit": nper, "preset": "REF_PERFACE"}
    out.update(modes["levmar_exact"])
    out["mode"] = "levmar-exact (BRDFGPU_JAC_FD_EXACT): p, info[0..9], ret of every fit equal the reference's bit for bit"
    out["fast"] = modes["fast"]
    out["fast"]["mode"] = "BRDFGPU_JAC_FD: 99.97 % of the converged fits within 1e-4 / 1e-6 of the reference (profiles/r02_parity.md)"
    if cpu:
        count = 3000
        dt, kind, nfev = cpu_batched_fits(count, nper)
        out["cpu_baseline"] = {"value": count / dt, "unit": "fits/s", "cores": 1, "kind": kind, "seconds_per_fit": dt / count,
                               "sample": "the first %d of the 65 536 configs[3] fits, one dlevmar_bc_dif call each (brdfdata.cpp:1119), %.1f s" % (count, dt),
                               "mean_nfev": nfev, "host_cores_available": os.cpu_count()}
        out["time_to_solution_ratio"] = (dt / count) / out["seconds_per_fit"]
        out["fast"]["time_to_solution_ratio"] = (dt / count) / out["fast"]["seconds_per_fit"]
    return out


def gather_leg(rig, cpu=True):
    """Stage 1 on the reference's scenes: img/cup (1 view, configs[0]) and img/bunny x 13 views (configs[2]); the CPU
    number is the oracle's gather (the reference's own cannot be built here) on the same arrays, one thread."""
    import oracle_lib as O
    import real_scenes as R
    import scene_lib as S
    A, ctx, peak = rig.A, rig.ctx, rig.peak
    out = []
    for name in ("cup", "bunny"):
        sc = R.load(name)
        synthetic = sc is None
        if synthetic:   # a mesh of the same size in front of one camera: same work, not the reference's data
            V, F = S.height_field(200, 97 if name == "cup" else 64, seed=5)
            imgs, dark = S.random_images(16, 800, 600, seed=6)
            cams = np.array([S.look_at_camera((40.0 * np.cos(a), 30.0 * np.sin(a), 250.0), (0.0, 0.0, 0.0)) for a in np.linspace(0, 3, 13)])
            sc = dict(V=V, F=F, imgs=imgs, dark=dark, cams=cams)
        cams = sc["cams"][:1] if name == "cup" else sc["cams"]
        scene = ctx.scene(sc["V"], sc["F"], sc["imgs"], dark=sc["dark"])

        def resident():
            s, _, nfit = scene.gather_resident(cams, model=A.BLINN_PHONG, channel=0, want_global=True, want_batch=False)
            s.free()
            return nfit
        resident()
        t0 = time.perf_counter()
        for _ in range(5):
            nfit = resident()
        wall_ms = (time.perf_counter() - t0) / 5 * 1e3
        dev_ms, _ = rig.timed(resident, 5)
        t0 = time.perf_counter()
        g = scene.gather(cams)
        host_ms = (time.perf_counter() - t0) * 1e3
        units = len(cams) * sc["F"].shape[0]
        algo = 930.0 * nfit   # ~0.93 KB per mapped face-view
        rec = {"scene": name + (" (synthetic stand-in: tests/_scenes absent)" if synthetic else ""), "views": len(cams), "faces": int(sc["F"].shape[0]),
               "fits": int(nfit), "samples_per_channel": int(nfit) * 16,
               "resident_call_ms_wall": wall_ms, "resident_call_ms_device": dev_ms, "host_arrays_call_ms_wall": host_ms,
               "face_views_per_s": units / (wall_ms * 1e-3),
               "roofline": {"bound": "hbm", "kernels": "k_project + compaction + k_gather_samples", "achieved": algo / (dev_ms * 1e-3) / 1e9,
                            "peak": peak, "unit": "GB/s", "frac": algo / (dev_ms * 1e-3) / 1e9 / peak, "algorithmic_bytes": algo,
                            "binding_limit": "fp64 divisions and square roots that bit-exactness fixes (2 normalisations per sample)"}}
        if cpu:
            clean = []
            for im in sc["imgs"]:
                w = im.copy()
                O.oracle().oracle_subtract_ambient(w.ctypes.data, sc["dark"].ctypes.data, w.size)
                clean.append(w)
            H, W = clean[0].shape[:2]
            led = S.led_table()
            t0 = time.perf_counter()
            tot = 0
            for cam in cams:
                tot += S.oracle_gather(sc["V"], sc["F"], cam, led, clean, W, H)["nfit"]
            dt = time.perf_counter() - t0
            rec["cpu_baseline"] = {"value": units / dt, "unit": "face-views/s", "cores": 1, "kind": "port", "seconds": dt,
                                   "sample": "oracle/gather_oracle.c on the same arrays, all %d view(s), %d fits" % (len(cams), tot)}
            rec["time_to_solution_ratio"] = dt / (wall_ms * 1e-3)
            rec["fits_equal_cpu"] = bool(tot == nfit)
        scene.free()
        out.append(rec)
    return out


def batched_sharded(rig, nfit_total):
    A, ctx, world, rank = rig.A, rig.ctx, rig.world, rig.rank
    lo, hi = rank * nfit_total // world, (rank + 1) * nfit_total // world
    b = ctx.batch_synth(hi - lo, 64, seed=2026, first_fit=lo)
    out = {"nfit_total": nfit_total, "samples_per_fit": 64}
    for name, mode in (("levmar_exact", A.JAC_FD_EXACT), ("fast", A.JAC_FD)):
        b.fit(A.REF_PERFACE, jac_mode=mode)
        ms, _ = rig.timed(lambda: b.fit(A.REF_PERFACE, jac_mode=mode), 3)
        out[name] = {"ms": ms, "fits_per_s": nfit_total / (ms * 1e-3)}
    b.free()
    return out


def run_gpu_arm(args):
    # stdout carries exactly ONE line (the JSON); libraries that write there (NCCL prints its version banner to
    # stdout when NCCL_DEBUG is set) are sent to stderr for the duration of the run
    sys.stdout.flush()
    json_fd = os.dup(1)
    os.dup2(2, 1)
    json_out = os.fdopen(json_fd, "w")

    rig = Rig(args)
    A, ctx, rank, world = rig.A, rig.ctx, rig.rank, rig.world
    extra = {}
    if world > 1:
        extra["parity"] = multi_gpu_parity(rig)     # fails the run when outside the bars

    clocks = ClockSampler(rig.local)
    if rank == 0:
        clocks.start()
    h = headline(rig, replicated=False)
    clk = clocks.stop() if rank == 0 else None
    s, ms, info, p, st = h["samples"], h["ms"], h["info"], h["p"], h["st"]
    n, n_total = N_PER_GPU, N_PER_GPU * world
    value = h["evals"] * n_total / (ms * 1e-3)

    # ---- dominant kernel of the step and its roofline ----
    peak, ncu = rig.peak, NCU.get("k_persistent_fit", {})
    swept = 24.0 * n * h["sweeps"]
    achieved = swept / (ms * 1e-3) / 1e9
    roofline = {"bound": "hbm", "kernel": "k_persistent_fit", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": ncu.get("dram_bytes"), "traffic_source": "ncu --set full, " + ncu.get("source", "?"),
                "peak_source": rig.peak_src, "algorithmic_bytes_per_launch": swept, "launch_ms": ms,
                "sweeps_per_launch": h["sweeps"], "levmar_counted_passes_per_launch": h["levmar_passes"],
                "binding_limit": "fp64 latency inside the sweeps (4 warps per scheduler, ptxas serialises the chains of a trip) + "
                                 "exchange/control latency between them (the shard is in shared memory: HBM is not the bound at this size)",
                "fp64_pipe_pct": ncu.get("fp64_pipe_pct"), "issue_slots_pct": ncu.get("issue_slots_pct"),
                "cycles": {"sweeps": st["cyc_sweep"], "exchange": st["cyc_exchange"], "total": st["cyc_total"]},
                "note": "achieved = 24 B x samples x PHYSICAL sweeps over the samples / kernel time: the rate at which the kernel walks its "
                        "(on-chip) samples, as a fraction of the HBM peak for scale.  One launch = one complete fit.  roofline_hbm is the "
                        "HBM-resident regime (10^8 samples) that BASELINE.json's 60 % target names."}

    e2e = e2e_single(rig, s) if world == 1 else e2e_multi(rig, s, h["drive"])
    extra["scaling_invariant"] = scripted_sweeps(rig, s)
    extra["scaling_invariant"]["headline_fit"] = {
        "iterations": int(info[5]), "nfev": int(info[7]), "sweeps_per_fit": h["sweeps"], "us_per_sweep": 1e3 * ms / h["sweeps"],
        "ranks_bit_identical": h["ranks_bit_identical"],
        "note": "the fit itself is a different problem for every N (N x 10^6 distinct samples): its iteration count, and with it "
                "the mix of Jacobian / trial / walk sweeps, changes with N, so value(N) / value(1) mixes exchange cost with trajectory"}
    if world > 1:
        s.free()
        if not args.quick:
            extra["scale_hbm"] = scale_hbm(rig)
            extra["bunny"] = bunny_sharded(rig)
            extra["batched"] = {"metric": "batched BRDF fits/sec", "preset": "REF_PERFACE", "n_gpus": world,
                                "strong_65536_total": batched_sharded(rig, 65536), "weak_65536_per_gpu": batched_sharded(rig, 65536 * world)}
            if not (extra["bunny"].get("pass", True)):
                if rank == 0:
                    sys.stderr.write("bench.py: bunny sharded fit outside the parity bars: %s\n" % json.dumps(extra["bunny"]))
                raise SystemExit(3)
    else:
        if not args.quick:
            extra["roofline_hbm"] = hbm_kernels(rig)
            extra["stages"] = {"batched": batched_leg(rig, cpu=not args.no_cpu), "gather": gather_leg(rig, cpu=not args.no_cpu),
                               "secant": secant_leg(rig)}
            extra["batched"] = extra["stages"]["batched"]   # (round-1 key)
            extra["bunny"] = bunny_sharded(rig)

    cpu = None
    top = {"seconds_per_fit": ms * 1e-3, "seconds_per_fit_e2e": e2e["seconds_per_fit"]}
    if rank == 0 and world == 1 and not args.no_cpu:
        dt, nfev, kind, p_cpu, info_cpu = cpu_reference_fit(n)
        cpu = {"value": nfev * n / dt, "unit": UNIT, "cores": 1, "kind": kind,
               "sample": "the full step once: n=%d, REF_GLOBAL, %d iterations, %d evaluations, %.1f s" % (n, int(info_cpu[5]), int(nfev), dt),
               "host_cores_available": os.cpu_count(), "seconds_per_fit": dt,
               "parity": {"p_rel_err_max": float(np.max(np.abs(p - p_cpu) / np.abs(p_cpu))),
                          "cost_rel_err": float(abs(info[1] - info_cpu[1]) / info_cpu[1]),
                          "p_gpu": [float(v) for v in p], "p_cpu": [float(v) for v in p_cpu]}}
        top["seconds_per_fit_cpu"] = dt
        top["time_to_solution"] = {"resident": dt / (ms * 1e-3), "e2e": dt / e2e["seconds_per_fit"],
                                   "note": "CPU seconds per fit / GPU seconds per fit on the same problem to the same answer: the like-for-like "
                                           "ratio.  The metric's ratio is larger by nfev_gpu / nfev_cpu = %.2f because the two trajectories "
                                           "count different numbers of evaluations (summation order, SURVEY.md Q13)." % (h["evals"] / nfev)}

    if rank == 0:
        out = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(3, args.warmup),
               "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
               "data": "synthetic",
               "config": {"workload": WORKLOAD, "samples_per_gpu": n, "samples_total": n_total, "preset": "REF_GLOBAL", "model": "blinn-phong",
                          "jacobian": "forward differences, delta=1 (levmar-exact)",
                          "driver": ("persistent cooperative kernel" + ("" if world == 1 else ", fused peer-memory all-reduce of the sums per evaluation"))
                          if h["drive"] == A.DRIVE_PERSISTENT else "host loop + NCCL all-reduce(10 f64) per evaluation",
                          "l2": "inputs (24 MB/GPU) are smaller than L2 by definition of the workload and live in shared memory for the whole fit; "
                                "roofline_hbm uses 2.4 GB inputs",
                          "iterations_per_fit": float(info[5]), "nfev_per_fit": float(info[7]), "stop_reason": int(info[6]),
                          "sweeps_per_fit": h["sweeps"], "fit_stats": st},
               "clocks": clk, "e2e": e2e, "gpu_launches": int(h["launches"]), "roofline": roofline, "cpu_baseline": cpu}
        out.update(top)
        out.update(extra)
        json_out.write(json.dumps(out) + "\n")
        json_out.flush()
    if world == 1:
        s.free()
    if rig.dist is not None:
        rig.dist.barrier()
        rig.dist.destroy_process_group()
    ctx.close()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--quick", action="store_true", help="headline only: skip the HBM-regime, stage and multi-GPU side measurements")
    ap.add_argument("--no-cpu", action="store_true", help="skip the CPU baseline legs")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_gpu_arm(args)


if __name__ == "__main__":
    main()
