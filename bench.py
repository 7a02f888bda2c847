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
    roofline  the dominant kernel of the step (the persistent fit kernel; at N>1 the K2 pass):
            algorithmic bytes = 24 B x samples x passes over the samples, / its CUDA-event duration
    roofline_hbm  K2 (fused residual + difference Jacobian + J^T J) alone at 10^8 samples, inputs
            far larger than L2 -- the regime BASELINE.json's 60 %-of-HBM target names
    cpu_baseline  the reference's single-threaded levmar timed on this box's host cores on the same
            workload
Multi-GPU (weak scaling): every rank holds its own 10^6-sample shard of ONE fit over N x 10^6
samples; each evaluation all-reduces 11 doubles.
"""
import argparse
import ctypes as C
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
# reference arm / cpu baseline: the reference's levmar on the host
# ----------------------------------------------------------------------------------------------
def cpu_reference_fit(n, itmax=None):
    """One dlevmar_bc_dif BRDF fit on the CPU.  Returns (seconds, nfev, kind)."""
    import oracle_lib as O
    import synth
    c, td, th, x = synth.samples(n, seed=SEED)
    ref = O.ref()
    lib, prefix, kind = (ref, "", "reference") if ref is not None else (O.oracle(), "oracle_", "port")
    preset = dict(O.REF_GLOBAL)
    if itmax is not None:
        preset["itmax"] = itmax
    O.oracle()
    t0 = time.perf_counter()
    ret, p, info = O.brdf_fit(lib, prefix, c, td, th, x, 1, preset)
    dt = time.perf_counter() - t0
    return dt, float(info[7]), kind, p, info


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
           "warmup": args.warmup, "ms_per_step": 1e3 * total_t / fits, "higher_is_better": True, "scaling": "weak",
           "vs_baseline": None, "dtype": "f64", "data": "synthetic", "timed_fits": fits,
           "config": {"workload": "synthetic single global BRDF fit, 10^6 samples, fp64 (BASELINE configs[1])",
                      "samples_per_gpu": n, "preset": "REF_GLOBAL", "model": "blinn-phong"},
           "cpu_baseline": {"value": value, "unit": UNIT, "cores": 1, "kind": kind, "sample": sample,
                            "host_cores_available": os.cpu_count()},
           "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
           "gpu_launches": 0}
    print(json.dumps(out))


# ----------------------------------------------------------------------------------------------
# GPU arm
# ----------------------------------------------------------------------------------------------
def run_gpu_arm(args):
    # stdout carries exactly ONE line (the JSON); libraries that write there (NCCL prints its version banner to
    # stdout when NCCL_DEBUG is set) are sent to stderr for the duration of the run
    sys.stdout.flush()
    json_fd = os.dup(1)
    os.dup2(2, 1)
    json_out = os.fdopen(json_fd, "w")

    import torch

    from brdf_b200 import api as A

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world > 1:
        raise SystemExit("--gpus %d but WORLD_SIZE=%d" % (args.gpus, world))
    torch.cuda.set_device(local)
    ctx = A.Context(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        ids = [A.comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(ids, src=0)
        ctx.comm_init(ids[0], rank, world)
        # peer buffers for the fused in-kernel exchange (CUDA IPC handles travel through torch.distributed)
        handles = [None] * world
        dist.all_gather_object(handles, ctx.peer_export())
        ctx.peer_attach(handles, rank, world)

    def barrier():
        ctx.synchronize()
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    stream = torch.cuda.ExternalStream(ctx.stream)
    n = N_PER_GPU
    n_total = n * world
    s = ctx.synth(n, SEED, start=rank * n)
    # BRDF_BENCH_DRIVE=host: one kernel + one NCCL all-reduce per evaluation instead of the persistent kernel
    drive = A.DRIVE_HOST if os.environ.get("BRDF_BENCH_DRIVE") == "host" else A.DRIVE_PERSISTENT

    def fit():
        return ctx.fit_global(s, A.REF_GLOBAL, drive=drive)

    for _ in range(max(3, args.warmup)):
        ret, p, info = fit()
    clocks = ClockSampler(local)
    barrier()
    if rank == 0:
        clocks.start()
    launches0 = ctx.launches
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record(stream)
    evals = passes = ref_passes = 0.0
    for _ in range(args.steps):
        ret, p, info = fit()
        evals += info[7]
        st = ctx.fit_stats()
        passes += st["jac_passes"] + st["cost_passes"]   # sweeps over the samples (a batched PG sweep counts once)
        # passes over the samples the REFERENCE algorithm makes for the same trajectory: one per Jacobian (fused
        # here; levmar makes m+1) and one per counted cost evaluation (speculative trial points are not counted)
        ref_passes += info[8] + (info[7] - 4.0 * info[8])
    ev1.record(stream)
    barrier()
    launches = ctx.launches - launches0
    ms = ev0.elapsed_time(ev1)
    clk = clocks.stop() if rank == 0 else None
    if dist is not None:
        t = torch.tensor([ms], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    value = evals * n_total / (ms * 1e-3)

    # ---- dominant kernel of the step and its roofline ----
    peak, peak_src = peaks()
    if drive == A.DRIVE_PERSISTENT:
        # the step IS one persistent kernel launch per rank (+ an 800-byte result copy); at N>1 the
        # cross-GPU exchange of the sums happens inside it (peer stores over NVLink)
        kernel, launches_per_step, kern_ms = "k_persistent_fit", 1, ms / args.steps
        algo_bytes = 24.0 * n * ref_passes / args.steps
    else:
        kernel = "k_normal_eq<forward>"
        reps = 50
        ctx.repeat(s, p, 1.0, 0, 5)
        ctx.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        ctx.repeat(s, p, 1.0, 0, reps)
        e1.record(stream)
        ctx.synchronize()
        kern_ms, launches_per_step, algo_bytes = e0.elapsed_time(e1) / reps, None, 24.0 * n
    achieved = algo_bytes / (kern_ms * 1e-3) / 1e9
    roofline = {"bound": "hbm", "kernel": kernel, "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": 40.98e6 if drive == A.DRIVE_PERSISTENT else None,
                "traffic_source": "ncu --set full, profiles/r01_ncu_tables.md r01_persist_spec (dram read + write of one k_persistent_fit launch)",
                "peak_source": peak_src, "algorithmic_bytes_per_launch": algo_bytes, "launch_ms": kern_ms,
                "sweeps_per_launch": passes / args.steps, "evaluations_per_launch": ref_passes / args.steps,
                "swept_bytes_per_launch": 24.0 * n * passes / args.steps,
                "note": "algorithmic bytes = 24 B/sample x the evaluations levmar counts for this trajectory (1 per fused "
                        "Jacobian, 1 per cost evaluation; SURVEY.md 8d).  At 10^6 samples/GPU the shard is held in shared "
                        "memory for the whole fit and the projected-gradient walk evaluates up to 8 trial points per sweep, "
                        "so almost none of these bytes move through HBM (ncu: 41 MB of DRAM traffic per launch) and "
                        "'achieved' is an algorithmic rate, not HBM utilisation; the kernel is bound by the FP64 pipe and "
                        "the per-evaluation exchange latency.  roofline_hbm is the HBM-resident regime (10^8 samples).  "
                        "Speculative Jacobians (fit_stats.spec_jac_*) answer some cost evaluations with a Jacobian sweep "
                        "whose sums the next iteration reuses; evaluations are counted as levmar counts them either way."}

    extra = {}
    if rank == 0 and world == 1 and not args.quick:
        extra = side_measurements(ctx, torch, stream, A, peak)
    if world > 1 and not args.quick:
        # batched mode on N GPUs (BASELINE configs[3]): independent fits, sharded by fit id, no communication
        def batched_sharded(nfit_total):
            lo, hi = rank * nfit_total // world, (rank + 1) * nfit_total // world
            b = ctx.batch_synth(hi - lo, 64, seed=2026, first_fit=lo)
            b.fit(A.REF_PERFACE)
            barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            for _ in range(3):
                b.fit(A.REF_PERFACE)
            e1.record(stream)
            barrier()
            t = torch.tensor([e0.elapsed_time(e1) / 3], device="cuda", dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            b.free()
            return {"nfit_total": nfit_total, "samples_per_fit": 64, "ms": float(t.item()), "fits_per_s": nfit_total / (float(t.item()) * 1e-3)}
        extra["batched"] = {"metric": "batched BRDF fits/sec", "preset": "REF_PERFACE", "n_gpus": world,
                            "strong_65536_total": batched_sharded(65536), "weak_65536_per_gpu": batched_sharded(65536 * world)}

    # ---- e2e: the levmar-signature call with pinned host buffers ----
    e2e = None
    if world == 1:
        c_h, t_h, x_h = s.download()
        angles = torch.empty(3 * n, dtype=torch.float64).pin_memory()
        xs = torch.empty(n, dtype=torch.float64).pin_memory()
        angles[:n] = torch.from_numpy(c_h); angles[n:2 * n] = torch.from_numpy(t_h); angles[2 * n:] = 0.0
        xs[:] = torch.from_numpy(x_h)
        extra_data = A.ExtraData(C.cast(angles.data_ptr(), A.dptr), 1)
        g = A.REF_GLOBAL
        lb, ub, opts = (np.array(g[k], dtype=np.float64) for k in ("lb", "ub", "opts"))
        fn = A.func_address("brdfgpu_BRDFFunc")

        def e2e_call():
            pp = np.array(g["p0"], dtype=np.float64)
            inf = np.zeros(10)
            r = A.lib().brdfgpu_dlevmar_bc_dif(fn, A._d(pp), C.cast(xs.data_ptr(), A.dptr), 3, n, A._d(lb), A._d(ub), None,
                                               g["itmax"], A._d(opts), A._d(inf), None, None,
                                               C.cast(C.pointer(extra_data), C.c_void_p))
            return r, pp, inf
        for _ in range(2):
            e2e_call()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        ev = 0.0
        for _ in range(args.steps):
            r, pp, inf = e2e_call()
            ev += inf[7]
        dt = time.perf_counter() - t0
        e2e = {"value": ev * n / dt, "unit": UNIT, "h2d_bytes_per_step": 3 * 8 * n, "d2h_bytes_per_step": 8 * (3 + 10) + 4,
               "ms_per_step": 1e3 * dt / args.steps, "call": "brdfgpu_dlevmar_bc_dif(brdfgpu_BRDFFunc, ...) with pinned host buffers",
               "p": [float(v) for v in pp]}
    else:
        # multi-GPU: the public call on resident shards; upload of the shard + fit + result read-back
        c_h, t_h, x_h = s.download()
        pin = [torch.from_numpy(a).pin_memory() for a in (c_h, t_h, x_h)]
        barrier()
        t0 = time.perf_counter()
        ev = 0.0
        h = C.c_void_p()
        ctx._ok(A.lib().brdfgpu_samples_upload(ctx.handle, n, C.cast(pin[0].data_ptr(), A.dptr), C.cast(pin[1].data_ptr(), A.dptr),
                                               C.cast(pin[2].data_ptr(), A.dptr), 1, C.byref(h)))
        ss = A.Samples(ctx, h)
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            # every step: this rank's shard host -> device, the fit over all ranks' shards, results back
            ctx._ok(A.lib().brdfgpu_samples_reload(ctx.handle, ss.handle, C.cast(pin[0].data_ptr(), A.dptr),
                                                   C.cast(pin[1].data_ptr(), A.dptr), C.cast(pin[2].data_ptr(), A.dptr)))
            r, pp, inf = ctx.fit_global(ss, A.REF_GLOBAL, drive=drive)
            ev += inf[7]
        ss.free()
        barrier()
        dt = torch.tensor([time.perf_counter() - t0], device="cuda", dtype=torch.float64)
        dist.all_reduce(dt, op=dist.ReduceOp.MAX)
        dt = float(dt.item())
        e2e = {"value": ev * n_total / dt, "unit": UNIT, "h2d_bytes_per_step": 3 * 8 * n, "d2h_bytes_per_step": 8 * 13 + 4,
               "ms_per_step": 1e3 * dt / args.steps, "call": "brdfgpu_samples_reload (pinned host shard -> device) + brdfgpu_fit_global per rank"}

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        dt, nfev, kind, p_cpu, info_cpu = cpu_reference_fit(n)
        cpu = {"value": nfev * n / dt, "unit": UNIT, "cores": 1, "kind": kind,
               "sample": "the full step once: n=%d, REF_GLOBAL, %d iterations, %.1f s" % (n, int(info_cpu[5]), dt),
               "host_cores_available": os.cpu_count(), "seconds_per_fit": dt,
               "time_to_solution_ratio": {"resident": dt / (ms / args.steps * 1e-3), "e2e": dt / (e2e["ms_per_step"] * 1e-3),
                                          "note": "same problem, same answer; the two trajectories differ in evaluation count "
                                                  "(summation order, SURVEY.md Q13), so seconds per fit is the like-for-like ratio"},
               "parity": {"p_rel_err_max": float(np.max(np.abs(p - p_cpu) / np.abs(p_cpu))),
                          "cost_rel_err": float(abs(info[1] - info_cpu[1]) / info_cpu[1]),
                          "p_gpu": [float(v) for v in p], "p_cpu": [float(v) for v in p_cpu]}}

    if rank == 0:
        out = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(3, args.warmup),
               "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
               "data": "synthetic",
               "config": {"workload": "synthetic single global BRDF fit, 10^6 samples, fp64 (BASELINE configs[1])",
                          "samples_per_gpu": n, "samples_total": n_total, "preset": "REF_GLOBAL", "model": "blinn-phong",
                          "jacobian": "forward differences, delta=1 (levmar-exact)",
                          "driver": ("persistent cooperative kernel" + ("" if world == 1 else ", fused peer-memory all-reduce of the 10 sums per evaluation"))
                          if drive == A.DRIVE_PERSISTENT else "host loop + NCCL all-reduce(10 f64) per evaluation",
                          "l2": "inputs (24 MB/GPU) are smaller than L2 by definition of the workload; roofline_hbm uses 2.4 GB inputs",
                          "iterations_per_fit": float(info[5]), "nfev_per_fit": float(info[7]), "stop_reason": int(info[6]),
                          "sweeps_per_fit": passes / args.steps, "fit_stats": st},
               "clocks": clk, "e2e": e2e, "gpu_launches": int(launches), "roofline": roofline, "cpu_baseline": cpu}
        out.update(extra)
        json_out.write(json.dumps(out) + "\n")
        json_out.flush()
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()
    ctx.close()


def side_measurements(ctx, torch, stream, A, peak):
    """HBM-resident roofline of K2/K3 at 10^8 samples and the batched-mode throughput (config 4)."""
    out = {}

    def timed(fn, reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        fn(reps)
        e1.record(stream)
        ctx.synchronize()
        return e0.elapsed_time(e1) / reps

    n = 100_000_000
    s = ctx.synth(n, SEED)
    p = [0.6, 0.35, 12.0]
    res = {}
    for kind, name in ((0, "k_normal_eq<forward>"), (1, "k_cost")):
        ctx.repeat(s, p, 1.0, kind, 3)
        ctx.synchronize()
        ms = timed(lambda r: ctx.repeat(s, p, 1.0, kind, r), 10)
        gbs = 24.0 * n / (ms * 1e-3) / 1e9
        res[name] = {"launch_ms": ms, "achieved": gbs, "frac": gbs / peak, "sample_visits_per_s": n / (ms * 1e-3)}
    k2 = res["k_normal_eq<forward>"]
    out["roofline_hbm"] = {"bound": "hbm", "kernel": "k_normal_eq<forward>", "samples": n, "achieved": k2["achieved"], "peak": peak,
                           "unit": "GB/s", "frac": k2["frac"], "traffic": 2.4048e9,
                           "traffic_source": "ncu --set full, profiles/r01_ncu_tables.md", "algorithmic_bytes_per_launch": 24.0 * n,
                           "launch_ms": k2["launch_ms"], "k_cost": res["k_cost"],
                           "note": "inputs 2.4 GB >> 126 MB L2, 10 back-to-back launches after 3 warm-ups"}
    # a COMPLETE fit at 10^8 samples (every Jacobian and trial evaluation streams 2.4 GB from HBM)
    ctx.fit_global(s, A.REF_GLOBAL)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    ret8, p8, info8 = ctx.fit_global(s, A.REF_GLOBAL)
    e1.record(stream)
    ctx.synchronize()
    st8 = ctx.fit_stats()
    ms8 = e0.elapsed_time(e1)
    sweeps8 = st8["jac_passes"] + st8["cost_passes"]
    out["roofline_hbm"]["complete_fit"] = {
        "samples": n, "ms": ms8, "iterations": float(info8[5]), "nfev": float(info8[7]), "sweeps": sweeps8,
        "driver": "persistent kernel" if st8["ctas"] else "one kernel per evaluation (K2 TMA-staged, K3)",
        "achieved": 24.0 * n * sweeps8 / (ms8 * 1e-3) / 1e9, "unit": "GB/s", "frac": 24.0 * n * sweeps8 / (ms8 * 1e-3) / 1e9 / peak,
        "sample_evals_per_s": float(info8[7]) * n / (ms8 * 1e-3), "p": [float(v) for v in p8]}
    del s
    # batched mode, BASELINE configs[3]: 65,536 fits x 64 samples
    nfit, nper = 65536, 64
    b = ctx.batch_synth(nfit, nper, seed=2026)
    for _ in range(2):
        b.fit(A.REF_PERFACE)
    ctx.synchronize()
    ms = timed(lambda r: [b.fit(A.REF_PERFACE) for _ in range(r)], 3)
    pp, info, ret = b.results()
    out["batched"] = {"metric": "batched BRDF fits/sec", "value": nfit / (ms * 1e-3), "unit": "fits/s", "nfit": nfit,
                      "samples_per_fit": nper, "ms_per_launch": ms, "preset": "REF_PERFACE",
                      "mean_iterations": float(info[:, 5].mean()), "mean_nfev": float(info[:, 7].mean()),
                      "sample_evals_per_s": float(info[:, 7].sum() * nper / (ms * 1e-3)),
                      "converged_fraction": float(np.isin(info[:, 6].astype(int), (1, 2, 6)).mean())}
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--quick", action="store_true", help="skip the 10^8-sample and batched side measurements")
    ap.add_argument("--no-cpu", action="store_true", help="skip the CPU baseline leg")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_gpu_arm(args)


if __name__ == "__main__":
    main()
