"""Condensed view of a bench.py JSON line (multi-GPU blocks included).   python profiles/show_bench.py file.json ..."""
import json
import sys

for path in sys.argv[1:]:
    d = json.load(open(path))
    print("== %s: N=%d value %.4g %s, %.3f ms/fit (%d it, %d nfev, %.0f sweeps), e2e %.3f ms" % (
        path, d["n_gpus"], d["value"], d["unit"], d["ms_per_step"], d["config"]["iterations_per_fit"], d["config"]["nfev_per_fit"],
        d["config"]["sweeps_per_fit"], d["e2e"]["ms_per_step"]))
    si = d.get("scaling_invariant")
    if si:
        c = si["cycles_per_sweep_cta0"]
        print("   scripted sweeps: %.2f us/sweep (cycles: sweep %.0f, exchange %.0f, phases %s); fit itself %.2f us/sweep" % (
            si["us_per_sweep"], c["sweep"], c["exchange"], [int(v) for v in c["exchange_phases"]], si["headline_fit"]["us_per_sweep"]))
    if "parity" in d:
        for k, v in d["parity"]["presets"].items():
            print("   parity %s: pass=%s ranks identical=%s |p-p1|=%.1e |p-oracle|=%.1e cost %.1e/%.1e" % (
                k, v.get("pass"), v["ranks_bit_identical"], v.get("p_rel_err_vs_1gpu", -1), v.get("p_rel_err_vs_oracle", -1),
                v.get("cost_rel_err_vs_1gpu", -1), v.get("cost_rel_err_vs_oracle", -1)))
    if "scale_hbm" in d:
        for r in d["scale_hbm"]["rows"]:
            print("   %-18s %9d/GPU %8.2f ms  %4d it %5d sweeps (%4d pts) %7.1f us/sweep  %6.0f GB/s/GPU = %.2f  %s" % (
                r["case"], r["samples_per_gpu"], r["ms_per_fit"], r["iterations"], r["sweeps"], r["trial_points"], r["us_per_sweep"],
                r["gbs_per_gpu"], r["frac_of_hbm_peak"], r["driver"][:34]))
    if "bunny" in d and "fits" in d["bunny"]:
        b = d["bunny"]
        print("   bunny: views/rank %s, %d fits, gather %.3f ms, fits %.1f ms, pass=%s; %s" % (
            b["views_per_rank"], b["fits_total"], b["gather_ms"], b["global_fits_ms"], b["pass"],
            [(f["channel"], f["iterations"], f["stop_reason"], "%.1f ms" % f["ms"]) for f in b["fits"]]))
    bt = d.get("batched")
    if bt and "strong_65536_total" in bt:
        for k in ("strong_65536_total", "weak_65536_per_gpu"):
            print("   batched %s: exact %.2f ms = %.3g fits/s | fast %.2f ms = %.3g fits/s" % (
                k, bt[k]["levmar_exact"]["ms"], bt[k]["levmar_exact"]["fits_per_s"], bt[k]["fast"]["ms"], bt[k]["fast"]["fits_per_s"]))
    elif bt:
        print("   batched: exact %.2f ms = %.3g fits/s (x%.0f CPU) | fast %.2f ms = %.3g fits/s" % (
            bt["ms_per_launch"], bt["value"], bt.get("time_to_solution_ratio", 0), bt["fast"]["ms_per_launch"], bt["fast"]["value"]))
