"""Tables of profiles/r02_scale.md from the committed bench lines (profiles/r02_bench_n{1,2,4,8}.json).
    python profiles/make_scale_md.py > /tmp/tables.md
"""
import json
import os

HERE = os.path.dirname(os.path.abspath(__file__))
D = {n: json.loads(open(os.path.join(HERE, "r02_bench_n%d.json" % n)).read().strip().splitlines()[-1]) for n in (1, 2, 4, 8)}


def g(d, k):
    return d.get(k) if d.get(k) is not None else d["config"].get(k)


print("## 1. Driver-visible parity (`parity`)\n")
print("| GPUs | preset | ranks bit-identical | max rel. parameter error vs 1 GPU | vs CPU oracle | cost error vs 1 GPU | vs oracle | stop reason (GPU / oracle) |")
print("|---|---|---|---|---|---|---|---|")
for n in (2, 4, 8):
    for name, r in g(D[n], "parity")["presets"].items():
        print("| %d | %s | %s | %.1e | %.1e | %.1e | %.1e | %d / %d |" % (
            n, name, r["ranks_bit_identical"], r["p_rel_err_vs_1gpu"], r["p_rel_err_vs_oracle"], r["cost_rel_err_vs_1gpu"],
            r["cost_rel_err_vs_oracle"], r["stop_reason"], r["oracle_stop_reason"]))

print("\n## 2. Global fit, 10^6 samples per GPU (weak; the headline `value`) and the trajectory-invariant figure\n")
print("| GPUs | samples | iterations | nfev | sweeps | ms per fit | sample-evals/s (`value`) | value / (N x value_1) | us per sweep (fit) | **us per sweep, scripted sequence** | exchange cycles per sweep (phases: warp sums, CTA sum + publish, collect, finish + peer hop) | e2e ms |")
print("|---|---|---|---|---|---|---|---|---|---|---|---|")
v1 = D[1]["value"]
for n in (1, 2, 4, 8):
    d = D[n]
    si = g(d, "scaling_invariant")
    hf = si["headline_fit"]
    cyc = si["cycles_per_sweep_cta0"]
    print("| %d | %.0e | %d | %d | %d | %.3f | %.3g | %.2f | %.2f | **%.2f** | %d (%s) | %.3f |" % (
        n, n * 1e6, hf["iterations"], hf["nfev"], hf["sweeps_per_fit"], d["ms_per_step"], d["value"], d["value"] / (n * v1),
        hf["us_per_sweep"], si["us_per_sweep"], cyc["exchange"], ", ".join("%d" % x for x in cyc["exchange_phases"]),
        d["e2e"]["ms_per_step"]))

print("\n## 3. BASELINE configs[4] at N GPUs (`scale_hbm`): 10^7 and 10^8 samples in total (strong), 10^8 per GPU (weak)\n")
print("| case | GPUs | samples per GPU | driver | ms per fit | iterations | sweeps (trial points) | us per sweep | GB/s per GPU (24 B x samples x sweeps) | of %d |" % g(D[2], "scale_hbm")["peak_gbs"])
print("|---|---|---|---|---|---|---|---|---|---|")
hb = g(D[1], "roofline_hbm") or D[1].get("roofline_hbm")
for case in ("strong 1e7 total", "strong 1e8 total", "weak 1e8 per GPU"):
    for n in (2, 4, 8):
        for r in g(D[n], "scale_hbm")["rows"]:
            if r["case"] == case:
                print("| %s | %d | %d | %s | %.2f | %d | %d (%d) | %.1f | %.0f | %.2f |" % (
                    case, n, r["samples_per_gpu"], r["driver"], r["ms_per_fit"], r["iterations"], r["sweeps"], r["trial_points"],
                    r["us_per_sweep"], r["gbs_per_gpu"], r["frac_of_hbm_peak"]))

print("\n## 4. BASELINE configs[2] (`bunny`): img/bunny through all 13 calibrations, gather sharded by view, one global fit per channel\n")
print("| GPUs | views per rank | fits (golden %d) | gather ms (max over ranks) | channel B: iterations, stop, ms | G | R | golden check |" % g(D[1], "bunny")["golden_fits_total"])
print("|---|---|---|---|---|---|---|---|")
for n in (1, 2, 4, 8):
    b = g(D[n], "bunny")
    f = {x["channel"]: x for x in b["fits"]}
    print("| %d | %s | %d | %.3f | %d, %d, %.1f | %d, %d, %.2f | %d, %d, %.2f | %s |" % (
        n, b["views_per_rank"], b["fits_total"], b["gather_ms"], f["B"]["iterations"], f["B"]["stop_reason"], f["B"]["ms"],
        f["G"]["iterations"], f["G"]["stop_reason"], f["G"]["ms"], f["R"]["iterations"], f["R"]["stop_reason"], f["R"]["ms"],
        "pass" if b["pass"] else "FAIL"))

print("\n## 5. BASELINE configs[3] (`batched`): 65 536 fits x 64 samples, sharded by fit id, no communication\n")
print("| GPUs | strong (65 536 in total): levmar-exact ms, fits/s | efficiency | fast ms, fits/s | efficiency | weak (65 536 per GPU): levmar-exact fits/s | efficiency | fast fits/s | efficiency |")
print("|---|---|---|---|---|---|---|---|---|")
b1 = g(D[1], "batched")
e1 = b1["value"]           # the N = 1 line reports the levmar-exact kernel as the stage value ...
f1 = b1["fast"]["value"]   # ... and the fast kernel next to it
e1ms = 65536 / e1 * 1e3
f1ms = 65536 / f1 * 1e3
print("| 1 | %.2f, %.3g | 1.00 | %.2f, %.3g | 1.00 | %.3g | 1.00 | %.3g | 1.00 |" % (e1ms, e1, f1ms, f1, e1, f1))
for n in (2, 4, 8):
    b = g(D[n], "batched")
    s, w = b["strong_65536_total"], b["weak_65536_per_gpu"]
    print("| %d | %.2f, %.3g | %.2f | %.2f, %.3g | %.2f | %.3g | %.2f | %.3g | %.2f |" % (
        n, s["levmar_exact"]["ms"], s["levmar_exact"]["fits_per_s"], s["levmar_exact"]["fits_per_s"] / (n * e1),
        s["fast"]["ms"], s["fast"]["fits_per_s"], s["fast"]["fits_per_s"] / (n * f1),
        w["levmar_exact"]["fits_per_s"], w["levmar_exact"]["fits_per_s"] / (n * e1), w["fast"]["fits_per_s"], w["fast"]["fits_per_s"] / (n * f1)))
