"""Per-fit parity classes of a batch of fits against the stored reference results (tests/golden/*_full_*.npz,
made by tests/golden/make_batched_full.py from the reference's own levmar)."""
import json
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
PAR_RTOL, COST_RTOL = 1e-4, 1e-6     # BASELINE.json north_star: parameters 1e-4 relative, final cost 1e-6 relative
CONVERGED = (1, 2, 6)                # info[6]: small gradient / small step / small residual
REASONS = {1: "small gradient", 2: "small step", 3: "itmax", 4: "singular", 5: "no further reduction", 6: "small residual", 7: "NaN/Inf"}


def load_full(name):
    z = np.load(os.path.join(HERE, "golden", name))
    return {k: z[k] for k in z.files}


def histogram(p, info, ret, ref):
    """Per reference stop reason: how many fits, and how many of them agree in each sense."""
    want_p = ref["p"].astype(np.float64)
    # the reference parameters are stored as float32 (6e-8 relative, far below the 1e-4 gate)
    p_ok = np.all(np.abs(p - want_p) <= PAR_RTOL * np.abs(want_p) + 1e-7, axis=1)
    cost_ok = np.abs(info[:, 1] - ref["cost"]) <= COST_RTOL * np.abs(ref["cost"]) + 1e-18
    cost_not_worse = info[:, 1] <= ref["cost"] * (1 + 1e-3) + 1e-15
    reason = info[:, 6].astype(int)
    same_reason = reason == ref["reason"]
    same_ret = (ret >= 0) == (ref["ret"] >= 0)
    out = {}
    for r in sorted(set(ref["reason"].tolist())):
        m = ref["reason"] == r
        out[int(r)] = dict(name=REASONS.get(int(r), "?"), fits=int(m.sum()), same_stop_reason=int((same_reason & m).sum()),
                           p_within_1e4=int((p_ok & m).sum()), cost_within_1e6=int((cost_ok & m).sum()),
                           strict=int((p_ok & cost_ok & m).sum()), cost_not_worse_1e3=int((cost_not_worse & m).sum()),
                           same_return_sign=int((same_ret & m).sum()),
                           gpu_reasons={int(k): int(v) for k, v in zip(*np.unique(reason[m], return_counts=True))})
    conv = np.isin(ref["reason"], CONVERGED)
    out["converged"] = dict(fits=int(conv.sum()), strict=int((p_ok & cost_ok & conv).sum()),
                            p_within_1e4=int((p_ok & conv).sum()), cost_within_1e6=int((cost_ok & conv).sum()))
    out["all"] = dict(fits=int(len(conv)), strict=int((p_ok & cost_ok).sum()), cost_within_1e6=int(cost_ok.sum()),
                      cost_not_worse_1e3=int(cost_not_worse.sum()), same_stop_reason=int(same_reason.sum()))
    return out, dict(p_ok=p_ok, cost_ok=cost_ok, conv=conv, cost_not_worse=cost_not_worse)


def record(name, hist):
    """Keep the histogram where the run's artefacts go (gpurun_out/ on the GPU box), for profiles/."""
    out = os.path.join(os.path.dirname(HERE), "gpurun_out")
    try:
        os.makedirs(out, exist_ok=True)
        with open(os.path.join(out, "parity_%s.json" % name), "w") as f:
            json.dump(hist, f, indent=1)
    except OSError:
        pass
