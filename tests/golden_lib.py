"""Loader for tests/golden/brdf_fits.json (outputs of the reference's levmar; see make_golden.py)."""
import json
import os

import numpy as np

import synth

HERE = os.path.dirname(os.path.abspath(__file__))


def load():
    with open(os.path.join(HERE, "golden", "brdf_fits.json")) as f:
        g = json.load(f)
    for case in g["global"]:
        case["p"] = np.array([float.fromhex(v) for v in case["p"]])
        case["info"] = np.array([float.fromhex(v) for v in case["info"]])
    for case in g["batch"]:
        case["p"] = np.array([float.fromhex(v) for v in case["p"]]).reshape(case["nfit"], 3)
        case["info"] = np.array([float.fromhex(v) for v in case["info"]]).reshape(case["nfit"], 10)
        case["ret"] = np.array(case["ret"])
    return g


def global_inputs(case):
    c, td, th, x = synth.samples(case["n"], model_id=case["model"], seed=case["seed"])
    if case["negate_every"]:
        td = td.copy(); th = th.copy()
        td[::case["negate_every"]] *= -1.0
        th[::case["negate_every"]] *= -1.0
    return c, td, th, x


def batch_inputs(case):
    return synth.batched(case["nfit"], case["nper"], model_id=case["model"], seed=case["seed"])
