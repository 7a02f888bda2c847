"""Synthetic BRDF samples of the shape BASELINE.json names.

Counter-based generator so any sample can be produced independently (numpy here, the same integer
recipe in brdf_b200/csrc/synth.cu on the device):

    u(i, s) = (splitmix64_finalise(seed + (4*i + s + 1) * 0x9E3779B97F4A7C15) >> 11) * 2**-53

    cosphi_i = u(i,0)   costhetadash_i = u(i,1)   costheta_i = u(i,2)      all in [0, 1)
    x_i = clamp(floor(255 * (kd*cosphi + coef*ks*t**n + (u(i,3) - 0.5) * 0.01)), 0, 255) / 255

i.e. the truth model plus +-0.005 uniform noise, quantised to 8 bits like the photographs.
"""
import numpy as np

GOLDEN = np.uint64(0x9E3779B97F4A7C15)
M1 = np.uint64(0xBF58476D1CE4E5B9)
M2 = np.uint64(0x94D049BB133111EB)
DEFAULT_SEED = 88172645463325252
TRUTH = (0.6, 0.35, 12.0)
PI = 3.1415926535897932384626433832795


def _finalise(z):
    z = (z ^ (z >> np.uint64(30))) * M1
    z = (z ^ (z >> np.uint64(27))) * M2
    return z ^ (z >> np.uint64(31))


def uniform(idx, stream, seed=DEFAULT_SEED):
    """u(i, s) for an array of sample indices (uint64 arithmetic wraps, as on the device)."""
    with np.errstate(over="ignore"):
        ctr = idx.astype(np.uint64) * np.uint64(4) + np.uint64(stream + 1)
        z = np.uint64(seed) + ctr * GOLDEN
        return (_finalise(z) >> np.uint64(11)).astype(np.float64) * (2.0 ** -53)


def model(p, cosphi, t, model_id):
    coef = 1.0 if model_id == 1 else ((p[2] + 2.0) / 2.0 * PI)
    return p[0] * cosphi + coef * p[1] * np.power(t, p[2])


def samples(n, truth=TRUTH, model_id=1, seed=DEFAULT_SEED, start=0):
    """Returns (cosphi, costhetadash, costheta, x) for samples start .. start+n-1."""
    idx = np.arange(start, start + n, dtype=np.uint64)
    cosphi = uniform(idx, 0, seed)
    costd = uniform(idx, 1, seed)
    costh = uniform(idx, 2, seed)
    noise = (uniform(idx, 3, seed) - 0.5) * 0.01
    t = costd if model_id == 1 else costh
    val = model(truth, cosphi, t, model_id) + noise
    x = np.clip(np.floor(255.0 * val), 0.0, 255.0) / 255.0
    return cosphi, costd, costh, x


def batched(nfit, nper, model_id=1, seed=DEFAULT_SEED):
    """Config-4 generator: per fit p* = (U(0.1,0.9), U(0.05,0.8), U(1,50)); samples as above.

    Fit f uses sample indices f*nper .. f*nper+nper-1 of stream `seed`, its truth comes from
    stream seed+1 at index f.  Returns (cosphi, costd, costh, x) shaped (nfit, nper) and truth (nfit,3).
    """
    fid = np.arange(nfit, dtype=np.uint64)
    kd = 0.1 + 0.8 * uniform(fid, 0, seed + 1)
    ks = 0.05 + 0.75 * uniform(fid, 1, seed + 1)
    nn = 1.0 + 49.0 * uniform(fid, 2, seed + 1)
    truth = np.stack([kd, ks, nn], axis=1)
    idx = np.arange(nfit * nper, dtype=np.uint64)
    cosphi = uniform(idx, 0, seed).reshape(nfit, nper)
    costd = uniform(idx, 1, seed).reshape(nfit, nper)
    costh = uniform(idx, 2, seed).reshape(nfit, nper)
    noise = ((uniform(idx, 3, seed) - 0.5) * 0.01).reshape(nfit, nper)
    t = costd if model_id == 1 else costh
    coef = 1.0 if model_id == 1 else ((nn + 2.0) / 2.0 * PI)[:, None]
    val = kd[:, None] * cosphi + coef * ks[:, None] * np.power(t, nn[:, None]) + noise
    x = np.clip(np.floor(255.0 * val), 0.0, 255.0) / 255.0
    return cosphi, costd, costh, x, truth
