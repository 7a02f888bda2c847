"""The C-ABI library loads without a GPU and exports every symbol include/brdfgpu.h declares; the
compute entry points fail loudly (LM_ERROR + message) when no CUDA device exists -- no CPU fallback."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from brdf_b200 import api as A

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "brdfgpu.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    names = set(re.findall(r"\b(brdfgpu_[A-Za-z0-9_]+)\s*\(", text))
    names -= {n for n in names if n.endswith("_t")}
    return sorted(names)


def test_header_symbols_are_exported():
    handle = C.CDLL(A.lib_path())
    names = declared_symbols()
    assert len(names) >= 45
    missing = [n for n in names if not hasattr(handle, n)]
    assert not missing, missing


def test_python_mirror_covers_header():
    assert sorted(A.SIGNATURES) == declared_symbols()


def test_no_torch_types_or_oracle_in_product():
    """The boundary is plain C; the product never references the oracle."""
    hdr = open(os.path.join(ROOT, "include", "brdfgpu.h")).read()
    code = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    assert "torch" not in code and "at::" not in code and "#include <cuda" not in code
    for dirpath, _, files in os.walk(os.path.join(ROOT, "brdf_b200")):
        for f in files:
            if f.endswith((".cu", ".cuh", ".py", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                # comments may NAME the oracle file that defines the gather arithmetic; nothing may
                # include, import, dlopen or call it
                for bad in ("oracle.h", "liboracle", "oracle_lib", "import oracle", "oracle_"):
                    assert bad not in src, (os.path.join(dirpath, f), bad)


def _has_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


@pytest.mark.skipif(_has_gpu(), reason="checks the behaviour WITHOUT a GPU")
def test_compute_fails_loudly_without_gpu():
    with pytest.raises(A.BrdfGpuError):
        A.Context()
    n = 8
    c = np.linspace(0.1, 0.9, n)
    extra, _keep = A.make_extra(c, c, c, 1)
    ret, p, info, _ = A.dlevmar_bc_dif([0.5, 1, 1], c, [0] * 3, [100] * 3, 10, [1e-3, 1e-15, 1e-15, 1e-20, 1e-6], extra)
    assert ret == A.LM_ERROR


def test_wrong_callback_is_rejected():
    """Any callback other than brdfgpu_BRDFFunc returns LM_ERROR: there is no CPU path to run it."""
    n = 8
    c = np.linspace(0.1, 0.9, n)
    extra, _keep = A.make_extra(c, c, c, 1)
    bogus = C.cast(A.lib().brdfgpu_version, C.c_void_p)
    ret, _, _, _ = A.dlevmar_bc_dif([0.5, 1, 1], c, [0] * 3, [100] * 3, 10, None, extra, func=bogus)
    assert ret == A.LM_ERROR


def test_m_other_than_3_is_rejected():
    """include/brdfgpu.h: the BRDF models have exactly 3 parameters; levmar's generic m (levmar.h:124-127) is refused
    before any GPU work, with the documented message."""
    n = 8
    c = np.linspace(0.1, 0.9, n)
    extra, _keep = A.make_extra(c, c, c, 1)
    for m in (2, 4, 8):
        ret, _, _, _ = A.dlevmar_bc_dif([0.5] * m, c, [0] * m, [100] * m, 10, None, extra)
        assert ret == A.LM_ERROR
        msg = A.lib().brdfgpu_last_error(None).decode()
        assert "exactly 3 parameters" in msg and "m = %d" % m in msg, msg
    ret, _, _, _ = A.dlevmar_dif([0.5] * 4, c, 10, None, extra)
    assert ret == A.LM_ERROR
    hdr = open(os.path.join(ROOT, "include", "brdfgpu.h")).read()
    assert "3 <= m" not in hdr and "BRDFGPU_NUM_PARAMS 3" in hdr
