"""brdf_b200/csrc/glibc_pow.cuh reproduces libm's pow() -- the function the reference's BRDFFunc calls once per
sample (brdfdata.cpp:981, 986) -- operation by operation; the levmar-exact batched mode stands on it.  The host
instantiation is compiled here and compared with libm's pow() bit for bit on millions of arguments: the BRDF
range, results in the subnormal range, negative bases, overflow, tiny exponents and the IEEE special values."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_port_equals_libm_pow(tmp_path):
    exe = tmp_path / "glibc_pow_check"
    src = os.path.join(ROOT, "tests", "native", "glibc_pow_check.cpp")
    # -ffp-contract=off: the port spells every fused operation itself; -mfma only makes __builtin_fma one instruction
    flags = ["-O2", "-std=c++17", "-ffp-contract=off"]
    if "fma" in open("/proc/cpuinfo").read():
        flags.append("-mfma")
    else:
        pytest.skip("host CPU without FMA: libm runs its non-FMA variant, which is a different instruction sequence")
    subprocess.check_call(["g++"] + flags + [src, "-o", str(exe)])
    out = subprocess.run([str(exe), "2"], capture_output=True, text=True)
    sys.stdout.write(out.stdout)
    assert out.returncode == 0 and " bad 0" in out.stdout, out.stdout[-2000:]


def test_tables_are_the_systems():
    """The generated table include equals what libm.so.6 of this image holds (profiles/tools/extract_pow_tables.py)."""
    inc = os.path.join(ROOT, "brdf_b200", "csrc", "glibc_pow_tables.inc")
    before = open(inc).read()
    libm = "/lib/x86_64-linux-gnu/libm.so.6"
    if not os.path.exists(libm):
        pytest.skip("no libm.so.6 at the usual place")
    subprocess.check_call([sys.executable, os.path.join(ROOT, "profiles", "tools", "extract_pow_tables.py"), libm],
                          stdout=subprocess.DEVNULL)
    after = open(inc).read()
    if after != before:
        open(inc, "w").write(before)
    assert after.split("\n", 1)[1] == before.split("\n", 1)[1]
