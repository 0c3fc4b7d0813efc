"""CPU: the Add rewrites of wdpm_b200/csrc/relax.cuh (no cap, sign gate) leave every bit of the reference
step (src/runoff.cl:24-55) unchanged - random and adversarial chains, compiled for the host."""
import ctypes as C
import subprocess
from pathlib import Path

import pytest

HERE = Path(__file__).resolve().parent
SRC = HERE / "emul" / "relax_equiv.cpp"


def _build():
    out = HERE / "emul" / "_build" / "librelax_equiv.so"
    out.parent.mkdir(exist_ok=True)
    deps = [SRC, HERE.parent / "wdpm_b200" / "csrc" / "relax.cuh"]
    if not out.exists() or any(d.stat().st_mtime > out.stat().st_mtime for d in deps):
        subprocess.run(["/usr/bin/g++", "-O2", "-std=c++17", "-ffp-contract=off", "-fno-fast-math", "-shared", "-fPIC",
                        str(SRC), "-o", str(out)], check=True)
    return C.CDLL(str(out))


@pytest.mark.parametrize("sfx", ["f64", "f32"])
def test_add_fast_step_is_bit_exact_and_the_cap_never_bites(sfx):
    fn = getattr(_build(), "relax_equiv_run_" + sfx)
    fn.restype = C.c_longlong
    bites = C.c_longlong()
    bad = fn(C.c_longlong(3_000_000), C.c_ulonglong(12345), C.byref(bites))
    assert bad == 0
    assert bites.value == 0  # mini(flow, wc) is a no-op in runoffadd (proof in relax.cuh)


def test_drain_fast_step_is_bit_exact_without_negative_zero_water():
    """push_drain_fast (gate and max(flow,0) folded into the factor) against the reference form; the solver
    uses it only with a zero threshold > 0, i.e. when no water value is -0.0."""
    fn = _build().relax_equiv_drain_f64
    fn.restype = C.c_longlong
    assert fn(C.c_longlong(3_000_000), C.c_ulonglong(777)) == 0
