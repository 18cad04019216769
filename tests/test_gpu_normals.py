"""The move draw on the special-function unit against its defining arithmetic (csrc/abs_rng.cuh).

`displacement_fast` must (a) stay inside the error constants its acceptance test is built on -- measured here
exhaustively over all 2^23 radius inputs and all 2^25 angle inputs of the Box-Muller map -- and (b) never accept a
draw whose truncated displacement differs from the defining path (`normal_pair`, which oracle/philox.py restates).
"""
import pytest

pytestmark = pytest.mark.gpu


def test_fast_displacement_is_provably_the_defining_one(native_lib):
    from covid19_agentbasedsimulation_b200 import _native
    r = _native.selftest_normals(device=0, sweep_log2=31)
    # each constant carries a factor of two over the measured maximum
    assert r["radius_error_over_bound"] <= 0.5, r
    assert r["angle_error_over_bound"] <= 0.5, r
    assert r["wrong"] == 0, r
    assert r["draws"] == 2 ** 31
    # the fast path must actually be the one that runs: well over 99 % of the draws at every tested amplitude
    assert r["accepted"] >= 0.98 * r["draws"], r
    assert r["radius_inputs_rejected"] <= 8, r
