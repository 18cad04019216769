"""CPU oracle for the COVID-ABS per-iteration agent step.

TEST INFRASTRUCTURE ONLY.  Nothing in the product package
(`covid19_agentbasedsimulation_b200`) may import, call, link or execute
anything under `oracle/`.  The only allowed users are `tests/`,
`__graft_entry__.smoke()` and the `cpu_baseline` / `--impl reference` legs of
`bench.py`, and there only as the checker or the timed CPU baseline.

Contents
--------
seq_oracle.py    restatement of the reference's *sequential* step
                 (covid_abs/abs.py:97-322) with a pluggable draw source.
                 Parity pinned: with the numpy legacy stream as draw source it
                 is bit-identical to the unmodified reference (see
                 tests/test_oracle_vs_reference.py, which imports
                 /root/reference when present, and tests/golden/*.json which
                 were generated from the reference by oracle/gen_golden.py).
philox.py        independent Philox4x32-10 + deterministic float32 Box-Muller
                 (known-answer vectors of Random123 in the tests).
sync_oracle.py   restatement of the same step under the GPU's *synchronous*
                 schedule with counter-based draws keyed (seed, agent, step);
                 the CUDA path must match this bit for bit.
"""
