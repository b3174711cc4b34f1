"""CPU oracle for the PicoPose correspondence hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``picopose_b200/`` imports this
package; only ``tests/``, ``__graft_entry__.smoke()`` and the CPU-baseline /
``--impl reference`` legs of ``bench.py`` may.  The oracle is a restatement in
plain torch-CPU / numpy of the algorithms in the reference's

    utils/matching.py, utils/corr_lookup.py, utils/correspondence.py,
    model/stage3/raft_decoder.py:30-53 (CorrelationPyramid)

Parity pinning: the reference ships no tests and no golden vectors
(SURVEY.md section 4), so the oracle is pinned against outputs of the reference
modules themselves, generated in the build container by
``oracle/make_golden.py`` (which imports /root/reference) and committed as
``tests/golden/*.npz``; ``tests/test_oracle_golden.py`` replays them.
"""
