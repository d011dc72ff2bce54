"""oracle/ -- TEST INFRASTRUCTURE ONLY.

CPU restatements of the LAS forward hot path of jiwidi/las-pytorch
(`model/las_model.py`, `utils/functions.py:54-77`).  Nothing under this package is part of
the product: only `tests/`, `__graft_entry__.smoke()` and the `cpu_baseline` / `--impl
reference` legs of `bench.py` may import it, and only as the checker or as the timed CPU
baseline.  `las_pytorch_b200` never imports it.

Parity pinning: the reference ships no tests, golden vectors or fixtures of its own
(SURVEY.md section 4 / 8c), so the restatements are pinned against outputs of the reference
itself: `tests/golden/make_golden.py` imports the unmodified reference from
`/root/reference` (three data-prep-only third-party modules stubbed), runs it on CPU in fp32
and fp64, and freezes inputs/weights/outputs under `tests/golden/*.npz`;
`tests/test_oracle_golden.py` checks both restatements against those files.
"""
