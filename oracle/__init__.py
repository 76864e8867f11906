"""CPU oracles: restatements of the reference's algorithms for the hot path -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Only tests/ (incl. tests/golden/make_golden*.py and tests/cond_cpu.py), __graft_entry__.smoke() and the cpu_baseline /
`--impl reference` legs of bench.py may import this package; nothing under ts-asr-whisper_b200/ does, and the product path
raises when its CUDA library is missing instead of falling back to anything here.  Every oracle is pinned to outputs of
the reference itself: stored goldens (tests/golden/*.npz) and, in the build container, live comparisons
(tests/test_reference_live.py).
"""
