"""ORACLE - test infrastructure only (CPU restatement of the reference hot path).

Nothing under oracle/ is part of the product. Only tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs may import it, and only as the checker or the
CPU baseline. Unet numerics: PARITY UNPINNED (reference is JAX/flax, not installable here, and its
tests pin shapes only); diffusion math: pinned by the reference's known-answer tests.
"""
