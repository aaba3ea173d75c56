"""CPU oracle — TEST INFRASTRUCTURE ONLY.

A CPU restatement of the reference's stage-1 per-frame path (keypoint math,
network wiring, losses, optimiser) with TensorFlow-1.12 op semantics.  Only
``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline /
``--impl reference`` legs may import it; the product package never does.

PARITY PINNING STATUS: the reference ships no tests, golden vectors or
checkpoints and TensorFlow 1.12 cannot be installed here, so the TF *op*
semantics (SAME padding, legacy bilinear resize, fused batch-norm, Adam) are
restated from TF-1.12 behaviour — "parity unpinned" for those.  The *wiring*
and keypoint math ARE pinned: ``oracle/tf_shim`` lets the reference's own,
unmodified ``utils/model.py`` and ``models/networks/__init__.py`` execute
eagerly on numpy/torch, and ``tests/golden/make_golden.py`` stores their
outputs as fixtures the oracle (and the CUDA path) are checked against.
"""
