"""CPU oracle for the DAD-3DNet image->3D-head hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is part of the product: only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs may import it, and only as
the checker (or as the timed CPU baseline), never as the thing shipped.  The product package
(``dad_3dheads_b200``) never imports this package and fails loudly when its CUDA library is missing.

PARITY UNPINNED: the reference (PinataFarms/DAD-3DHeads) ships no tests, golden vectors or known-answer fixtures for
this path, its python package cannot be imported in the build image (albumentations, smplx,
pytorchcv, hydra ... are absent) and the released checkpoint is not available offline.  The functions below are
line-by-line restatements of the reference files they cite plus restatements of the two pinned third-party
dependencies that hold the arithmetic (smplx==0.1.26 ``lbs``; pytorchcv==0.0.65 ``resnet50``), checked against
closed-form identities (tests/test_oracle_identities.py), not against reference outputs.
"""
