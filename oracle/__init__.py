"""CPU oracle for the DAD-3DNet image->3D-head hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is part of the product: only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs may import it, and only as
the checker (or as the timed CPU baseline), never as the thing shipped.  The product package
(``dad_3dheads_b200``) never imports this package and fails loudly when its CUDA library is missing.

PINNED TO THE REFERENCE'S OWN SOURCE (round 2).  The reference (PinataFarms/DAD-3DHeads) ships no tests or golden vectors
for this path, but its arithmetic source is in /root/reference and only third-party imports keep it from running here.
``oracle/ref_harness.py`` supplies shims for exactly those absent packages (``oracle/ref_shims``) and runs the UNMODIFIED
reference files -- predictor.py, model_training/head_mesh.py, model/flame.py, model/utils.py, model/flame_regression.py,
model/bifpn.py, model/encoders.py -- from /root/reference, or from their byte-compiled twin ``oracle/_ref``
(``oracle/build_ref.py``; git-ignored, travels to the GPU box).  The restatements in this package
(flame_oracle / encoder_oracle / predictor_oracle / resize_oracle) agree with that reference to <= 1e-12 in fp64 and to
fp32 round-off in fp32 (tests/test_oracle_pinned.py, live and against the committed reference outputs
tests/golden/reference_*.npz made by tools/make_reference_golden.py); the packed asset equals the reference's FLAMELayer
buffers bit for bit.  Residue that is NOT reference-owned and therefore restated twice (shim and oracle, compared with each
other): ``smplx==0.1.26`` ``lbs`` and the ``pytorchcv==0.0.65`` ResNet-50 body; albumentations is restated over the real cv2.
"""
