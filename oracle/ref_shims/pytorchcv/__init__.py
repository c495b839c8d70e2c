"""Shim of pytorchcv==0.0.65 (absent): only ``model_provider.get_model("resnet50")`` (reference: encoders.py:5,21)."""
