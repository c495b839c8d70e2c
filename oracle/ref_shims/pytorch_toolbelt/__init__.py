"""Shim of pytorch_toolbelt (absent): ``utils.read_rgb_image`` (demo.py:5), ``utils.transfer_weights`` (model/utils.py:27)
and ``modules.instantiate_activation_block`` (layers.py:22)."""
