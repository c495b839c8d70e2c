"""Import stub of scikit-image (absent; model_training/data/utils.py:5 imports skimage.io.imread at module level)."""
