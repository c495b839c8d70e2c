"""albumentations.augmentations.geometric.py3round (imported by predictor.py:12)."""


def py3round(number):
    """Unified rounding in all python versions (round half to even)."""
    if abs(round(number) - number) == 0.5:
        return int(2.0 * round(number / 2.0))
    return int(round(number))
