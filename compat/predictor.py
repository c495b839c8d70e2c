"""reference path: predictor.py -> dad_3dheads_b200.predictor"""
from dad_3dheads_b200.predictor import (FaceMeshPredictor, calculate_paddings, load_yaml, model_exists,  # noqa: F401
                                        py3round)
