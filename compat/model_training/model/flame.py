"""reference path: model_training/model/flame.py -> dad_3dheads_b200.flame (+ the RPY helpers, flame.py:104,254-264)"""
from collections import namedtuple

import numpy as np
from scipy.spatial.transform import Rotation

from dad_3dheads_b200.flame import (EYE_COEFFS, FLAME_CONSTS, JAW_COEFFS, MAX_EXPRESSION, MAX_SHAPE,  # noqa: F401
                                    MESH_OFFSET_Z, NECK_COEFFS, ROT_COEFFS, FLAMELayer, FlameParams)
from model_training.model.utils import rot_mat_from_6dof

RPY = namedtuple("RPY", ["roll", "pitch", "yaw"])


def limit_angle(angle, pi=180.0):
    """Angle in degrees wrapped into [-pi, pi] (flame.py:238-251)."""
    if angle < -pi:
        angle = angle + (-2 * (int(angle / pi) // 2)) * pi
    if angle > pi:
        angle = angle - (2 * ((int(angle / pi) + 1) // 2)) * pi
    return angle


def rotation_mat_from_flame_params(flame_params):
    return rot_mat_from_6dof(flame_params.rotation).cpu().numpy()[0]


def calculate_rpy(flame_params) -> RPY:
    rot_mat = np.transpose(rotation_mat_from_flame_params(flame_params))
    angle = Rotation.from_matrix(rot_mat).as_euler("xyz", degrees=True)
    roll, pitch, yaw = list(map(limit_angle, [angle[2], angle[0] - 180, angle[1]]))
    return RPY(roll=roll, pitch=pitch, yaw=yaw)
