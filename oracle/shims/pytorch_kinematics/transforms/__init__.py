from oracle.tp_pytorch_kinematics import (  # noqa: F401
    Transform3d, Translate, euler_angles_to_matrix, quaternion_to_matrix, matrix_to_rotation_6d,
    random_rotation, random_rotations, axis_and_angle_to_matrix_33)
from . import rotation_conversions  # noqa: F401
