from oracle.tp_pytorch_kinematics import (  # noqa: F401
    euler_angles_to_matrix, quaternion_to_matrix, matrix_to_rotation_6d, random_rotation, random_rotations,
    axis_and_angle_to_matrix_33)


def matrix_to_pos_rot(m):
    """Only reached from the reference's GUI branch (chamfer.py:103); not on the hot path."""
    raise NotImplementedError
