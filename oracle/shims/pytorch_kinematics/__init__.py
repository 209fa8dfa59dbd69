"""Shim: see oracle/shims/README.md and oracle/tp_pytorch_kinematics.py."""
from oracle.tp_pytorch_kinematics import (  # noqa: F401
    Transform3d, Translate, Chain, Frame, Link, Joint, Visual, build_serial_chain_from_urdf,
    rotation_conversions, euler_angles_to_matrix, quaternion_to_matrix, matrix_to_rotation_6d,
    random_rotation, random_rotations, axis_and_angle_to_matrix_33)
from . import transforms  # noqa: F401
