"""pytorch3d.transforms stand-in: the oracle's restatement IS the declared semantics."""
from oracle.transforms import (  # noqa: F401
    quaternion_to_matrix, matrix_to_quaternion, quaternion_multiply, quaternion_invert,
    quaternion_apply, quaternion_raw_multiply, so3_exponential_map, axis_angle_to_matrix,
    random_quaternions, standardize_quaternion)
