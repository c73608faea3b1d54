"""Drop-in for the hot functions of the reference's common/rot.py (lines 44-193)."""
from ..functional import MatrixToAxisAngleFunction


def matrix_to_axis_angle(matrix):
    """(...,3,3) rotation matrices -> (...,3) axis-angle; same branch behaviour and gradients as
    common/rot.py:180-193 (best-conditioned quaternion, 0.1 floor, small-angle series)."""
    if matrix.size(-1) != 3 or matrix.size(-2) != 3:
        raise ValueError(f"Invalid rotation matrix shape {matrix.shape}.")
    return MatrixToAxisAngleFunction.apply(matrix)
