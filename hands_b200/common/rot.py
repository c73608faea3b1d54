"""Drop-in for the hot functions of the reference's common/rot.py (lines 44-193)."""
from ..functional import MatrixToAxisAngleFunction, Rot6dToRotmatFunction


def matrix_to_axis_angle(matrix):
    """(...,3,3) rotation matrices -> (...,3) axis-angle; same branch behaviour and gradients as
    common/rot.py:180-193 (best-conditioned quaternion, 0.1 floor, small-angle series)."""
    if matrix.size(-1) != 3 or matrix.size(-2) != 3:
        raise ValueError(f"Invalid rotation matrix shape {matrix.shape}.")
    return MatrixToAxisAngleFunction.apply(matrix)


def rot6d_to_rotmat(x):
    """common/rot.py:367-381: (B,6) or (B,6k) -> (B*k,3,3); a1 = x[0,2,4], a2 = x[1,3,5] (reshape(-1,3,2)), columns b1,b2,b3."""
    return Rot6dToRotmatFunction.apply(x.reshape(-1, 6), "cols_paired")


def rotation_6d_to_matrix(d6):
    """pytorch3d.transforms.rotation_6d_to_matrix as the reference calls it (src/nets/hand_heads/hand_hmr.py:85-87):
    (...,6) -> (...,3,3); a1 = d6[:3], a2 = d6[3:], rows b1,b2,b3."""
    return Rot6dToRotmatFunction.apply(d6.reshape(-1, 6), "rows").reshape(d6.shape[:-1] + (3, 3))


def rot6d_to_rotmat_hamer(x):
    """src/models/hamer_light/geometry.py:47-62 (reshape(-1,2,3).permute(0,2,1)) == src/models/handoccnet_light/mano_head.py:132-141
    `rot6d2mat`: (B,6k) -> (B*k,3,3); a1 = x[0:3], a2 = x[3:6], columns b1,b2,b3."""
    return Rot6dToRotmatFunction.apply(x.reshape(-1, 6), "cols")


rot6d2mat = rot6d_to_rotmat_hamer
