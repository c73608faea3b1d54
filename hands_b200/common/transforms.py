"""Drop-in for the hot functions of the reference's common/transforms.py (lines 69-77, 316-345)."""
from ..functional import Project2DFunction


def project2d_batch(K, pts_cam):
    """K (B,3,3), pts_cam (B,N,3) -> (B,N,2) pixels: (K X)[:2] / (K X)[2]  (transforms.py:316-329)."""
    return Project2DFunction.apply(K, pts_cam, 0.0)


def project2d_norm_batch(K, pts_cam, patch_width):
    """project2d_batch followed by normalize_kp2d (transforms.py:332-345), one kernel."""
    return Project2DFunction.apply(K, pts_cam, float(patch_width))
