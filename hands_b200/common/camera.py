"""Drop-in for the hot functions of the reference's common/camera.py (lines 10-29, 456-474)."""
from ..functional import WeakToPerspFunction, persp_to_weak


def weak_perspective_to_perspective_torch(weak_perspective_camera, focal_length, img_res, min_s):
    """[s,tx,ty] -> [tx,ty, 2f/(img_res*max(s,min_s)+1e-9)]  (common/camera.py:456-474)."""
    return WeakToPerspFunction.apply(weak_perspective_camera, focal_length, img_res, min_s)


def perspective_to_weak_perspective_torch(perspective_camera, focal_length, img_res):
    """[tx,ty,tz] -> [2f/(img_res*tz+1e-9), tx, ty]  (common/camera.py:10-29)."""
    return persp_to_weak(perspective_camera, focal_length, img_res)
