"""Batched Perspective Crop Layer (seam 4 of SURVEY.md §8(b); no batched function exists in the
reference, where the closure at src/datasets/hands_light_dataset.py:354-467 runs per sample on the
CPU inside the data loader)."""
import torch

from .functional import PerspectiveCropFunction, RotApplyFunction


def perspective_crop(img, bbox_xyxy, K, img_res=224, crops_per_img=1):
    """img (Bi,3,img_res,img_res) fp32 CUDA; bbox_xyxy (Bi*crops_per_img,4) int [x0,y0,x1,y1] inside the
    image; K (Bi*crops_per_img,3,3).  Crop c samples image c // crops_per_img.
    Returns crop (Bi*crops_per_img,3,img_res,img_res) and R_virt2orig (Bi*crops_per_img,3,3).
    Semantics = lines 425-467 of the reference closure; gradient flows to img."""
    if img.shape[-1] != img_res or img.shape[-2] != img_res:
        raise ValueError(f"img must be {img_res}x{img_res} (the reference crops from the resized full image)")
    return PerspectiveCropFunction.apply(img, bbox_xyxy, K, int(crops_per_img))


def apply_virtual_rotation(R_virt2orig, pose):
    """pose (B,16,3,3): returns a copy with pose[:,0] <- R_virt2orig @ pose[:,0]
    (src/models/hands_light/model.py:330-334, written out-of-place so autograd stays valid)."""
    rotated = RotApplyFunction.apply(R_virt2orig, pose[:, 0])
    return torch.cat([rotated[:, None], pose[:, 1:]], dim=1)
