"""Drop-in for normalize_kp2d / unormalize_kp2d (reference common/data_utils.py:361-373).

Both are single affine maps; on the hot path they are fused into the projection kernel
(`project2d_norm_batch`, `MANOHead`).  The stand-alone forms are one fused torch expression each, kept
so callers such as src/models/generic/wrapper.py:118-134 keep working unchanged.
"""
import torch


def normalize_kp2d(kp2d: torch.Tensor, img_res):
    """pixels -> [-1,1]: 2*x/img_res - 1 on the first two channels; further channels pass through."""
    if kp2d.dim() != 3:
        raise AssertionError(f"kp2d must be (B,N,C), got {tuple(kp2d.shape)}")
    xy = (2.0 * kp2d[..., :2]) / img_res - 1.0
    if kp2d.shape[2] == 2:
        return xy
    return torch.cat([xy, kp2d[..., 2:]], dim=2)


def unormalize_kp2d(kp2d_normalized: torch.Tensor, img_res):
    """[-1,1] -> pixels: 0.5*img_res*(x+1)."""
    if kp2d_normalized.dim() != 3 or kp2d_normalized.shape[2] != 2:
        raise AssertionError(f"kp2d_normalized must be (B,N,2), got {tuple(kp2d_normalized.shape)}")
    return (kp2d_normalized + 1) * (0.5 * img_res)
