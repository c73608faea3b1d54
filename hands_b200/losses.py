"""Key-point loss terms and metric partial sums computed straight from the head's outputs.

Drop-in for the terms of the reference's loss that read the geometry path's outputs
(`src/callbacks/loss/loss_arctic_sf.py:70-92,131-136`: `hand_kp3d_loss` on `mano.j3d.cam.*`, `joints_loss` on
`mano.j2d.norm.*`, MSE, masked by `joints_valid_*`, gated by `meta_info['is_j3d_loss'|'is_j2d_loss']`, `.mean()`),
and for the evaluation metrics on the same tensors (`common/metrics.py:23-55` as called from
`src/utils/eval_modules.py:95-118,373,407-421`): instead of per-metric `.cpu()` round trips the kernels leave
partial sums + counts on the device, ready for `hands_b200.distributed.PackedMetrics`.
"""
import torch

from . import _lib
from .functional import _f32c, _ptr, _stream

NOJ = 21


class KeypointLossFunction(torch.autograd.Function):
    """(j3d_cam, j2d_norm, gt_j3d_cam, gt_j2d_norm, joints_valid, hand_valid, gate_j3d, gate_j2d, img_res)
    -> (loss_kp3d, loss_kp2d, sums[8]); gradients flow to j3d_cam and j2d_norm only (the targets are data)."""

    @staticmethod
    def forward(ctx, j3d, j2d, gt3, gt2, jv, hv, gate3, gate2, img_res):
        lib = _lib.load()
        B = j3d.shape[0]
        j3d, gt3 = _f32c(j3d, "j3d_cam", (B, NOJ, 3)), _f32c(gt3, "gt_j3d_cam", (B, NOJ, 3))
        j2d, gt2 = _f32c(j2d, "j2d_norm", (B, NOJ, 2)), _f32c(gt2, "gt_j2d_norm", (B, NOJ, 2))
        jv = _f32c(jv, "joints_valid", (B, NOJ))
        hv, gate3, gate2 = _f32c(hv, "hand_valid", (B,)), _f32c(gate3, "gate_j3d", (B,)), _f32c(gate2, "gate_j2d", (B,))
        dev = j3d.device
        partial = torch.empty(max(B, 1), _lib.KP_SUMS, dtype=torch.float32, device=dev)
        sums = torch.empty(_lib.KP_SUMS, dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            rc = lib.hb_kp_loss_fwd(_ptr(j3d), _ptr(j2d), _ptr(gt3), _ptr(gt2), _ptr(jv), _ptr(hv), _ptr(gate3), _ptr(gate2), B, float(img_res),
                                    _ptr(partial), _ptr(sums), _stream())
        _lib.check(rc, "hb_kp_loss_fwd")
        ctx.save_for_backward(j3d, j2d, gt3, gt2, jv, gate3, gate2)
        ctx.set_materialize_grads(False)
        n = max(B, 1) * NOJ
        ctx.mark_non_differentiable(sums)
        return sums[0] / (n * 3), sums[1] / (n * 2), sums

    @staticmethod
    def backward(ctx, g3, g2, _g_sums):
        lib = _lib.load()
        j3d, j2d, gt3, gt2, jv, gate3, gate2 = ctx.saved_tensors
        B = j3d.shape[0]
        g3 = None if g3 is None else g3.reshape(1).contiguous().float()
        g2 = None if g2 is None else g2.reshape(1).contiguous().float()
        g_j3d = torch.empty_like(j3d) if ctx.needs_input_grad[0] else None
        g_j2d = torch.empty_like(j2d) if ctx.needs_input_grad[1] else None
        with torch.cuda.device(j3d.device):
            rc = lib.hb_kp_loss_bwd(_ptr(j3d), _ptr(j2d), _ptr(gt3), _ptr(gt2), _ptr(jv), _ptr(gate3), _ptr(gate2), B, _ptr(g3), _ptr(g2),
                                    _ptr(g_j3d), _ptr(g_j2d), _stream())
        _lib.check(rc, "hb_kp_loss_bwd")
        return g_j3d, g_j2d, None, None, None, None, None, None, None


def keypoint_losses(j3d_cam, j2d_norm, gt_j3d_cam, gt_j2d_norm, joints_valid, hand_valid=None, gate_j3d=None, gate_j2d=None, img_res=224):
    """Returns (loss_kp3d, loss_kp2d, sums) for one hand side; `sums` is the 8-float device vector described in
    include/hands_b200.h (MPJPE-RA and pixel-error numerators / counts in [2..5])."""
    return KeypointLossFunction.apply(j3d_cam, j2d_norm, gt_j3d_cam, gt_j2d_norm, joints_valid, hand_valid, gate_j3d, gate_j2d, float(img_res))


def mrrpe_sums(j3d_cam_r, j3d_cam_l, gt_j3d_cam_r, gt_j3d_cam_l, valid=None):
    """common/metrics.py:47-55 as partial sums: tensor([sum of valid distances, number of valid samples, 0...])."""
    lib = _lib.load()
    B = j3d_cam_r.shape[0]
    args = [_f32c(t, n, (B, NOJ, 3)) for t, n in ((j3d_cam_r, "j3d_cam_r"), (j3d_cam_l, "j3d_cam_l"), (gt_j3d_cam_r, "gt_j3d_cam_r"), (gt_j3d_cam_l, "gt_j3d_cam_l"))]
    valid = _f32c(valid, "valid", (B,))
    dev = args[0].device
    partial = torch.empty(max(B, 1), _lib.KP_SUMS, dtype=torch.float32, device=dev)
    sums = torch.empty(_lib.KP_SUMS, dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        rc = lib.hb_mrrpe(*[_ptr(t) for t in args], _ptr(valid), B, _ptr(partial), _ptr(sums), _stream())
    _lib.check(rc, "hb_mrrpe")
    return sums
