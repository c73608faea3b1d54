"""Drop-in for the reference's src/callbacks/process/process_arctic.py:4-75 (`process_data_light`), the GT side of every
training / validation step: two no-grad MANO forwards of the ground-truth parameters, the mean-offset translation into
camera space, the GT camera translation and its weak-perspective form.  ~250 torch ops in the reference; here three
launches per hand side for the MANO forward plus one for the glue."""
import torch

from .... import _lib
from ....functional import _f32c, _ptr, _stream


def _gt_side(mano, pose, betas, j3d_full, K, img_res):
    with torch.no_grad():
        out = mano(betas=betas, hand_pose=pose[:, 3:], global_orient=pose[:, :3], transl=None)
        B = betas.shape[0]
        joints, verts = _f32c(out.joints, "joints", (B, 21, 3)), _f32c(out.vertices, "vertices", (B, 778, 3))
        j3d_full = _f32c(j3d_full, "mano.j3d.full", (B, 21, 3))
        K = _f32c(K, "intrinsics", (B, 3, 3))
        v3d_cam, cam_t, cam_t_wp = torch.empty_like(verts), torch.empty(B, 3, dtype=torch.float32, device=verts.device), torch.empty(B, 3, dtype=torch.float32, device=verts.device)
        with torch.cuda.device(verts.device):
            rc = _lib.load().hb_gt_process(_ptr(joints), _ptr(verts), _ptr(j3d_full), _ptr(K), B, float(img_res), _ptr(v3d_cam), _ptr(cam_t), _ptr(cam_t_wp), _stream())
        _lib.check(rc, "hb_gt_process")
    return joints, verts, v3d_cam, cam_t, cam_t_wp


def process_data_light(models, inputs, targets, meta_info, mode, args, field_max=float("inf")):
    """Same signature, same keys written into `targets` as the reference (process_arctic.py:4-75).  `models["mano_r"]`,
    `models["mano_l"]` are `hands_b200.common.body_models.build_mano_aa(...)` layers."""
    img_res = args.img_res if hasattr(args, "img_res") else args["img_res"]
    K = meta_info["intrinsics"]
    for side, key in (("r", "mano_r"), ("l", "mano_l")):
        joints, verts, v3d_cam, cam_t, cam_t_wp = _gt_side(models[key], targets[f"mano.pose.{side}"], targets[f"mano.beta.{side}"],
                                                          targets[f"mano.j3d.full.{side}"], K, img_res)
        targets[f"mano.joints3d.{side}"] = joints      # MANO canonical space
        targets[f"mano.vertices.{side}"] = verts
        targets[f"mano.cam_t.{side}"] = cam_t
        targets[f"mano.cam_t.wp.{side}"] = cam_t_wp
        targets[f"mano.v3d.cam.{side}"] = v3d_cam
        targets[f"mano.j3d.cam.{side}"] = targets[f"mano.j3d.full.{side}"]
    return inputs, targets, meta_info
