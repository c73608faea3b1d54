"""CPU ORACLE — TEST INFRASTRUCTURE ONLY.  Never imported by the product (`hands_b200/`).

A plain-torch (CPU, fp32 or fp64) restatement of the reference's geometry hot path.  Only
`tests/`, `__graft_entry__.smoke()` and the `cpu_baseline` / `--impl reference` legs of
`bench.py` may import this file, and there only as the checker / the CPU baseline.

Parity status
-------------
* Reference-owned steps (log map, camera, projection, key-point normalisation, the
  Perspective Crop Layer) are PINNED: `tests/golden/make_golden.py` ran the reference's own
  functions from `/root/reference` in the build container and committed their inputs/outputs
  as fixtures; `tests/test_oracle_golden.py` checks this file against them.
* The MANO arithmetic itself (shape/pose blendshapes, Rodrigues, kinematic chain, LBS,
  fingertip selector) lives in the third-party package `smplx` (upstream vchoutas/smplx,
  `smplx/lbs.py`, `smplx/body_models.py::MANO`; the reference pins no version, ARCTIC's setup
  doc uses smplx==0.1.28 with `vertex_joint_selector` enabled), which is neither under
  `/root/reference` nor installed here, and the reference has no tests or golden vectors.
  For that part: **parity unpinned** — this file restates the published algorithm
   and is anchored on the reference's call sites
  (`src/nets/hand_heads/mano_head.py:34-38`, `src/callbacks/process/process_arctic.py:16-34`,
  `src/arctic/processing.py:175-190`) and on invariance / known-answer tests.

Every function cites the reference lines it follows.
"""
import math

import torch
import torch.nn.functional as F

TIP_IDS = (744, 320, 443, 554, 671)  # smplx vertex_ids['mano']: thumb,index,middle,ring,pinky


# --------------------------------------------------------------------------------------
# log map: rotation matrix -> quaternion -> axis-angle     (common/rot.py:44-193)
# --------------------------------------------------------------------------------------
def sqrt_positive_part(x):
    """sqrt(max(0,x)) with a zero sub-gradient at x<=0 (common/rot.py:44-52)."""
    pos = x > 0
    safe = torch.where(pos, x, torch.ones_like(x))
    return torch.where(pos, torch.sqrt(safe), torch.zeros_like(x))


def matrix_to_quaternion(m):
    """(...,3,3) -> (...,4) real-first; best-conditioned of four candidates
    (common/rot.py:118-177: floor 0.1 at :169-170, argmax pick at :175-177)."""
    lead = m.shape[:-2]
    m = m.reshape(lead + (9,))
    m00, m01, m02, m10, m11, m12, m20, m21, m22 = m.unbind(-1)
    q_abs = sqrt_positive_part(
        torch.stack(
            [1.0 + m00 + m11 + m22, 1.0 + m00 - m11 - m22, 1.0 - m00 + m11 - m22, 1.0 - m00 - m11 + m22],
            dim=-1,
        )
    )
    sq = q_abs * q_abs
    cand = torch.stack(
        [
            torch.stack([sq[..., 0], m21 - m12, m02 - m20, m10 - m01], dim=-1),
            torch.stack([m21 - m12, sq[..., 1], m10 + m01, m02 + m20], dim=-1),
            torch.stack([m02 - m20, m10 + m01, sq[..., 2], m12 + m21], dim=-1),
            torch.stack([m10 - m01, m20 + m02, m21 + m12, sq[..., 3]], dim=-1),
        ],
        dim=-2,
    )
    floor = torch.full_like(q_abs, 0.1)
    cand = cand / (2.0 * torch.maximum(q_abs, floor)[..., None])
    idx = q_abs.argmax(dim=-1)
    pick = idx[..., None, None].expand(lead + (1, 4))
    return torch.gather(cand, -2, pick).squeeze(-2)


def quaternion_to_axis_angle(q):
    """(...,4) -> (...,3) (common/rot.py:55-83: atan2, small-angle series 0.5 - a^2/48 for |a|<1e-6)."""
    xyz = q[..., 1:]
    n = torch.linalg.vector_norm(xyz, dim=-1, keepdim=True)
    half = torch.atan2(n, q[..., :1])
    ang = 2 * half
    small = ang.abs() < 1e-6
    safe_ang = torch.where(small, torch.ones_like(ang), ang)
    regular = torch.sin(half) / safe_ang
    series = 0.5 - (ang * ang) / 48
    return xyz / torch.where(small, series, regular)


def matrix_to_axis_angle(m):
    """common/rot.py:180-193."""
    return quaternion_to_axis_angle(matrix_to_quaternion(m))


# --------------------------------------------------------------------------------------
# 6D rotation representation -> rotation matrix (three layouts in the reference)
# --------------------------------------------------------------------------------------
def _gram_schmidt(a1, a2):
    b1 = F.normalize(a1, dim=-1)
    b2 = F.normalize(a2 - torch.einsum("...i,...i->...", b1, a2).unsqueeze(-1) * b1, dim=-1)
    b3 = torch.cross(b1, b2, dim=-1)
    return b1, b2, b3


def rot6d_to_rotmat_paired(x):
    """common/rot.py:367-381: reshape(-1,3,2); a1 = x[:, :, 0], a2 = x[:, :, 1]; b1,b2,b3 stacked as COLUMNS."""
    x = x.reshape(-1, 3, 2)
    return torch.stack(_gram_schmidt(x[:, :, 0], x[:, :, 1]), dim=-1)


def rot6d_to_rotmat_cols(x):
    """src/models/hamer_light/geometry.py:47-62: reshape(-1,2,3).permute(0,2,1).contiguous(); a1 = x[:, :, 0] (= x[0:3]),
    a2 = x[:, :, 1] (= x[3:6]); COLUMNS.  (Same numbers as `rot6d2mat` up to the last bit: the slices are strided here.)"""
    x = x.reshape(-1, 2, 3).permute(0, 2, 1).contiguous()
    return torch.stack(_gram_schmidt(x[:, :, 0], x[:, :, 1]), dim=-1)


def rot6d2mat(x):
    """src/models/handoccnet_light/mano_head.py:132-141: a1 = x[:, 0:3], a2 = x[:, 3:6]; COLUMNS."""
    x = x.reshape(-1, 6)
    return torch.stack(_gram_schmidt(x[:, 0:3], x[:, 3:6]), dim=-1)


def rotation_6d_to_matrix(d6):
    """pytorch3d.transforms.rotation_6d_to_matrix, called at src/nets/hand_heads/hand_hmr.py:85-87.  pytorch3d is a
    third-party dependency absent here (no version pinned by the reference): restated from its published algorithm
    (a1 = d6[..., :3], a2 = d6[..., 3:], Gram-Schmidt, b1,b2,b3 stacked as ROWS); numerically the transpose of
    `rot6d2mat`, which IS pinned by a golden fixture."""
    return torch.stack(_gram_schmidt(d6[..., :3], d6[..., 3:]), dim=-2)


# --------------------------------------------------------------------------------------
# smplx MANO forward  [smplx-recalled; SURVEY.md Appendix A steps 1-9]
# --------------------------------------------------------------------------------------
def batch_rodrigues(rot_vecs):
    """(N,3) -> (N,3,3).  smplx/lbs.py::batch_rodrigues: angle = ||r + 1e-8|| (added to each
    component before the norm), R = I + sin*K + (1-cos)*K@K."""
    n = rot_vecs.shape[0]
    angle = torch.linalg.vector_norm(rot_vecs + 1e-8, dim=1, keepdim=True)
    d = rot_vecs / angle
    c = torch.cos(angle)[:, None]
    s = torch.sin(angle)[:, None]
    rx, ry, rz = d[:, 0], d[:, 1], d[:, 2]
    z = torch.zeros_like(rx)
    K = torch.stack([z, -rz, ry, rz, z, -rx, -ry, rx, z], dim=1).reshape(n, 3, 3)
    eye = torch.eye(3, dtype=rot_vecs.dtype)[None]
    return eye + s * K + (1 - c) * torch.bmm(K, K)


def batch_rigid_transform(rot_mats, joints, parents):
    """smplx/lbs.py::batch_rigid_transform.  rot_mats (B,16,3,3), joints (B,16,3).
    Returns posed joints (B,16,3) and relative transforms A (B,16,4,4)."""
    B, NJ = joints.shape[:2]
    rel = joints.clone()
    rel[:, 1:] = joints[:, 1:] - joints[:, parents[1:]]
    bottom = torch.zeros(B, NJ, 1, 4, dtype=joints.dtype)
    bottom[..., 3] = 1
    M = torch.cat([torch.cat([rot_mats, rel[..., None]], dim=-1), bottom], dim=-2)  # (B,16,4,4)
    chain = [M[:, 0]]
    for i in range(1, NJ):
        chain.append(torch.matmul(chain[int(parents[i])], M[:, i]))
    G = torch.stack(chain, dim=1)
    posed = G[:, :, :3, 3]
    jh = torch.cat([joints, torch.zeros(B, NJ, 1, dtype=joints.dtype)], dim=2)[..., None]
    init_bone = torch.matmul(G, jh)  # (B,16,4,1)
    A = G - F.pad(init_bone, [3, 0])
    return posed, A


def mano_forward(buf, betas, global_orient, hand_pose, transl=None, return_aux=False):
    """smplx.MANO(use_pca=False).forward as the reference constructs it
    (common/body_models.py:92-99) and calls it (mano_head.py:34-38).
    buf: dict with v_template (778,3), shapedirs (778,3,10), posedirs (135,2334),
    J_regressor (16,778), lbs_weights (778,16), parents (16), pose_mean (48).
    Returns vertices (B,778,3), joints (B,21,3)."""
    dt = betas.dtype
    c = {k: (v.to(dt) if v.is_floating_point() else v) for k, v in buf.items()}
    B = betas.shape[0]
    full_pose = torch.cat([global_orient, hand_pose], dim=1) + c["pose_mean"]
    v_shaped = c["v_template"] + torch.einsum("bl,mkl->bmk", betas, c["shapedirs"])
    J = torch.einsum("bik,ji->bjk", v_shaped, c["J_regressor"])
    R = batch_rodrigues(full_pose.reshape(-1, 3)).reshape(B, -1, 3, 3)
    pose_feature = (R[:, 1:] - torch.eye(3, dtype=dt)).reshape(B, -1)
    v_posed = v_shaped + torch.matmul(pose_feature, c["posedirs"]).reshape(B, -1, 3)
    parents = c["parents"].tolist()
    J_posed, A = batch_rigid_transform(R, J, parents)
    W = c["lbs_weights"][None].expand(B, -1, -1)
    T = torch.matmul(W, A.reshape(B, 16, 16)).reshape(B, -1, 4, 4)
    vh = torch.cat([v_posed, torch.ones(B, v_posed.shape[1], 1, dtype=dt)], dim=2)
    verts = torch.matmul(T, vh[..., None])[:, :, :3, 0]
    tips = verts[:, list(TIP_IDS)]
    joints = torch.cat([J_posed, tips], dim=1)
    if transl is not None:
        verts = verts + transl[:, None]
        joints = joints + transl[:, None]
    if return_aux:
        return verts, joints, {"v_posed": v_posed, "A": A, "R": R, "J": J}
    return verts, joints


# --------------------------------------------------------------------------------------
# camera + projection      (common/camera.py, common/transforms.py, common/data_utils.py)
# --------------------------------------------------------------------------------------
def weak_perspective_to_perspective(cam, focal_length, img_res, min_s):
    """common/camera.py:456-474."""
    s = torch.clamp(cam[:, 0], min_s)
    return torch.stack([cam[:, 1], cam[:, 2], 2 * focal_length / (img_res * s + 1e-9)], dim=-1)


def perspective_to_weak_perspective(cam_t, focal_length, img_res):
    """common/camera.py:10-29."""
    return torch.stack([2 * focal_length / (img_res * cam_t[:, 2] + 1e-9), cam_t[:, 0], cam_t[:, 1]], dim=-1)


def project2d_batch(K, pts_cam):
    """common/transforms.py:316-329 with to_xy_batch :69-77 (divide by z, no eps)."""
    p = torch.bmm(K, pts_cam.permute(0, 2, 1)).permute(0, 2, 1)
    return p[:, :, :2] / p[:, :, 2:3]


def normalize_kp2d(kp2d, img_res):
    """common/data_utils.py:361-365."""
    out = kp2d.clone()
    out[:, :, :2] = 2.0 * kp2d[:, :, :2] / img_res - 1.0
    return out


def unormalize_kp2d(kp2d_norm, img_res):
    """common/data_utils.py:368-373."""
    return 0.5 * img_res * (kp2d_norm + 1)


def mano_head_forward(buf, rotmat, shape, cam, K, img_res=224.0, min_s=0.1):
    """src/nets/hand_heads/mano_head.py:21-65 without the xdict packaging.
    rotmat (B,16,3,3) or (B,48) axis-angle.  Returns a plain dict with the nine un-postfixed keys."""
    pose_in = rotmat
    aa = rotmat
    if rotmat.shape[-1] != 48:
        aa = matrix_to_axis_angle(rotmat.reshape(-1, 3, 3)).reshape(-1, 48)
    verts, joints = mano_forward(buf, shape, aa[:, :3], aa[:, 3:])
    f = (K[:, 0, 0] + K[:, 1, 1]) / 2.0
    cam_t = weak_perspective_to_perspective(cam, f, img_res, min_s)
    j3d_cam = joints + cam_t[:, None, :]
    v3d_cam = verts + cam_t[:, None, :]
    j2d = normalize_kp2d(project2d_batch(K, j3d_cam), img_res)
    return {
        "cam_t.wp": cam,
        "cam_t": cam_t,
        "joints3d": joints,
        "vertices": verts,
        "j3d.cam": j3d_cam,
        "v3d.cam": v3d_cam,
        "j2d.norm": j2d,
        "beta": shape,
        "pose": pose_in.clone(),
    }


# --------------------------------------------------------------------------------------
# key-point losses and metrics on the head's outputs
# --------------------------------------------------------------------------------------
def keypoint_losses(j3d, j2d, gt3, gt2, joints_valid, gate3=None, gate2=None):
    """src/callbacks/loss/loss_arctic_sf.py:70-92,131-136 + src/utils/loss_modules.py:62-73 (keypoint_3d_loss),
    :88-95 (hand_kp3d_loss), :116-125 (joints_loss): MSE, root-relative for 3D, masked by joints_valid, gated per
    sample, reduced with .mean()."""
    B = j3d.shape[0]
    d3 = ((j3d - j3d[:, :1]) - (gt3 - gt3[:, :1])) ** 2 * joints_valid[:, :, None]
    d2 = (j2d - gt2) ** 2 * joints_valid[:, :, None]
    d3, d2 = d3.reshape(B, -1), d2.reshape(B, -1)
    if gate3 is not None:
        d3 = d3 * gate3[..., None]
    if gate2 is not None:
        d2 = d2 * gate2[..., None]
    return d3.mean(), d2.mean()


def keypoint_metric_sums(j3d, j2d, gt3, gt2, joints_valid, hand_valid, img_res):
    """common/metrics.py:23-45 as called from src/utils/eval_modules.py:95-118 (root-relative, per-hand validity, mean over
    joints) and :407-421 (pixel error on data_utils.unormalize_kp2d'ed key-points, per-joint validity x hand validity):
    numerators and counts of the nan-means the reference takes afterwards."""
    dist3 = (((gt3 - gt3[:, :1]) - (j3d - j3d[:, :1])) ** 2).sum(dim=2).sqrt().mean(dim=1)
    half = 0.5 * img_res
    dist2 = (((half * (gt2 + 1)) - (half * (j2d + 1))) ** 2).sum(dim=2).sqrt()
    v2 = joints_valid * hand_valid[:, None]
    return (dist3 * hand_valid).sum(), hand_valid.sum(), (dist2 * v2).sum(), v2.sum()


def vector_loss_term(pred, gt, valid=None, gate=None, pred2=None):
    """One masked vector MSE term of compute_loss_light: src/utils/loss_modules.py:99-113 (vector_loss, MSE, return_mean=False:
    dist.reshape(B,-1) * is_valid[...,None]; zeros when is_valid.sum()==0, which is the same number), the optional second
    prediction against the same target (`cam_t.wp.init`, loss_arctic_sf.py:116-129), the per-sample gate of :134-145 and the
    .mean() of :146-158."""
    B = pred.shape[0]
    d = ((pred - gt) ** 2).reshape(B, -1)
    if pred2 is not None:
        d = d + ((pred2 - gt) ** 2).reshape(B, -1)
    if valid is not None:
        d = d * valid[..., None]
    if gate is not None:
        d = d * gate[..., None]
    return d.mean()


def axis_angle_to_matrix(aa):
    """pytorch3d axis_angle_to_matrix (loss_arctic_sf.py:48-49) = common/rot.py:754-784 (axis_angle_to_quaternion) then
    :86-115 (quaternion_to_matrix)."""
    ang = torch.norm(aa, p=2, dim=-1, keepdim=True)
    half = ang * 0.5
    small = ang.abs() < 1e-6
    soa = torch.where(small, 0.5 - ang * ang / 48, torch.sin(half) / torch.where(small, torch.ones_like(ang), ang))
    q = torch.cat([torch.cos(half), aa * soa], dim=-1)
    r, i, j, k = torch.unbind(q, -1)
    two_s = 2.0 / (q * q).sum(-1)
    o = torch.stack([1 - two_s * (j * j + k * k), two_s * (i * j - k * r), two_s * (i * k + j * r),
                     two_s * (i * j + k * r), 1 - two_s * (i * i + k * k), two_s * (j * k - i * r),
                     two_s * (i * k - j * r), two_s * (j * k + i * r), 1 - two_s * (i * i + j * j)], -1)
    return o.reshape(aa.shape[:-1] + (3, 3))


def mrrpe_sums(root_r, root_l, gt_root_r, gt_root_l, valid):
    """common/metrics.py:47-55."""
    d = (((root_l - root_r) - (gt_root_l - gt_root_r)) ** 2).sum(dim=1).sqrt()
    return (d * valid).sum(), valid.sum()


# --------------------------------------------------------------------------------------
# GT side of a step      (src/callbacks/process/process_arctic.py:4-75)
# --------------------------------------------------------------------------------------
def process_gt_side(buf, pose, betas, j3d_full, K, img_res):
    """One hand side of process_data_light: canonical MANO forward of the GT parameters (:16-21), mean-offset translation
    into camera space (:42-46), GT camera translation (:49-52) and its weak-perspective form (:58-65)."""
    verts, joints = mano_forward(buf, betas, pose[:, :3], pose[:, 3:])
    Tr0 = (j3d_full - joints).mean(dim=1)
    cam_t = j3d_full[:, 0] - joints[:, 0]
    f = (K[:, 0, 0] + K[:, 1, 1]) / 2.0
    return {"joints3d": joints, "vertices": verts, "v3d.cam": verts + Tr0[:, None, :], "cam_t": cam_t,
            "cam_t.wp": perspective_to_weak_perspective(cam_t, f, img_res)}


# --------------------------------------------------------------------------------------
# KPE features      (src/datasets/hands_light_dataset.py:259-279, src/models/hands_light/model.py:444-460)
# --------------------------------------------------------------------------------------
def kpe_angles(bbox, K):
    """Per box (numpy float64 arctan2, cast to float32): centre angles (2) and corner angles (8)."""
    import numpy as np

    bbox, K = np.asarray(bbox, dtype=np.float64), np.asarray(K, dtype=np.float64)
    center, corner = [], []
    for b, k in zip(bbox, K):
        c = (b[:2] + b[2:]) / 2.0
        center.append(np.array([np.arctan2(c[0] - k[0, 2], k[0, 0]), np.arctan2(c[1] - k[1, 2], k[1, 1])]).astype(np.float32))
        cr = np.array([[b[0], b[1]], [b[0], b[3]], [b[2], b[1]], [b[2], b[3]]])
        cr = np.stack([cr[:, 0] - k[0, 2], cr[:, 1] - k[1, 2]], axis=-1)
        corner.append(np.arctan2(cr, np.array([[k[0, 0], k[1, 1]]])).flatten().astype(np.float32))
    return torch.from_numpy(np.stack(center)), torch.from_numpy(np.stack(corner))


def kpe_pos_enc(angle, L):
    """model.py:444-460 (compute_center_pos_enc / compute_corner_pos_enc are the same arithmetic)."""
    bz, c = angle.shape
    freq_expand = 2 ** torch.arange(L).unsqueeze(0).repeat(bz, 1).reshape(bz, -1, 1)
    angle_expand = angle.reshape(bz, 1, c)
    return torch.stack([torch.sin(freq_expand * angle_expand), torch.cos(freq_expand * angle_expand)], dim=-1).reshape(bz, -1).float()


# --------------------------------------------------------------------------------------
# Perspective Crop Layer      (src/datasets/hands_light_dataset.py:354-467)
# --------------------------------------------------------------------------------------
def virtual_camera_rotation(p):
    """hands_light_dataset.py:357-366.  p: 3 python floats / float64."""
    x, y = float(p[0]), float(p[1])
    n1x = math.sqrt(1 + x * x)
    d1x = 1 / n1x
    d1xy = 1 / math.sqrt(1 + x * x + y * y)
    d1xy1x = 1 / math.sqrt((1 + x * x + y * y) * (1 + x * x))
    return torch.tensor(
        [
            [d1x, -x * y * d1xy1x, x * d1xy],
            [0.0, n1x * d1xy, y * d1xy],
            [-x * d1x, -y * d1xy1x, d1xy],
        ],
        dtype=torch.float64,
    )


def virtual_intrinsics(p, K, size_wh):
    """hands_light_dataset.py:368-386 (focal_at_image_plane and slant_compensation on)."""
    p = [float(v) for v in p]
    plen = math.sqrt(p[0] ** 2 + p[1] ** 2 + p[2] ** 2)
    sx = 1.0 / math.sqrt(p[0] ** 2 + p[2] ** 2)
    sy = math.sqrt(p[0] ** 2 + 1) / math.sqrt(p[0] ** 2 + p[1] ** 2 + 1)
    Kv = torch.zeros(3, 3, dtype=torch.float64)
    Kv[0, 0] = plen * float(K[0, 0]) / (size_wh[0] * sx)
    Kv[1, 1] = plen * float(K[1, 1]) / (size_wh[1] * sy)
    Kv[0, 2] = 0.5
    Kv[1, 2] = 0.5
    Kv[2, 2] = 1.0
    return Kv


def pcl_homography(bbox, K, img_res=224):
    """hands_light_dataset.py:425-454: bbox [x0,y0,x1,y1] + intrinsics -> (P_virt2orig fp32 (3,3),
    R_virt2orig fp32 (3,3), s int).  All arithmetic in float64, cast at the end (:443-445)."""
    K64 = K.to(torch.float64)
    x0, y0, x1, y1 = [float(v) for v in bbox]
    cx, cy = (x0 + x1) / 2, (y0 + y1) / 2
    w, h = x1 - x0, y1 - y0
    s = int(max(w, h))
    if s == 0:
        s = int(img_res)
    p = torch.linalg.inv(K64) @ torch.tensor([cx, cy, 1.0], dtype=torch.float64)
    R = virtual_camera_rotation(p)
    Kv = virtual_intrinsics(p, K64, (s, s))
    P = K64 @ (R @ torch.linalg.inv(Kv))
    return P.float(), R.float(), s


def perspective_grid(P, img_res, s):
    """hands_light_dataset.py:388-423 with transform_to_pytorch=True.  Returns (s,s,2) fp32."""
    xs = torch.linspace(0, 1, s)
    ys = torch.linspace(0, 1, s)
    rs, cs = torch.meshgrid([xs, ys], indexing="ij")
    pv = torch.stack([rs, cs, torch.ones_like(rs)]).reshape(3, -1)
    q = torch.matmul(P, pv)
    q = q[:2] / (1e-8 + q[2:3])
    g = q.reshape(2, s, s).permute(2, 1, 0).clone()
    g /= img_res
    g *= 2
    g -= 1
    return g


def perspective_crop(img, bbox, K, img_res=224):
    """Batched restatement of hands_light_dataset.py:425-467 (one crop per row).
    img (B,3,H,W) fp32, bbox (B,4) int, K (B,3,3).  Returns crop (B,3,img_res,img_res), R (B,3,3).
    Differentiable w.r.t. img (the reference runs this in the data loader, forward only)."""
    crops, rots = [], []
    for b in range(img.shape[0]):
        P, R, s = pcl_homography(bbox[b].tolist(), K[b], img_res)
        grid = perspective_grid(P, img_res, s)
        mid = F.grid_sample(img[b : b + 1], grid[None], mode="bilinear", padding_mode="zeros", align_corners=False)
        out = F.interpolate(mid, size=(img_res, img_res), mode="bilinear", align_corners=True)
        crops.append(out[0])
        rots.append(R)
    return torch.stack(crops), torch.stack(rots)


def pcl_fix_global_orient(R_virt2orig, pose):
    """src/models/hands_light/model.py:330-334: pose[:,0] <- R_virt2orig @ pose[:,0]."""
    out = pose.clone()
    out[:, 0] = torch.bmm(R_virt2orig, pose[:, 0])
    return out


# --------------------------------------------------------------------------------------
# mesh index constants     (common/body_models.py:35-72)
# --------------------------------------------------------------------------------------
SEAL_RING = (120, 108, 79, 78, 121, 214, 215, 279, 239, 234, 92, 38, 122, 118, 117, 119)
CIRCLE_V_ID = (108, 79, 78, 121, 214, 215, 279, 239, 234, 92, 38, 122, 118, 117, 119, 120)


def seal_faces(is_rhand):
    """16 wrist-sealing triangles fanning around the extra centre vertex 778
    (common/body_models.py:35-52; flipped winding for the left hand :66-68)."""
    ring = list(SEAL_RING)
    tris = [[ring[i], ring[(i + 1) % 16], 778] for i in range(16)]
    t = torch.tensor(tris, dtype=torch.int64)
    if not is_rhand:
        t = t[:, [1, 0, 2]]
    return t


def seal_mano_mesh(v3d, faces, is_rhand):
    """common/body_models.py:60-72."""
    centre = v3d[:, list(CIRCLE_V_ID)].mean(dim=1, keepdim=True)
    return torch.cat([v3d, centre], dim=1), torch.cat([faces, seal_faces(is_rhand).to(faces.device)], dim=0)
