"""Self-checks of the smplx-MANO restatement (parity unpinned by the reference: SURVEY.md §8(c)).
Known answers and invariances that any correct MANO forward must satisfy."""
import math

import torch

from hands_b200.synthetic import PARENTS, TIP_IDS, synthetic_head_inputs, synthetic_mano_buffers
from oracle import geometry_oracle as O


def _aa_inputs(B, seed, dtype=torch.float64):
    g = torch.Generator().manual_seed(seed)
    pose = torch.randn(B, 48, generator=g, dtype=dtype) * 0.3
    betas = torch.randn(B, 10, generator=g, dtype=dtype)
    return pose, betas


def test_index_constants_bit_exact():
    buf = synthetic_mano_buffers(True)
    assert buf["parents"].tolist() == [-1, 0, 1, 2, 0, 4, 5, 0, 7, 8, 0, 10, 11, 0, 13, 14] == PARENTS
    assert list(O.TIP_IDS) == [744, 320, 443, 554, 671] == TIP_IDS
    assert O.seal_faces(True).shape == (16, 3) and O.seal_faces(True)[0].tolist() == [120, 108, 778]
    assert O.seal_faces(False)[0].tolist() == [108, 120, 778]
    assert O.seal_faces(True)[-1].tolist() == [119, 120, 778]
    v = torch.randn(2, 778, 3)
    sv, sf = O.seal_mano_mesh(v, buf["faces"], True)
    assert sv.shape == (2, 779, 3) and sf.shape == (1554, 3)
    torch.testing.assert_close(sv[:, 778], v[:, list(O.CIRCLE_V_ID)].mean(1))


def test_rest_pose_identity():
    buf = synthetic_mano_buffers(True, flat_hand=True)
    B = 3
    z = torch.zeros(B, 48, dtype=torch.float64)
    verts, joints = O.mano_forward(buf, torch.zeros(B, 10, dtype=torch.float64), z[:, :3], z[:, 3:])
    vt = buf["v_template"].double()
    torch.testing.assert_close(verts, vt.expand(B, -1, -1), rtol=0, atol=1e-8)  # the +1e-8 in Rodrigues leaves a 1.7e-8 rad rotation
    torch.testing.assert_close(joints[:, :16], (buf["J_regressor"].double() @ vt).expand(B, -1, -1), rtol=0, atol=1e-8)  # the +1e-8 in Rodrigues leaves a 1.7e-8 rad rotation
    torch.testing.assert_close(joints[:, 16:], vt[list(O.TIP_IDS)].expand(B, -1, -1), rtol=0, atol=1e-8)  # the +1e-8 in Rodrigues leaves a 1.7e-8 rad rotation
    assert joints.shape == (B, 21, 3)


def test_translation_and_global_rotation_equivariance():
    buf = synthetic_mano_buffers(False)
    pose, betas = _aa_inputs(4, 1)
    v0, j0 = O.mano_forward(buf, betas, pose[:, :3], pose[:, 3:])
    t = torch.randn(4, 3, dtype=torch.float64)
    v1, j1 = O.mano_forward(buf, betas, pose[:, :3], pose[:, 3:], transl=t)
    torch.testing.assert_close(v1, v0 + t[:, None])
    torch.testing.assert_close(j1, j0 + t[:, None])
    # a different global orientation rotates the whole hand rigidly about the root joint
    R0 = O.batch_rodrigues(pose[:, :3] + buf["pose_mean"][:3].double())
    g2 = torch.randn(4, 3, dtype=torch.float64) * 0.7
    R1 = O.batch_rodrigues(g2)
    v2, j2 = O.mano_forward(buf, betas, g2, pose[:, 3:])
    root = j0[:, :1]
    rel = R1 @ R0.transpose(1, 2)
    torch.testing.assert_close(v2 - root, (v0 - root) @ rel.transpose(1, 2), atol=1e-9, rtol=1e-7)
    torch.testing.assert_close(j2 - root, (j0 - root) @ rel.transpose(1, 2), atol=1e-9, rtol=1e-7)


def test_left_right_use_their_own_buffers():
    pose, betas = _aa_inputs(2, 2)
    vr, _ = O.mano_forward(synthetic_mano_buffers(True), betas, pose[:, :3], pose[:, 3:])
    vl, _ = O.mano_forward(synthetic_mano_buffers(False), betas, pose[:, :3], pose[:, 3:])
    assert (vr - vl).abs().max() > 1e-3


def test_rotation_roundtrip_and_small_angle_branch():
    rotmat, _, _, _ = synthetic_head_inputs(64, seed=0)
    R = rotmat.reshape(-1, 3, 3)
    aa = O.matrix_to_axis_angle(R)
    back = O.batch_rodrigues(aa)
    assert (back - R).abs().max() < 4e-6
    eye = torch.eye(3).expand(5, 3, 3)
    assert O.matrix_to_axis_angle(eye).abs().max() == 0.0
    # angle near pi: a non-w quaternion candidate is chosen
    rot_pi, _, _, _ = synthetic_head_inputs(4, seed=1, edge="near_pi")
    R = rot_pi.reshape(-1, 3, 3).double()
    aa = O.matrix_to_axis_angle(R)
    ang = aa.norm(dim=1)
    # no quaternion standardisation on this path, so the angle may come out as theta or 2pi-theta
    assert (torch.minimum((ang - (math.pi - 1e-3)).abs(), (ang - (math.pi + 1e-3)).abs()) < 1e-6).all()
    assert (O.batch_rodrigues(aa) - R).abs().max() < 1e-6  # inputs are fp32-rounded matrices


def test_logmap_backward_is_tangent_projection_not_identity():
    rotmat, _, _, _ = synthetic_head_inputs(2, seed=3)
    R = rotmat.reshape(-1, 3, 3).double().requires_grad_(True)
    out = O.batch_rodrigues(O.matrix_to_axis_angle(R))
    w = torch.randn(out.shape, dtype=torch.float64, generator=torch.Generator().manual_seed(0))
    (g,) = torch.autograd.grad((out * w).sum(), R)
    assert ((g - w).norm() / w.norm()) > 0.3


def test_head_gradcheck_fp64():
    buf = synthetic_mano_buffers(True)
    rotmat, betas, cam, K = synthetic_head_inputs(2, seed=5)
    rotmat, betas, cam, K = rotmat.double().requires_grad_(True), betas.double().requires_grad_(True), cam.double().requires_grad_(True), K.double()

    def fn(r, b, c):
        o = O.mano_head_forward(buf, r, b, c, K)
        return o["v3d.cam"][:, ::97], o["j3d.cam"], o["j2d.norm"]

    assert torch.autograd.gradcheck(fn, (rotmat, betas, cam), eps=1e-6, atol=1e-5, rtol=1e-4, nondet_tol=0)


def test_small_scale_clamp_zero_gradient():
    buf = synthetic_mano_buffers(True)
    rotmat, betas, cam, K = synthetic_head_inputs(4, seed=6, small_s_frac=0.5)
    cam = cam.clone().requires_grad_(True)
    o = O.mano_head_forward(buf, rotmat, betas, cam, K)
    o["j2d.norm"].sum().backward()
    assert (cam.grad[:2, 0] == 0).all() and (cam.grad[2:, 0] != 0).all()
    f = (K[:, 0, 0] + K[:, 1, 1]) / 2
    torch.testing.assert_close(o["cam_t"][:2, 2], 2 * f[:2] / (224.0 * 0.1 + 1e-9))
