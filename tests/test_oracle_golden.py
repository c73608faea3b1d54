"""The oracle against the fixtures produced by the reference's own code (tests/golden/make_golden.py)."""
import os

import numpy as np
import torch

import pytest

from hands_b200.synthetic import synthetic_head_inputs, synthetic_mano_buffers, synthetic_pcl_inputs
from oracle import geometry_oracle as O


def _load(golden_dir, name):
    return {k: v for k, v in np.load(os.path.join(golden_dir, name)).items()}


def test_logmap_matches_reference(golden_dir):
    g = _load(golden_dir, "logmap.npz")
    R = torch.from_numpy(g["R"]).requires_grad_(True)
    quat = O.matrix_to_quaternion(R)
    aa = O.matrix_to_axis_angle(R)
    assert np.array_equal(quat.detach().numpy(), g["quat"])          # same ops, same order -> bit-exact
    assert np.array_equal(aa.detach().numpy(), g["aa"])
    (gR,) = torch.autograd.grad((aa * torch.from_numpy(g["w"])).sum(), R)
    np.testing.assert_allclose(gR.numpy(), g["gR"], rtol=1e-6, atol=1e-6)


def test_camera_projection_matches_reference(golden_dir):
    g = _load(golden_dir, "camera_projection.npz")
    cam = torch.from_numpy(g["cam"]).requires_grad_(True)
    f = torch.from_numpy(g["f"])
    K = torch.from_numpy(g["K"])
    cam_t = O.weak_perspective_to_perspective(cam, f, 224, 0.1)
    assert np.array_equal(cam_t.detach().numpy(), g["cam_t"])
    (g_cam,) = torch.autograd.grad((cam_t * torch.from_numpy(g["w3"])).sum(), cam)
    np.testing.assert_allclose(g_cam.numpy(), g["g_cam"], rtol=1e-6, atol=0)
    assert np.array_equal(O.perspective_to_weak_perspective(cam_t.detach(), f, 224).numpy(), g["wp"])
    pts = torch.from_numpy(g["pts"]).requires_grad_(True)
    j2d = O.project2d_batch(K, pts)
    assert np.array_equal(j2d.detach().numpy(), g["j2d"])
    j2d_n = O.normalize_kp2d(j2d, 224)
    assert np.array_equal(j2d_n.detach().numpy(), g["j2d_norm"])
    (g_pts,) = torch.autograd.grad((j2d_n * torch.from_numpy(g["w2"])).sum(), pts)
    np.testing.assert_allclose(g_pts.numpy(), g["g_pts"], rtol=1e-6, atol=1e-7)
    assert np.array_equal(O.unormalize_kp2d(j2d_n.detach(), 224).numpy(), g["j2d_un"])


def test_pcl_small_matches_reference(golden_dir):
    g = _load(golden_dir, "pcl.npz")
    img = torch.from_numpy(g["small_img"])
    for j in range(8):
        b = j // 2
        K = torch.from_numpy(g["small_K"][b])
        P, R, s = O.pcl_homography(g["small_bbox"][j].tolist(), K, 64)
        assert s == int(g["small_s"][j])
        assert np.array_equal(R.numpy(), g["small_rot"][j])
        grid = O.perspective_grid(P, 64, s)
        assert np.array_equal(grid.numpy(), g[f"grid_{j}"])
        crop, R2 = O.perspective_crop(img[b : b + 1], torch.from_numpy(g["small_bbox"][j : j + 1]), K[None].float(), 64)
        assert np.array_equal(crop[0].numpy(), g["small_crop"][j])


def test_pcl_full_matches_reference(golden_dir):
    g = _load(golden_dir, "pcl.npz")
    img, bbox, K = synthetic_pcl_inputs(4, seed=int(g["full_seed"]), img_res=224)
    # torch's own CPU bilinear kernels differ in the last ulp between 1 and N threads (measured
    # 2.4e-7 on N(0,1) images), so bit-equality is asserted single-threaded like the generator ran.
    nt = torch.get_num_threads()
    try:
        for threads, exact in ((1, True), (nt, False)):
            torch.set_num_threads(threads)
            for j in range(4):
                b = (j // 2) * 2
                crop, R = O.perspective_crop(img[b : b + 1], bbox[j : j + 1], K[b : b + 1], 224)
                got = crop[0, :, ::4, ::4].numpy()
                if exact:
                    assert np.array_equal(got, g["full_crop_sub4"][j])
                else:
                    np.testing.assert_allclose(got, g["full_crop_sub4"][j], rtol=0, atol=1e-6)
                assert np.array_equal(R[0].numpy(), g["full_rot"][j])
    finally:
        torch.set_num_threads(nt)


def test_rot6d_matches_reference(golden_dir):
    """The three 6D -> rotation-matrix conversions against the reference's own outputs (common/rot.py:367-381,
    hamer_light/geometry.py:47-62, handoccnet_light/mano_head.py:132-141)."""
    d = np.load(os.path.join(golden_dir, "rot6d.npz"))
    x = torch.from_numpy(d["x"])
    for name, fn in (("paired", O.rot6d_to_rotmat_paired), ("cols", O.rot6d_to_rotmat_cols)):
        xi = x.clone().requires_grad_(True)
        R = fn(xi)
        assert torch.equal(R.detach(), torch.from_numpy(d[f"R_{name}"]))
        (gx,) = torch.autograd.grad((R * torch.from_numpy(d[f"w_{name}"])).sum(), xi)
        ref = torch.from_numpy(d[f"gx_{name}"])
        assert float((gx - ref).abs().max() / ref.abs().max()) <= 1e-6
    assert torch.equal(O.rot6d2mat(x), torch.from_numpy(d["R_handoccnet"]))
    # the pytorch3d variant (absent here) is the row-stacked form of the same Gram-Schmidt
    assert torch.equal(O.rotation_6d_to_matrix(x), O.rot6d2mat(x).transpose(1, 2))
    Rm = O.rotation_6d_to_matrix(x[:48].double())
    eye = torch.eye(3, dtype=torch.float64).expand_as(Rm)
    assert float((Rm @ Rm.transpose(1, 2) - eye).abs().max()) < 1e-12 and float((torch.linalg.det(Rm) - 1).abs().max()) < 1e-12


def test_keypoint_losses_and_metrics_match_reference(golden_dir):
    """Loss terms / metrics on the path's outputs against the reference's own functions (loss_modules.py, metrics.py)."""
    d = {k: torch.from_numpy(np.asarray(v)) for k, v in np.load(os.path.join(golden_dir, "kp_loss.npz")).items()}
    j3d, j2d = d["j3d"].clone().requires_grad_(True), d["j2d"].clone().requires_grad_(True)
    l3, l2 = O.keypoint_losses(j3d, j2d, d["gt3"], d["gt2"], d["jv"], d["gate3"], d["gate2"])
    assert abs(float(l3.detach()) - float(d["loss3"])) <= 1e-6 * float(d["loss3"]) and abs(float(l2.detach()) - float(d["loss2"])) <= 1e-6 * float(d["loss2"])
    g3, g2 = torch.autograd.grad(5.0 * l3 + 3.0 * l2, (j3d, j2d))
    assert float((g3 - d["g3"]).abs().max()) <= 1e-6 * float(d["g3"].abs().max()) and float((g2 - d["g2"]).abs().max()) <= 1e-6 * float(d["g2"].abs().max())
    s2, n2, s4, n4 = O.keypoint_metric_sums(d["j3d"], d["j2d"], d["gt3"], d["gt2"], d["jv"], d["hv"], 224.0)
    assert abs(float(s2 / n2) - float(np.nanmean(d["mpjpe"].numpy()))) <= 1e-6
    assert abs(float(s4 / n4) - float(np.nanmean(d["pix"].numpy()))) <= 1e-4
    assert int(n2) == int(np.isfinite(d["mpjpe"].numpy()).sum()) and int(n4) == int(np.isfinite(d["pix"].numpy()).sum())
    sm, nm = O.mrrpe_sums(d["j3d"][:, 0], d["j3d_l"][:, 0], d["gt3"][:, 0], d["gt3_l"][:, 0], d["hv"])
    assert abs(float(sm / nm) - float(np.nanmean(d["mrrpe"].numpy()))) <= 1e-6


def test_process_gt_matches_reference(golden_dir):
    """GT side of a step: the oracle against the reference's own process_data_light (process_arctic.py:4-75) run on the
    same MANO restatement -- pins the glue (mean-offset translation, GT camera translation, weak-perspective camera)."""
    from hands_b200.synthetic import synthetic_mano_buffers

    d = {k: torch.from_numpy(v) for k, v in np.load(os.path.join(golden_dir, "process_gt.npz")).items()}
    for side, is_rhand in (("r", True), ("l", False)):
        o = O.process_gt_side(synthetic_mano_buffers(is_rhand), d[f"in_mano.pose.{side}"], d[f"in_mano.beta.{side}"], d[f"in_mano.j3d.full.{side}"], d["K"], 224)
        for key, name in (("joints3d", "joints3d"), ("vertices", "vertices"), ("v3d.cam", "v3d.cam"), ("cam_t", "cam_t"), ("cam_t.wp", "cam_t.wp")):
            assert torch.equal(o[key], d[f"out_mano.{name}.{side}"]), (side, key)
        assert torch.equal(d[f"out_mano.j3d.cam.{side}"], d[f"in_mano.j3d.full.{side}"])


def test_kpe_features_match_reference(golden_dir):
    """KPE angles and sinusoidal encodings against the reference's own lines (hands_light_dataset.py:259-279, model.py:444-460)."""
    d = np.load(os.path.join(golden_dir, "kpe.npz"))
    center, corner = O.kpe_angles(d["bbox"], d["K"])
    assert np.array_equal(center.numpy(), d["center"]) and np.array_equal(corner.numpy(), d["corner"])
    L = int(d["L"])
    assert torch.equal(O.kpe_pos_enc(center, L), torch.from_numpy(d["center_enc"]))
    assert torch.equal(O.kpe_pos_enc(corner, L), torch.from_numpy(d["corner_enc"]))


def test_mesh_constants_match_reference(golden_dir):
    """a21: the PRODUCT's wrist-seal table and `seal_mano_mesh` (hands_b200/common/body_models.py) and the oracle's, against
    the reference's own table and function (common/body_models.py:35-72, exec'd by make_golden.py).  Bit-exact: indices."""
    from hands_b200.common import body_models as P

    g = _load(golden_dir, "mesh_constants.npz")
    assert np.array_equal(np.asarray(P.SEAL_FACES_R, dtype=np.int64), g["SEAL_FACES_R"])
    assert np.array_equal(np.asarray(P.CIRCLE_V_ID, dtype=np.int64), g["CIRCLE_V_ID"])
    v3d, faces = torch.from_numpy(g["v3d"]), torch.from_numpy(g["faces"])
    for name, is_rhand in (("r", True), ("l", False)):
        for impl in (P.seal_mano_mesh, O.seal_mano_mesh):
            sv, sf = impl(v3d, faces, is_rhand)
            assert sv.shape == (3, 779, 3) and sf.shape == (1554, 3) and sf.dtype == torch.int64
            assert torch.equal(sf, torch.from_numpy(g[f"sealed_f_{name}"]))
            assert torch.equal(sv, torch.from_numpy(g[f"sealed_v_{name}"]))


def test_mano_head_chain_matches_reference_source(golden_dir):
    """a11: the oracle's head chain against the reference's OWN MANOHead.forward source (mano_head.py:21-65) run over the same
    MANO layer (make_golden.py::golden_mano_head_reference_source) -- outputs and gradients to 1e-6 relative."""
    g = _load(golden_dir, "mano_head_ref.npz")
    for name, is_rhand in (("r", True), ("l", False)):
        rotmat, betas, cam, K = synthetic_head_inputs(12, seed=int(g[f"seed_{name}"]), small_s_frac=0.25)
        buf = synthetic_mano_buffers(is_rhand)
        r, b, c = rotmat.clone().requires_grad_(True), betas.clone().requires_grad_(True), cam.clone().requires_grad_(True)
        out = O.mano_head_forward(buf, r, b, c, K, 224.0, 0.1)
        for k in ["cam_t", "joints3d", "vertices", "j3d.cam", "v3d.cam", "j2d.norm", "pose", "beta", "cam_t.wp"]:
            ref = torch.from_numpy(g[f"{k}_{name}"])   # (torch's matmul blocking depends on the thread count: not always bit-equal)
            assert float((out[k].detach() - ref).abs().max()) <= 1e-6 * float(ref.abs().max()), k
        loss = sum((out[k] * torch.from_numpy(g[f"w_{k}_{name}"])).sum() for k in ("v3d.cam", "j3d.cam", "j2d.norm"))
        grads = torch.autograd.grad(loss, (r, b, c))
        for gname, got in zip(("g_rotmat", "g_betas", "g_cam"), grads):
            ref = torch.from_numpy(g[f"{gname}_{name}"])
            assert float((got - ref).abs().max() / ref.abs().max()) <= 1e-6, gname
        aa = torch.from_numpy(g[f"aa_{name}"])
        out2 = O.mano_head_forward(buf, aa, betas, cam, K, 224.0, 0.1)
        assert float((out2["j2d.norm"] - torch.from_numpy(g[f"aa_j2d.norm_{name}"])).abs().max()) <= 1e-6
        assert float((out2["v3d.cam"] - torch.from_numpy(g[f"aa_v3d.cam_{name}"])).abs().max()) <= 1e-6 * float(out2["v3d.cam"].abs().max())


def test_mano_layer_state_dict_is_smplx_shaped():
    """Strict checkpoint exchange with the reference (common/abstract_pl.py:42-44 loads strictly): the drop-in's persistent
    keys are the smplx.MANO(use_pca=False) set [smplx-recalled]; helper tensors that differ between smplx versions are tolerated;
    anything else is still an error.  Copies / pickles of a module never carry the device handle cache."""
    import copy
    import pickle

    from hands_b200.common.body_models import build_mano_aa

    m = build_mano_aa(True, synthetic=True)
    assert sorted(m.state_dict().keys()) == sorted([
        "J_regressor", "betas", "faces_tensor", "global_orient", "hand_mean", "hand_pose", "lbs_weights", "parents", "pose_mean",
        "posedirs", "shapedirs", "v_template", "vertex_joint_selector.extra_joints_idxs"])
    assert m.tip_ids.tolist() == [744, 320, 443, 554, 671]
    sd = dict(m.state_dict())
    sd["body_pose"] = torch.zeros(1, 3)
    del sd["hand_mean"]
    m2 = copy.deepcopy(m)
    m2.load_state_dict(sd, strict=True)
    sd["not_a_mano_key"] = torch.zeros(1)
    with pytest.raises(RuntimeError):
        m2.load_state_dict(sd, strict=True)
    m._handles[("cuda", 0)] = ("stamp", object())     # stand-in for a live device handle
    assert pickle.loads(pickle.dumps(m))._handles == {} and copy.deepcopy(m)._handles == {}


def test_loss_light_terms_match_reference(golden_dir):
    """f2: the oracle's masked vector-MSE term and axis_angle_to_matrix against the reference's OWN compute_loss_light
    (loss_arctic_sf.py:20-158, exec'd from the file by make_golden.py) on every cam_t / transl / pose / beta key."""
    g = _load(golden_dir, "loss_light.npz")
    P = {k[5:]: torch.from_numpy(v) for k, v in g.items() if k.startswith("pred:")}
    G = {k[3:]: torch.from_numpy(v) for k, v in g.items() if k.startswith("gt:")}
    M = {k[5:]: torch.from_numpy(v) for k, v in g.items() if k.startswith("meta:")}
    close = lambda a, key: abs(float(a) - float(g["loss:" + key])) <= 2e-6 * max(1.0, abs(float(g["loss:" + key])))  # noqa: E731
    assert float((O.axis_angle_to_matrix(G["mano.pose.r"].reshape(-1, 3)) - torch.from_numpy(g["gt_rotmat_r"])).abs().max()) <= 2e-7
    for sd, valid in (("r", G["right_valid"]), ("l", G["left_valid"])):
        assert close(O.vector_loss_term(P[f"mano.cam_t.wp.{sd}"], G[f"mano.cam_t.wp.{sd}"], valid, M["is_cam_loss"], pred2=P[f"mano.cam_t.wp.init.{sd}"]), f"loss/mano/cam_t/{sd}")
        gt_pose = O.axis_angle_to_matrix(G[f"mano.pose.{sd}"].reshape(-1, 3)).reshape(-1, 16, 3, 3)
        assert close(O.vector_loss_term(P[f"mano.pose.{sd}"], gt_pose, valid, M["is_pose_loss"]), f"loss/mano/pose/{sd}")
        assert close(O.vector_loss_term(P[f"mano.beta.{sd}"], G[f"mano.beta.{sd}"], valid, M["is_beta_loss"]), f"loss/mano/beta/{sd}")
        l3, l2 = O.keypoint_losses(P[f"mano.j3d.cam.{sd}"], P[f"mano.j2d.norm.{sd}"], G[f"mano.j3d.cam.{sd}"], G[f"mano.j2d.norm.{sd}"], G[f"joints_valid_{sd}"],
                                   M["is_j3d_loss"], M["is_j2d_loss"])
        assert close(l3, f"loss/mano/kp3d/{sd}") and close(l2, f"loss/mano/kp2d/{sd}")
    assert close(O.vector_loss_term(P["mano.cam_t.wp.l"] - P["mano.cam_t.wp.r"], G["mano.cam_t.wp.l"] - G["mano.cam_t.wp.r"], G["right_valid"] * G["left_valid"], M["is_cam_loss"]),
                 "loss/mano/transl/l")


def test_mano_decimator_matches_reference(golden_dir):
    """f4: the drop-in MANODecimator (one GEMM over the (778, 3B) view) against the reference's own class
    (common/body_models.py:11-32, exec'd by make_golden.py with a synthetic decimation matrix)."""
    from hands_b200.common.body_models import MANODecimator

    g = _load(golden_dir, "decimator.npz")
    dec = MANODecimator(data={"D_right": g["D_right"], "D_left": g["D_left"], "other": np.zeros(3)})
    verts = torch.from_numpy(g["verts"])
    for name, flag in (("r", True), ("l", False)):
        out = dec.downsample(verts, flag)
        assert out.shape == (5, 195, 3) and out.is_contiguous()
        assert float((out - torch.from_numpy(g[f"sub_{name}"])).abs().max()) <= 1e-7
