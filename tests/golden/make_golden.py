"""Generate golden fixtures from the REFERENCE's own code (run in the build container only).

    python tests/golden/make_golden.py            # needs /root/reference (read-only)

The reference is Python, so its path-owned functions are imported from where they lie and run
on CPU; the Perspective-Crop-Layer closure (a nested block inside a dataset `__getitem__`,
`src/datasets/hands_light_dataset.py:354-467`) is exec'd from the file's own lines at
generation time -- nothing is copied into this repository.  The resulting .npz files are
committed; tests never read /root/reference.

`smplx` (the MANO arithmetic) cannot be imported (absent), so no golden exists for it: that
part of the oracle is "parity unpinned" (see oracle/geometry_oracle.py header).
"""
import math
import os
import sys
import textwrap
import types

import numpy as np
import torch
import torch.nn.functional as F

REF = os.environ.get("HANDS_REFERENCE", "/root/reference")
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, REF)
sys.path.insert(0, os.path.join(HERE, "..", ".."))

import common.camera as ref_camera  # noqa: E402
import common.data_utils as ref_data_utils  # noqa: E402
import common.rot as ref_rot  # noqa: E402
import common.transforms as ref_tf  # noqa: E402

from hands_b200.synthetic import random_rotmats, synthetic_head_inputs, synthetic_pcl_inputs  # noqa: E402


def golden_logmap():
    g = torch.Generator().manual_seed(7)
    R = torch.cat(
        [
            random_rotmats(40, g),
            random_rotmats(8, g, edge="identity"),
            random_rotmats(16, g, edge="near_pi"),
        ]
    )
    # small rotations around random axes (exercise the small-angle neighbourhood)
    tiny = ref_rot.quaternion_to_matrix(
        torch.nn.functional.normalize(
            torch.cat([torch.ones(8, 1), 1e-4 * torch.randn(8, 3, generator=g)], dim=1), dim=1
        )
    )
    R = torch.cat([R, tiny]).contiguous().requires_grad_(True)
    aa = ref_rot.matrix_to_axis_angle(R)
    w = torch.randn(aa.shape, generator=g)
    (gR,) = torch.autograd.grad((aa * w).sum(), R)
    quat = ref_rot.matrix_to_quaternion(R.detach())
    np.savez_compressed(
        os.path.join(HERE, "logmap.npz"),
        R=R.detach().numpy(), aa=aa.detach().numpy(), quat=quat.numpy(), w=w.numpy(), gR=gR.numpy(),
    )


def golden_rot6d():
    """The reference's three 6D -> rotation-matrix conversions.  common/rot.py, hamer_light/geometry.py and
    handoccnet_light/mano_head.py import and run here; pytorch3d (hand_hmr.py:85-87) is absent, so its variant has no
    fixture of its own (it is the transpose of the column-stacked one)."""
    import importlib.util

    def load(path, name):
        spec = importlib.util.spec_from_file_location(name, os.path.join(REF, path))
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        return mod

    hamer_geometry = load("src/models/hamer_light/geometry.py", "ref_hamer_geometry")
    g = torch.Generator().manual_seed(11)
    x = torch.randn(64, 6, generator=g)
    x[48:56] *= 1e-3        # tiny vectors
    x[56:60, 3:] = x[56:60, :3] * 1.5 + 1e-3 * torch.randn(4, 3, generator=g)   # nearly parallel (contiguous layout)
    out = {"x": x.numpy()}
    for name, fn in (("paired", ref_rot.rot6d_to_rotmat), ("cols", hamer_geometry.rot6d_to_rotmat)):
        xi = x.clone().requires_grad_(True)
        R = fn(xi)
        w = torch.randn(R.shape, generator=g)
        (gx,) = torch.autograd.grad((R * w).sum(), xi)
        out.update({f"R_{name}": R.detach().numpy(), f"w_{name}": w.numpy(), f"gx_{name}": gx.numpy()})
    try:
        occ = load("src/models/handoccnet_light/mano_head.py", "ref_handoccnet_mano_head")
        out["R_handoccnet"] = occ.rot6d2mat(x).numpy()
    except Exception as e:  # noqa: BLE001  (module-level imports of that file may be missing here)
        print("handoccnet rot6d2mat not importable:", type(e).__name__, e)
    np.savez_compressed(os.path.join(HERE, "rot6d.npz"), **out)


def golden_kp_loss():
    """Loss terms and metrics that read the path's outputs, from the reference's own functions
    (src/utils/loss_modules.py, common/metrics.py, common/data_utils.py)."""
    import src.utils.loss_modules as ref_lm
    import common.metrics as ref_metrics

    g = torch.Generator().manual_seed(21)
    B = 37
    j3d = (0.1 * torch.randn(B, 21, 3, generator=g) + torch.tensor([0.0, 0.0, 0.6])).requires_grad_(True)
    gt3 = 0.1 * torch.randn(B, 21, 3, generator=g) + torch.tensor([0.0, 0.0, 0.6])
    j2d = (0.5 * torch.randn(B, 21, 2, generator=g)).requires_grad_(True)
    gt2 = 0.5 * torch.randn(B, 21, 2, generator=g)
    jv = (torch.rand(B, 21, generator=g) > 0.2).float()
    hv = (torch.rand(B, generator=g) > 0.15).float()
    gate3 = (torch.rand(B, generator=g) > 0.3).float()
    gate2 = (torch.rand(B, generator=g) > 0.3).float()
    mse = torch.nn.MSELoss(reduction="none")
    l3 = ref_lm.hand_kp3d_loss(j3d, gt3, mse, jv, return_mean=False)                 # loss_arctic_sf.py:87-92
    l2 = ref_lm.joints_loss(j2d, gt2, criterion=mse, jts_valid=jv, return_mean=False)  # :70-83
    l3 = (l3.reshape(B, -1) * gate3[..., None]).mean()                               # :131-136, :146-158
    l2 = (l2.reshape(B, -1) * gate2[..., None]).mean()
    g3, g2 = torch.autograd.grad(5.0 * l3 + 3.0 * l2, (j3d, j2d))
    with torch.no_grad():
        ra = lambda t: t - t[:, :1, :]  # noqa: E731  (eval_modules.py:105-108)
        mp = ref_metrics.compute_joint3d_error(ra(gt3), ra(j3d), hv).mean(axis=1)     # eval_modules.py:111-118
        pix = ref_metrics.compute_pixel_error(ref_data_utils.unormalize_kp2d(gt2, 224), ref_data_utils.unormalize_kp2d(j2d.detach(), 224), jv * hv.view(-1, 1))
        j3d_l = j3d.detach() + 0.05 * torch.randn(B, 21, 3, generator=g)
        gt3_l = gt3 + 0.05 * torch.randn(B, 21, 3, generator=g)
        mrrpe = ref_metrics.compute_mrrpe(gt3[:, 0], gt3_l[:, 0], j3d.detach()[:, 0], j3d_l[:, 0], hv)
    np.savez_compressed(
        os.path.join(HERE, "kp_loss.npz"),
        j3d=j3d.detach().numpy(), gt3=gt3.numpy(), j2d=j2d.detach().numpy(), gt2=gt2.numpy(), jv=jv.numpy(), hv=hv.numpy(),
        gate3=gate3.numpy(), gate2=gate2.numpy(), loss3=l3.detach().numpy(), loss2=l2.detach().numpy(), g3=g3.numpy(), g2=g2.numpy(),
        mpjpe=mp, pix=pix, j3d_l=j3d_l.numpy(), gt3_l=gt3_l.numpy(), mrrpe=mrrpe,
    )


def golden_process_gt():
    """The reference's own process_data_light (src/callbacks/process/process_arctic.py:4-75), run with MANO layers that wrap
    the oracle's smplx restatement on the seeded synthetic constants (smplx itself is absent): pins the glue arithmetic
    (mean-offset translation, GT camera translation, weak-perspective camera) against the reference's code."""
    import importlib.util
    from types import SimpleNamespace

    from hands_b200.synthetic import synthetic_mano_buffers
    from oracle import geometry_oracle as O

    spec = importlib.util.spec_from_file_location("ref_process_arctic", os.path.join(REF, "src/callbacks/process/process_arctic.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)

    def layer(is_rhand):
        buf = synthetic_mano_buffers(is_rhand)

        def fwd(betas, hand_pose, global_orient, transl=None):
            v, j = O.mano_forward(buf, betas, global_orient, hand_pose, transl)
            return SimpleNamespace(vertices=v, joints=j)

        return fwd

    g = torch.Generator().manual_seed(31)
    B = 9
    targets = {}
    for side in ("r", "l"):
        targets[f"mano.pose.{side}"] = 0.3 * torch.randn(B, 48, generator=g)
        targets[f"mano.beta.{side}"] = torch.randn(B, 10, generator=g)
        targets[f"mano.j3d.full.{side}"] = 0.08 * torch.randn(B, 21, 3, generator=g) + torch.tensor([0.05, -0.02, 0.7])
    f = 300 + 1200 * torch.rand(B, generator=g)
    K = torch.zeros(B, 3, 3)
    K[:, 0, 0], K[:, 1, 1], K[:, 0, 2], K[:, 1, 2], K[:, 2, 2] = f, f * 1.01, 112.0, 112.0, 1.0
    inp = {k: v.clone() for k, v in targets.items()}
    _, out, _ = mod.process_data_light({"mano_r": layer(True), "mano_l": layer(False)}, {}, targets, {"intrinsics": K}, "train", SimpleNamespace(img_res=224))
    save = {"in_" + k: v.numpy() for k, v in inp.items()}
    save["K"] = K.numpy()
    save.update({"out_" + k: v.numpy() for k, v in out.items() if k not in inp})
    np.savez_compressed(os.path.join(HERE, "process_gt.npz"), **save)


def golden_kpe():
    """KPE angles (dataset file lines 259-279, exec'd from the file like the PCL closure) and the sinusoidal encodings
    (`compute_center_pos_enc` / `compute_corner_pos_enc`, src/models/hands_light/model.py:444-460, exec'd from the file:
    the module itself needs pytorch3d)."""
    with open(os.path.join(REF, "src/datasets/hands_light_dataset.py")) as fh:
        lines = fh.readlines()
    assert "if 'center' in args.pos_enc" in lines[258] and "targets['corner.r']" in lines[278], (lines[258], lines[278])
    angle_src = textwrap.dedent("".join(lines[258:279]))
    with open(os.path.join(REF, "src/models/hands_light/model.py")) as fh:
        mlines = fh.readlines()
    assert "def compute_center_pos_enc" in mlines[443] and "return corner_pos_enc" in mlines[459], (mlines[443], mlines[459])
    enc_ns = {"torch": torch}
    exec(compile(textwrap.dedent("".join(mlines[443:460])), "<reference pos enc>", "exec"), enc_ns)
    _, bbox, K = synthetic_pcl_inputs(12, seed=5, img_res=224, smin=40, smax=200)
    K = K.clone()
    K[:, 0, 2] += torch.linspace(-8, 8, 12)   # principal point off the image centre
    K[:, 1, 1] *= 1.03
    center, corner = [], []
    for q in range(0, 12, 2):
        ns = {"np": np, "args": types.SimpleNamespace(pos_enc="center_corner"), "inputs": {}, "targets": {},
              "r_bbox": bbox[q].numpy(), "l_bbox": bbox[q + 1].numpy(), "intrx_for_enc": K[q].numpy().astype(np.float64)}
        exec(compile(angle_src, "<reference kpe angles>", "exec"), ns)
        center += [ns["inputs"]["r_center_angle"]]
        corner += [ns["inputs"]["r_corner_angle"]]
        ns["intrx_for_enc"] = K[q + 1].numpy().astype(np.float64)   # the left box with its own intrinsics row
        exec(compile(angle_src, "<reference kpe angles>", "exec"), ns)
        center += [ns["inputs"]["l_center_angle"]]
        corner += [ns["inputs"]["l_corner_angle"]]
    center, corner = torch.from_numpy(np.stack(center)), torch.from_numpy(np.stack(corner))
    L = 4
    me = types.SimpleNamespace(args=types.SimpleNamespace(n_freq_pos_enc=L))
    np.savez_compressed(os.path.join(HERE, "kpe.npz"), bbox=bbox.numpy(), K=K.numpy(), center=center.numpy(), corner=corner.numpy(), L=L,
                        center_enc=enc_ns["compute_center_pos_enc"](me, center).numpy(), corner_enc=enc_ns["compute_corner_pos_enc"](me, corner).numpy())


def golden_camera_projection():
    B = 32
    rotmat, betas, cam, K = synthetic_head_inputs(B, seed=3, small_s_frac=0.25)
    f = (K[:, 0, 0] + K[:, 1, 1]) / 2
    cam = cam.clone().requires_grad_(True)
    cam_t = ref_camera.weak_perspective_to_perspective_torch(cam, focal_length=f, img_res=224, min_s=0.1)
    g = torch.Generator().manual_seed(11)
    w3 = torch.randn(B, 3, generator=g)
    (g_cam,) = torch.autograd.grad((cam_t * w3).sum(), cam)
    wp = ref_camera.perspective_to_weak_perspective_torch(cam_t.detach(), f, 224)
    pts = (torch.randn(B, 21, 3, generator=g) * 0.05 + cam_t.detach()[:, None]).requires_grad_(True)
    j2d = ref_tf.project2d_batch(K, pts)
    j2d_n = ref_data_utils.normalize_kp2d(j2d, 224)
    w2 = torch.randn(B, 21, 2, generator=g)
    (g_pts,) = torch.autograd.grad((j2d_n * w2).sum(), pts)
    j2d_un = ref_data_utils.unormalize_kp2d(j2d_n.detach(), 224)
    np.savez_compressed(
        os.path.join(HERE, "camera_projection.npz"),
        cam=cam.detach().numpy(), K=K.numpy(), f=f.numpy(), cam_t=cam_t.detach().numpy(), w3=w3.numpy(),
        g_cam=g_cam.numpy(), wp=wp.numpy(), pts=pts.detach().numpy(), j2d=j2d.detach().numpy(),
        j2d_norm=j2d_n.detach().numpy(), w2=w2.numpy(), g_pts=g_pts.numpy(), j2d_un=j2d_un.numpy(),
    )


def _pcl_closure_source():
    """Lines 357-467 of the reference dataset file (the body of `if 'pcl' in args.pos_enc:`)."""
    path = os.path.join(REF, "src/datasets/hands_light_dataset.py")
    with open(path) as fh:
        lines = fh.readlines()
    assert "if 'pcl' in args.pos_enc" in lines[353], lines[353]
    return textwrap.dedent("".join(lines[355:467]))


def run_reference_pcl(img, r_bbox, l_bbox, intrx, img_res):
    ns = {
        "math": math, "np": np, "torch": torch, "F": F,
        "inputs": {"img": img, "r_bbox": np.asarray(r_bbox), "l_bbox": np.asarray(l_bbox)},
        "intrx": np.asarray(intrx, dtype=np.float64),
        "args": types.SimpleNamespace(img_res=img_res, pos_enc="pcl"),
    }
    exec(compile(_pcl_closure_source(), "<reference pcl closure>", "exec"), ns)
    i = ns["inputs"]
    return i["r_img"], i["l_img"], i["r_rot"], i["l_rot"], ns["r_grid_perspective"], ns["l_grid_perspective"]


def golden_pcl():
    out = {}
    # (a) small images, full outputs.  img_res=64; bboxes include: generic, square, zero-size (s -> img_res),
    #     centred on the principal point (R = I), touching the border.
    res = 64
    g = torch.Generator().manual_seed(5)
    imgs = torch.randn(4, 3, res, res, generator=g)
    Ks = np.array([[[90.0, 0, 32], [0, 90.0, 32], [0, 0, 1]],
                   [[60.0, 0, 30], [0, 75.0, 34], [0, 0, 1]],
                   [[150.0, 0, 32], [0, 150.0, 32], [0, 0, 1]],
                   [[40.0, 0, 32], [0, 40.0, 32], [0, 0, 1]]])
    r_boxes = np.array([[5, 8, 40, 30], [20, 20, 44, 44], [10, 10, 10, 10], [0, 0, 63, 63]], dtype=np.int16)
    l_boxes = np.array([[30, 2, 60, 50], [12, 12, 52, 52], [40, 3, 62, 20], [1, 30, 18, 62]], dtype=np.int16)
    crops, rots, grids_sizes = [], [], []
    for b in range(4):
        r_img, l_img, r_rot, l_rot, r_grid, l_grid = run_reference_pcl(imgs[b], r_boxes[b], l_boxes[b], Ks[b], res)
        crops += [r_img.numpy(), l_img.numpy()]
        rots += [r_rot.numpy(), l_rot.numpy()]
        grids_sizes += [r_grid.shape[0], l_grid.shape[0]]
        out[f"grid_{2*b}"] = r_grid.numpy()
        out[f"grid_{2*b+1}"] = l_grid.numpy()
    out["small_img"] = imgs.numpy()
    out["small_K"] = Ks
    out["small_bbox"] = np.stack([r_boxes, l_boxes], axis=1).reshape(8, 4)  # row 2b = right, 2b+1 = left
    out["small_crop"] = np.stack(crops)
    out["small_rot"] = np.stack(rots)
    out["small_s"] = np.array(grids_sizes)
    # (b) full-size 224 case from the synthetic generator (inputs regenerated from the seed in the test);
    #     outputs stored sub-sampled [::4, ::4] to keep the fixture small.
    img, bbox, K = synthetic_pcl_inputs(4, seed=9, img_res=224)
    sub, rots, sizes = [], [], []
    for b in range(0, 4, 2):
        r_img, l_img, r_rot, l_rot, r_grid, l_grid = run_reference_pcl(
            img[b], bbox[b].numpy().astype(np.int16), bbox[b + 1].numpy().astype(np.int16), K[b].numpy(), 224)
        sub += [r_img.numpy()[:, ::4, ::4], l_img.numpy()[:, ::4, ::4]]
        rots += [r_rot.numpy(), l_rot.numpy()]
        sizes += [r_grid.shape[0], l_grid.shape[0]]
    out["full_seed"] = np.array(9)
    out["full_crop_sub4"] = np.stack(sub)       # crop j uses image (j//2)*2, bbox j, K (j//2)*2
    out["full_rot"] = np.stack(rots)
    out["full_s"] = np.array(sizes)
    np.savez_compressed(os.path.join(HERE, "pcl.npz"), **out)


def golden_mesh_constants():
    """Wrist-seal constants and `seal_mano_mesh` of the reference (common/body_models.py:35-72).  The module itself does
    not import (trimesh/smplx/$MANO_DIR at import), so the three top-level statements are exec'd from the file's own AST --
    nothing is copied into this repository."""
    import ast

    with open(os.path.join(REF, "common", "body_models.py")) as fh:
        tree = ast.parse(fh.read())
    want = {"SEAL_FACES_R", "CIRCLE_V_ID", "seal_mano_mesh"}
    body = [n for n in tree.body
            if (isinstance(n, ast.Assign) and any(getattr(t, "id", None) in want for t in n.targets))
            or (isinstance(n, ast.FunctionDef) and n.name in want)]
    assert len(body) == 3
    ns = {"np": np, "torch": torch}
    exec(compile(ast.Module(body=body, type_ignores=[]), "ref_body_models_constants", "exec"), ns)
    g = torch.Generator().manual_seed(21)
    v3d = torch.randn(3, 778, 3, generator=g)
    faces = torch.randint(0, 778, (1538, 3), generator=g)
    out = {"SEAL_FACES_R": np.array(ns["SEAL_FACES_R"], dtype=np.int64), "CIRCLE_V_ID": np.asarray(ns["CIRCLE_V_ID"], dtype=np.int64),
           "v3d": v3d.numpy(), "faces": faces.numpy()}
    for name, is_rhand in (("r", True), ("l", False)):
        sv, sf = ns["seal_mano_mesh"](v3d, faces, is_rhand)
        out[f"sealed_v_{name}"] = sv.numpy()
        out[f"sealed_f_{name}"] = sf.numpy()
    np.savez_compressed(os.path.join(HERE, "mesh_constants.npz"), **out)


def golden_mano_head_reference_source():
    """The reference's OWN `MANOHead` (src/nets/hand_heads/mano_head.py:12-65), imported from where it lies and run on CPU.
    Its only un-importable dependency, `common.body_models` (needs smplx + the licensed pickles), is replaced by a stub
    module whose `build_mano_aa` returns the oracle's MANO restatement over the seeded synthetic buffers; every other
    module it imports (common.rot / camera / transforms / data_utils / xdict) is the reference's.  This pins the head chain
    a10-a14 (log map -> MANO call -> camera -> projection -> normalisation -> nine xdict keys) GIVEN the MANO layer; the
    MANO arithmetic inside the stub stays 'parity unpinned'."""
    import importlib.util
    from collections import namedtuple

    from hands_b200.synthetic import synthetic_mano_buffers
    from oracle import geometry_oracle as O

    Out = namedtuple("Out", ["vertices", "joints"])

    class OracleMano(torch.nn.Module):
        def __init__(self, is_rhand):
            super().__init__()
            self.buf = synthetic_mano_buffers(is_rhand)

        def forward(self, betas=None, global_orient=None, hand_pose=None, transl=None):
            return Out(*O.mano_forward(self.buf, betas, global_orient, hand_pose, transl))

    stub = types.ModuleType("common.body_models")
    stub.build_mano_aa = lambda is_rhand, create_transl=False, flat_hand=False: OracleMano(is_rhand)
    saved = sys.modules.get("common.body_models")
    sys.modules["common.body_models"] = stub
    try:
        spec = importlib.util.spec_from_file_location("ref_mano_head", os.path.join(REF, "src", "nets", "hand_heads", "mano_head.py"))
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
    finally:
        if saved is None:
            del sys.modules["common.body_models"]
        else:
            sys.modules["common.body_models"] = saved
    out = {}
    B = 12
    for name, is_rhand in (("r", True), ("l", False)):
        head = mod.MANOHead(is_rhand, 1000.0, 224.0)
        rotmat, betas, cam, K = synthetic_head_inputs(B, seed=41 + int(is_rhand), small_s_frac=0.25)
        g = torch.Generator().manual_seed(5 + int(is_rhand))
        w = {"v3d.cam": torch.randn(B, 778, 3, generator=g), "j3d.cam": torch.randn(B, 21, 3, generator=g), "j2d.norm": torch.randn(B, 21, 2, generator=g)}
        r, b, c = rotmat.clone().requires_grad_(True), betas.clone().requires_grad_(True), cam.clone().requires_grad_(True)
        o = head(r, b, c, K)
        pf = "." + name
        assert sorted(o.keys()) == sorted(k + pf for k in ["cam_t.wp", "cam_t", "joints3d", "vertices", "j3d.cam", "v3d.cam", "j2d.norm", "beta", "pose"])
        loss = sum((o[k + pf] * w[k]).sum() for k in w)
        gr, gb, gc = torch.autograd.grad(loss, (r, b, c))
        out[f"seed_{name}"] = np.array(41 + int(is_rhand))
        for k in ["cam_t", "joints3d", "vertices", "j3d.cam", "v3d.cam", "j2d.norm", "pose", "beta", "cam_t.wp"]:
            out[f"{k}_{name}"] = o[k + pf].detach().numpy()
        for k, v in w.items():
            out[f"w_{k}_{name}"] = v.numpy()
        out[f"g_rotmat_{name}"], out[f"g_betas_{name}"], out[f"g_cam_{name}"] = gr.numpy(), gb.numpy(), gc.numpy()
        # axis-angle input branch (rotmat.shape[-1] == 48, mano_head.py:31)
        aa = ref_rot.matrix_to_axis_angle(rotmat.reshape(-1, 3, 3)).reshape(-1, 48)
        o2 = head(aa, betas, cam, K)
        out[f"aa_{name}"] = aa.numpy()
        out[f"aa_j2d.norm_{name}"] = o2["j2d.norm" + pf].detach().numpy()
        out[f"aa_v3d.cam_{name}"] = o2["v3d.cam" + pf].detach().numpy()
    np.savez_compressed(os.path.join(HERE, "mano_head_ref.npz"), **{k: (v.astype(np.float32) if v.dtype == np.float64 else v) for k, v in out.items()})


def golden_loss_light():
    """The reference's own `compute_loss_light` (src/callbacks/loss/loss_arctic_sf.py:20-171).  The module imports pytorch3d
    (absent) for one function, `axis_angle_to_matrix`; the function definition is exec'd from the file's AST with that name
    bound to the reference's own port of it (common/rot.py: quaternion_to_matrix(axis_angle_to_quaternion(.)), the same
    pytorch3d code) and every other name bound to src/utils/loss_modules.py -- nothing is copied."""
    import ast

    import src.utils.loss_modules as ref_lm

    with open(os.path.join(REF, "src", "callbacks", "loss", "loss_arctic_sf.py")) as fh:
        tree = ast.parse(fh.read())
    fn = [n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name == "compute_loss_light"]
    assert len(fn) == 1
    ns = {"torch": torch, "nn": torch.nn, "l1_loss": torch.nn.L1Loss(reduction="none"), "mse_loss": torch.nn.MSELoss(reduction="none"),
          "axis_angle_to_matrix": lambda aa: ref_rot.quaternion_to_matrix(ref_rot.axis_angle_to_quaternion(aa))}
    for name in ("compute_contact_devi_loss", "hand_kp3d_loss", "joints_loss", "mano_loss", "object_kp3d_loss", "vector_loss", "grasp_loss"):
        ns[name] = getattr(ref_lm, name)
    exec(compile(ast.Module(body=fn, type_ignores=[]), "ref_compute_loss_light", "exec"), ns)
    g = torch.Generator().manual_seed(31)
    B = 24
    rn = lambda *sh: torch.randn(*sh, generator=g)  # noqa: E731
    mask = lambda p: (torch.rand(B, generator=g) > p).float()  # noqa: E731
    pred, gt = {}, {}
    for sd in ("r", "l"):
        pred[f"mano.beta.{sd}"] = rn(B, 10)
        pred[f"mano.pose.{sd}"] = random_rotmats(B * 16, g).reshape(B, 16, 3, 3)
        pred[f"mano.j3d.cam.{sd}"] = 0.1 * rn(B, 21, 3) + torch.tensor([0.0, 0.0, 0.6])
        pred[f"mano.j2d.norm.{sd}"] = 0.5 * rn(B, 21, 2)
        pred[f"mano.cam_t.wp.{sd}"] = rn(B, 3)
        pred[f"mano.cam_t.wp.init.{sd}"] = rn(B, 3)
        gt[f"mano.pose.{sd}"] = 0.4 * rn(B, 48)
        gt[f"mano.pose.{sd}"][0, :3] = 0.0          # small-angle branch of axis_angle_to_quaternion
        gt[f"mano.beta.{sd}"] = rn(B, 10)
        gt[f"mano.j3d.cam.{sd}"] = 0.1 * rn(B, 21, 3) + torch.tensor([0.0, 0.0, 0.6])
        gt[f"mano.j2d.norm.{sd}"] = 0.5 * rn(B, 21, 2)
        gt[f"mano.cam_t.wp.{sd}"] = rn(B, 3)
        gt[f"joints_valid_{sd}"] = (torch.rand(B, 21, generator=g) > 0.2).float()
    gt["is_valid"], gt["right_valid"], gt["left_valid"] = mask(0.1), mask(0.2), mask(0.2)
    meta = {k: mask(0.3) for k in ("is_cam_loss", "is_j2d_loss", "is_j3d_loss", "is_pose_loss", "is_beta_loss")}
    leaves = {k: v.clone().requires_grad_(True) for k, v in pred.items()}
    class Args(dict):     # easydict stand-in: .get() and attribute access (args.regress_center_corner, :196)
        __getattr__ = dict.get

    out = ns["compute_loss_light"](leaves, gt, meta, Args(regress_center_corner=False))
    total = sum(l * w for l, w in out.values())
    grads = torch.autograd.grad(total, list(leaves.values()), allow_unused=True)
    sav = {"B": np.array(B)}
    for k, v in pred.items():
        sav["pred:" + k] = v.numpy()
    for k, v in gt.items():
        sav["gt:" + k] = v.numpy()
    for k, v in meta.items():
        sav["meta:" + k] = v.numpy()
    for k, (l, w) in out.items():
        sav["loss:" + k] = l.detach().numpy()
        sav["weight:" + k] = np.array(w)
    for k, gr in zip(leaves, grads):
        sav["grad:" + k] = (gr if gr is not None else torch.zeros_like(pred[k])).numpy()
    sav["gt_rotmat_r"] = ns["axis_angle_to_matrix"](gt["mano.pose.r"].reshape(-1, 3)).numpy()
    np.savez_compressed(os.path.join(HERE, "loss_light.npz"), **sav)


def golden_decimator():
    """`MANODecimator.downsample` of the reference (common/body_models.py:11-32), class definition exec'd from the file's AST,
    fed a synthetic 195x778 decimation matrix through a temporary $DATA_DIR (the real .npy is ARCTIC data, unavailable)."""
    import ast
    import tempfile

    with open(os.path.join(REF, "common", "body_models.py")) as fh:
        tree = ast.parse(fh.read())
    cls = [n for n in tree.body if isinstance(n, ast.ClassDef) and n.name == "MANODecimator"]
    assert len(cls) == 1
    ns = {"np": np, "torch": torch, "os": os}
    exec(compile(ast.Module(body=cls, type_ignores=[]), "ref_mano_decimator", "exec"), ns)
    g = torch.Generator().manual_seed(51)
    D = {}
    for flag in ("right", "left"):
        d = torch.zeros(195, 778)
        idx = torch.randint(0, 778, (195, 3), generator=g)
        wts = torch.rand(195, 3, generator=g)
        d.scatter_(1, idx, wts / wts.sum(1, keepdim=True))      # decimation rows: a few positive weights summing to one
        D[f"D_{flag}"] = d.numpy()
    verts = 0.05 * torch.randn(5, 778, 3, generator=g)
    old = os.environ.get("DATA_DIR")
    with tempfile.TemporaryDirectory() as tmp:
        path = os.path.join(tmp, "arctic", "data", "arctic_data", "data", "meta")
        os.makedirs(path)
        np.save(os.path.join(path, "mano_decimator_195.npy"), D, allow_pickle=True)
        os.environ["DATA_DIR"] = tmp
        try:
            dec = ns["MANODecimator"]()
            out_r, out_l = dec.downsample(verts, True), dec.downsample(verts, False)
        finally:
            if old is None:
                del os.environ["DATA_DIR"]
            else:
                os.environ["DATA_DIR"] = old
    np.savez_compressed(os.path.join(HERE, "decimator.npz"), D_right=D["D_right"], D_left=D["D_left"], verts=verts.numpy(), sub_r=out_r.numpy(), sub_l=out_l.numpy())


if __name__ == "__main__":
    torch.set_num_threads(1)
    golden_logmap()
    golden_camera_projection()
    golden_pcl()
    golden_rot6d()
    golden_kp_loss()
    golden_process_gt()
    golden_kpe()
    golden_mesh_constants()
    golden_mano_head_reference_source()
    golden_loss_light()
    golden_decimator()
    for f in sorted(os.listdir(HERE)):
        if f.endswith(".npz"):
            print(f, os.path.getsize(os.path.join(HERE, f)), "bytes")
