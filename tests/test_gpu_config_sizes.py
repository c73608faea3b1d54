"""GPU parity at the BASELINE.json config sizes, against the CPU oracle (not against the CUDA path itself), plus the
reference-side binding tests of seams 1-3.

  C2  MANOHead fwd+bwd, B = 1024, both hand sides: six outputs + three gradients
  C3  PerspectiveCropLayer, 1024 crops of 3x224x224: crops and g_img
  C4  one sample batch in the C4 shape: 2 crops per source image, R_virt2orig chained into both heads as pre_rot

Tolerances: BASELINE.json north_star (vertices/joints 1e-5 relative, key-points 1e-3 px, gradients 1e-4 relative,
indices bit-exact).  Crops: 1e-5 relative to the crop's value range in the default (fast-division) forward, 1e-6 absolute
and > 99.9 % bit-equal pixels in the exact mode (HB_PCL_EXACT=1 / hb_pcl_set_exact(1)).
"""
import os

import numpy as np
import pytest
import torch

from _tol import tol_check
from hands_b200.synthetic import synthetic_head_inputs, synthetic_mano_buffers, synthetic_pcl_inputs
from oracle import geometry_oracle as O

pytestmark = pytest.mark.gpu
IMG_RES = 224.0
KEYS6 = ("vertices", "joints3d", "v3d.cam", "j3d.cam", "j2d.norm", "cam_t")


def rel(got, ref):
    ref = ref.double()
    return float((got.double().cpu() - ref).abs().max() / ref.abs().max().clamp_min(1e-30))


@pytest.fixture(scope="module")
def dev():
    assert torch.cuda.is_available(), "gpu tests need a CUDA device"
    return torch.device("cuda:0")


@pytest.fixture(scope="module")
def heads(dev):
    from hands_b200.src.nets.hand_heads.mano_head import MANOHead

    return {s: MANOHead(s, 1000.0, IMG_RES, synthetic=True).to(dev) for s in (True, False)}


def _oracle(is_rhand, rotmat, betas, cam, K, dtype):
    return O.mano_head_forward(synthetic_mano_buffers(is_rhand), rotmat.to(dtype), betas.to(dtype), cam.to(dtype), K.to(dtype), IMG_RES, 0.1)


def _grads(fn, rotmat, betas, cam, K, w):
    r, b, c = rotmat.clone().requires_grad_(True), betas.clone().requires_grad_(True), cam.clone().requires_grad_(True)
    out = fn(r, b, c, K)
    loss = sum((out[k] * w[k].to(out[k].device, out[k].dtype)).sum() for k in w)
    return out, torch.autograd.grad(loss, (r, b, c))


@pytest.mark.parametrize("is_rhand", [True, False])
def test_C2_head_fwd_bwd_B1024_against_oracle(heads, dev, is_rhand):
    B = 1024
    rotmat, betas, cam, K = synthetic_head_inputs(B, seed=2024 + int(is_rhand), small_s_frac=0.05)
    g = torch.Generator().manual_seed(3)
    w = {"v3d.cam": torch.randn(B, 778, 3, generator=g), "j3d.cam": torch.randn(B, 21, 3, generator=g), "j2d.norm": torch.randn(B, 21, 2, generator=g)}
    pf = ".r" if is_rhand else ".l"

    def ours(r, b, c, k):
        o = heads[is_rhand](r, b, c, k)
        return {key: o[key + pf] for key in KEYS6}

    out, got = _grads(ours, rotmat.to(dev), betas.to(dev), cam.to(dev), K.to(dev), w)
    o64, g64 = _grads(lambda r, b, c, k: _oracle(is_rhand, r, b, c, k, torch.float64), rotmat.double(), betas.double(), cam.double(), K.double(), w)
    o32, g32 = _grads(lambda r, b, c, k: _oracle(is_rhand, r, b, c, k, torch.float32), rotmat, betas, cam, K, w)
    for key in ("vertices", "joints3d", "v3d.cam", "j3d.cam", "cam_t"):
        tol_check(f"C2[{pf}].{key}", rel(out[key], o64[key].detach()), 1e-5, rel(o32[key].detach(), o64[key].detach()))
    px = (out["j2d.norm"].double().cpu() - o64["j2d.norm"].detach()).abs().max() * IMG_RES / 2
    own_px = (o32["j2d.norm"].detach().double() - o64["j2d.norm"].detach()).abs().max() * IMG_RES / 2
    tol_check(f"C2[{pf}].j2d_px", px, 1e-3, own_px)
    for name, a, r64, r32 in zip(("rotmat", "betas", "cam"), got, g64, g32):
        tol_check(f"C2[{pf}].g_{name}", rel(a, r64), 1e-4, rel(r32, r64))
    assert torch.equal(out["joints3d"][:, 16:], out["vertices"][:, list(O.TIP_IDS)])   # index work: bit-exact


def _oracle_pcl(img, bbox, K, res, cpi, w):
    nt = torch.get_num_threads()
    torch.set_num_threads(1)   # torch's bilinear kernels differ by 1 ulp between 1 and N threads; the goldens are 1-thread
    crops, rots, grads = [], [], []
    try:
        for b in range(img.shape[0]):   # one autograd leaf per source image (a slice of one big leaf costs a batch-sized zero gradient per crop)
            xr = img[b : b + 1].clone().requires_grad_(True)
            sl = slice(b * cpi, (b + 1) * cpi)
            c, r = O.perspective_crop(xr.expand(cpi, -1, -1, -1), bbox[sl], K[sl], res)
            (gx,) = torch.autograd.grad((c * w[sl]).sum(), xr)
            crops.append(c.detach()); rots.append(r); grads.append(gx)
    finally:
        torch.set_num_threads(nt)
    return torch.cat(crops), torch.cat(rots), torch.cat(grads)


_C3 = {}


def _c3_case():
    """Inputs and oracle results of the C3 case, computed once (the single-threaded oracle takes a few seconds for 1024 crops)."""
    if not _C3:
        n, res = 1024, 224
        img, bbox, K = synthetic_pcl_inputs(n, seed=33, img_res=res, smin=res // 4, smax=3 * res // 4)
        w = torch.randn(n, 3, res, res, generator=torch.Generator().manual_seed(6))
        _C3["v"] = (img, bbox, K, w) + _oracle_pcl(img, bbox, K, res, 1, w)
    return _C3["v"]


@pytest.mark.parametrize("exact", [0, 1])
def test_C3_pcl_1024_crops_against_oracle(dev, exact):
    from hands_b200 import _lib
    from hands_b200.pcl import perspective_crop

    res = 224
    img, bbox, K, w, ref_crop, ref_rot, ref_g = _c3_case()
    prev = _lib.load().hb_pcl_set_exact(exact)
    try:
        x = img.to(dev).requires_grad_(True)
        crop, rot = perspective_crop(x, bbox.to(dev), K.to(dev), img_res=res)
        (g_img,) = torch.autograd.grad((crop * w.to(dev)).sum(), x)
        crop, rot, g_img = crop.detach().cpu(), rot.cpu(), g_img.cpu()
    finally:
        _lib.load().hb_pcl_set_exact(prev)
    assert (rot - ref_rot).abs().max() <= 1.2e-7
    err = (crop - ref_crop).abs()
    if exact:
        # P differs from numpy's LAPACK inverse in the last fp32 bit for some crops (DESIGN.md "PCL exactness"): 1e-5 px in
        # the sample position, up to 5e-5 on white noise
        assert err.max() <= 5e-5, float(err.max())
        assert float((err <= 1e-6).float().mean()) > 0.99
    tol_check(f"C3[exact={exact}].crop_rel", float(err.max() / ref_crop.abs().max()), 1e-5)
    tol_check(f"C3[exact={exact}].g_img", rel(g_img, ref_g), 1e-4)


def test_C4_sample_batch_against_oracle(heads, dev):
    """One C4-shaped batch: S source images, two crops each (right, left), R_virt2orig of each crop chained into its hand's
    MANOHead as `pre_rot` (hands_light/model.py:330-334), gradients through both."""
    from hands_b200.pcl import perspective_crop

    S, res = 48, 224
    n = 2 * S
    img, bbox, Kc = synthetic_pcl_inputs(n, seed=44, img_res=res, smin=res // 4, smax=3 * res // 4)
    img = img[:S].contiguous()
    w = torch.randn(n, 3, res, res, generator=torch.Generator().manual_seed(8))
    x = img.to(dev).requires_grad_(True)
    crop, rot = perspective_crop(x, bbox.to(dev), Kc.to(dev), img_res=res, crops_per_img=2)
    (g_img,) = torch.autograd.grad((crop * w.to(dev)).sum(), x)
    ref_crop, ref_rot, ref_g = _oracle_pcl(img, bbox, Kc, res, 2, w)
    tol_check("C4.crop_rel", float((crop.detach().cpu() - ref_crop).abs().max() / ref_crop.abs().max()), 1e-5)
    tol_check("C4.g_img", rel(g_img, ref_g), 1e-4)
    rot_s = rot.view(S, 2, 3, 3)
    for side, is_rhand in enumerate((True, False)):
        pf = ".r" if is_rhand else ".l"
        rotmat, betas, cam, K = synthetic_head_inputs(S, seed=70 + side, small_s_frac=0.1)
        g = torch.Generator().manual_seed(side)
        wk = {"v3d.cam": torch.randn(S, 778, 3, generator=g), "j3d.cam": torch.randn(S, 21, 3, generator=g), "j2d.norm": torch.randn(S, 21, 2, generator=g)}
        Rv = rot_s[:, side].contiguous()

        def ours(r, b, c, k):
            o = heads[is_rhand](r, b, c, k, pre_rot=Rv)
            return {key: o[key + pf] for key in KEYS6 + ("pose",)}

        def oracle(dtype):
            Rc = ref_rot.view(S, 2, 3, 3)[:, side].to(dtype)

            def f(r, b, c, k):
                o = _oracle(is_rhand, O.pcl_fix_global_orient(Rc, r), b, c, k, dtype)
                return o
            return f

        out, got = _grads(ours, rotmat.to(dev), betas.to(dev), cam.to(dev), K.to(dev), wk)
        o64, g64 = _grads(oracle(torch.float64), rotmat.double(), betas.double(), cam.double(), K.double(), wk)
        o32, g32 = _grads(oracle(torch.float32), rotmat, betas, cam, K, wk)
        for key in ("vertices", "v3d.cam", "j3d.cam"):
            tol_check(f"C4[{pf}].{key}", rel(out[key], o64[key].detach()), 1e-5, rel(o32[key].detach(), o64[key].detach()))
        px = (out["j2d.norm"].double().cpu() - o64["j2d.norm"].detach()).abs().max() * IMG_RES / 2
        tol_check(f"C4[{pf}].j2d_px", px, 1e-3, (o32["j2d.norm"].detach().double() - o64["j2d.norm"].detach()).abs().max() * IMG_RES / 2)
        for name, a, r64, r32 in zip(("rotmat", "betas", "cam"), got, g64, g32):
            tol_check(f"C4[{pf}].g_{name}", rel(a, r64), 1e-4, rel(r32, r64))
        # the returned pose is the ROTATED one, as in the reference (the in-place bmm precedes the head)
        tol_check(f"C4[{pf}].pose", rel(out["pose"], o64["pose"].detach()), 1e-6)


# ---- reference-side binding (seams 1-3) --------------------------------------------------------------------------------
def _reference_call_sequence(mano_layer, rotmat, shape, cam, K, img_res):
    """The call sequence of the reference's MANOHead.forward (src/nets/hand_heads/mano_head.py:30-51), issued against the
    DROP-IN modules under the reference's own import names (common.rot / camera / transforms / data_utils / body_models)."""
    from hands_b200.common import camera, data_utils, rot, transforms as tf

    aa = rotmat
    if rotmat.shape[-1] != 48:
        aa = rot.matrix_to_axis_angle(rotmat.reshape(-1, 3, 3)).reshape(-1, 48)
    mo = mano_layer(betas=shape, hand_pose=aa[:, 3:], global_orient=aa[:, :3])
    f = (K[:, 0, 0] + K[:, 1, 1]) / 2.0
    cam_t = camera.weak_perspective_to_perspective_torch(cam, focal_length=f, img_res=img_res, min_s=0.1)
    j3d = mo.joints + cam_t[:, None, :]
    v3d = mo.vertices + cam_t[:, None, :]
    j2d = data_utils.normalize_kp2d(tf.project2d_batch(K, j3d), img_res)
    return {"cam_t": cam_t, "joints3d": mo.joints, "vertices": mo.vertices, "j3d.cam": j3d, "v3d.cam": v3d, "j2d.norm": j2d}


@pytest.mark.parametrize("is_rhand", [True, False])
def test_reference_mano_head_source_binding(heads, dev, golden_dir, is_rhand):
    """Seam 1 + 3: (i) the fused MANOHead against the output of the reference's OWN MANOHead.forward source
    (tests/golden/mano_head_ref.npz: mano_head.py:21-65 imported from the reference and run over the oracle's MANO layer);
    (ii) the same call sequence issued op by op through the drop-in modules (build_mano_aa layer + common.rot / camera /
    transforms / data_utils) against the fused head: <= 1e-6, gradients <= 1e-5."""
    g = np.load(os.path.join(golden_dir, "mano_head_ref.npz"))
    nm = "r" if is_rhand else "l"
    pf = "." + nm
    B = 12
    rotmat, betas, cam, K = synthetic_head_inputs(B, seed=int(g[f"seed_{nm}"]), small_s_frac=0.25)
    w = {k: torch.from_numpy(g[f"w_{k}_{nm}"]) for k in ("v3d.cam", "j3d.cam", "j2d.norm")}
    head = heads[is_rhand]

    def fused(r, b, c, k):
        o = head(r, b, c, k)
        return {key: o[key + pf] for key in KEYS6 + ("pose", "beta", "cam_t.wp")}

    out, got = _grads(fused, rotmat.to(dev), betas.to(dev), cam.to(dev), K.to(dev), w)
    for key in ("vertices", "joints3d", "v3d.cam", "j3d.cam", "cam_t"):
        tol_check(f"refsrc[{nm}].{key}", rel(out[key], torch.from_numpy(g[f"{key}_{nm}"])), 1e-5)
    tol_check(f"refsrc[{nm}].j2d_px", (out["j2d.norm"].cpu() - torch.from_numpy(g[f"j2d.norm_{nm}"])).abs().max() * IMG_RES / 2, 1e-3)
    for key in ("pose", "beta", "cam_t.wp"):
        assert torch.equal(out[key].cpu(), torch.from_numpy(g[f"{key}_{nm}"])), key
    for name, a in zip(("rotmat", "betas", "cam"), got):
        ref = torch.from_numpy(g[f"g_{name}_{nm}"])
        tol_check(f"refsrc[{nm}].g_{name}", rel(a, ref), 1e-4)
    # axis-angle input branch
    aa = torch.from_numpy(g[f"aa_{nm}"]).to(dev)
    o2 = head(aa, betas.to(dev), cam.to(dev), K.to(dev))
    tol_check(f"refsrc[{nm}].aa.v3d", rel(o2["v3d.cam" + pf], torch.from_numpy(g[f"aa_v3d.cam_{nm}"])), 1e-5)
    tol_check(f"refsrc[{nm}].aa.j2d_px", (o2["j2d.norm" + pf].cpu() - torch.from_numpy(g[f"aa_j2d.norm_{nm}"])).abs().max() * IMG_RES / 2, 1e-3)
    # (ii) op-by-op through the drop-ins
    seq_out, seq_g = _grads(lambda r, b, c, k: _reference_call_sequence(head.mano, r, b, c, k, IMG_RES), rotmat.to(dev), betas.to(dev), cam.to(dev), K.to(dev), w)
    for key in KEYS6:
        tol_check(f"dropin_seq[{nm}].{key}", rel(seq_out[key], out[key].detach().cpu()), 1e-6)
    for name, a, b in zip(("rotmat", "betas", "cam"), seq_g, got):
        tol_check(f"dropin_seq[{nm}].g_{name}", rel(a, b.cpu()), 1e-5)


def test_projection_small_depth(dev):
    """`to_xy_batch` divides by z with no eps (common/transforms.py:76) -- tiny and negative depths
    must give the same huge / sign-flipped pixels as the reference formula, and z == 0 the same inf/nan pattern."""
    from hands_b200.common import transforms as tf

    B, N = 4, 7
    g = torch.Generator().manual_seed(12)
    pts = torch.randn(B, N, 3, generator=g)
    pts[0, :, 2] = 1e-6
    pts[1, :, 2] = -1e-3
    pts[2, :, 2] = 1e-20
    pts[3, 0, 2] = 0.0
    _, _, _, K = synthetic_head_inputs(B, seed=1)
    got = tf.project2d_batch(K.to(dev), pts.to(dev)).cpu()
    ref = O.project2d_batch(K.double(), pts.double())
    fin = torch.isfinite(ref) & (ref.abs() < 1e30)
    assert torch.isfinite(got[fin]).all()
    tol_check("project_small_z", float(((got.double() - ref)[fin].abs() / ref[fin].abs().clamp_min(1.0)).max()), 1e-5)
    assert not torch.isfinite(got[3, 0]).all()   # z == 0: inf or nan, never a silently clamped number
