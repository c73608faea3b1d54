"""Seeded MANO-shaped constant buffers and synthetic inputs.

The licensed ``MANO_{RIGHT,LEFT}.pkl`` files read by the reference
(``common/body_models.py:90-99``) are not available offline, so benchmarks, tests and
``smoke()`` use random buffers with the real model's shapes and scales.  The recipe is
SURVEY.md §8(d): seed 0 for the right hand, 1 for the left.

Everything here is plain CPU torch; nothing in this file touches the GPU.
"""
import math

import torch

NUM_VERTS = 778
NUM_JOINTS = 16
NUM_BETAS = 10
NUM_POSE_FEAT = 135
NUM_FACES = 1538
# kintree of the MANO hand: five 3-joint chains off the wrist (index, middle, pinky, ring, thumb)
PARENTS = [-1, 0, 1, 2, 0, 4, 5, 0, 7, 8, 0, 10, 11, 0, 13, 14]
# smplx vertex_ids['mano']: thumb, index, middle, ring, pinky finger tips
TIP_IDS = [744, 320, 443, 554, 671]


def synthetic_mano_buffers(is_rhand=True, flat_hand=False, sparse_weights=False, seed=None):
    """Return a dict of fp32/int64 CPU tensors shaped like the smplx MANO buffers."""
    if seed is None:
        seed = 0 if is_rhand else 1
    g = torch.Generator().manual_seed(seed)

    def randn(*shape, std=1.0):
        return torch.randn(*shape, generator=g, dtype=torch.float64) * std

    v_template = randn(NUM_VERTS, 3, std=0.03)
    shapedirs = randn(NUM_VERTS, 3, NUM_BETAS, std=0.005)
    posedirs_raw = randn(NUM_VERTS, 3, NUM_POSE_FEAT, std=0.002)
    posedirs = posedirs_raw.reshape(-1, NUM_POSE_FEAT).T.contiguous()  # (135, 2334), column 3v+k
    J_regressor = torch.rand(NUM_JOINTS, NUM_VERTS, generator=g, dtype=torch.float64)
    J_regressor = J_regressor / J_regressor.sum(dim=1, keepdim=True)
    lbs = torch.rand(NUM_VERTS, NUM_JOINTS, generator=g, dtype=torch.float64) ** 8
    if sparse_weights:
        # MANO-like: keep the 4 largest weights per vertex
        top = lbs.topk(4, dim=1)
        lbs = torch.zeros_like(lbs).scatter_(1, top.indices, top.values)
    lbs = lbs / lbs.sum(dim=1, keepdim=True)
    hands_mean = randn(45, std=0.1)
    if flat_hand:
        hands_mean = torch.zeros_like(hands_mean)
    pose_mean = torch.cat([torch.zeros(3, dtype=torch.float64), hands_mean])
    faces = torch.randint(0, NUM_VERTS, (NUM_FACES, 3), generator=g, dtype=torch.int64)
    return {
        "v_template": v_template.float(),
        "shapedirs": shapedirs.float(),
        "posedirs": posedirs.float(),
        "J_regressor": J_regressor.float(),
        "lbs_weights": lbs.float(),
        "parents": torch.tensor(PARENTS, dtype=torch.int64),
        "pose_mean": pose_mean.float(),
        "faces": faces,
        "tip_ids": torch.tensor(TIP_IDS, dtype=torch.int64),
    }


def random_rotmats(n, generator, edge=None):
    """(n,3,3) rotation matrices: Gram-Schmidt of N(0,1) 6-vectors."""
    x = torch.randn(n, 6, generator=generator, dtype=torch.float64)
    a, b = x[:, :3], x[:, 3:]
    e1 = a / a.norm(dim=1, keepdim=True)
    b = b - (e1 * b).sum(1, keepdim=True) * e1
    e2 = b / b.norm(dim=1, keepdim=True)
    e3 = torch.linalg.cross(e1, e2)
    R = torch.stack([e1, e2, e3], dim=-1)
    if edge == "identity":
        R = torch.eye(3, dtype=torch.float64).expand(n, 3, 3).clone()
    elif edge == "near_pi":
        axis = e1
        ang = math.pi - 1e-3
        K = torch.zeros(n, 3, 3, dtype=torch.float64)
        K[:, 0, 1], K[:, 0, 2] = -axis[:, 2], axis[:, 1]
        K[:, 1, 0], K[:, 1, 2] = axis[:, 2], -axis[:, 0]
        K[:, 2, 0], K[:, 2, 1] = -axis[:, 1], axis[:, 0]
        R = torch.eye(3, dtype=torch.float64) + math.sin(ang) * K + (1 - math.cos(ang)) * K @ K
    return R.float()


def synthetic_head_inputs(B, seed=0, img_res=224.0, edge=None, small_s_frac=0.0):
    """Inputs of ``MANOHead.forward``: rotmat (B,16,3,3), betas (B,10), cam (B,3), K (B,3,3)."""
    g = torch.Generator().manual_seed(1000 + seed)
    rotmat = random_rotmats(B * NUM_JOINTS, g, edge=edge).reshape(B, NUM_JOINTS, 3, 3)
    betas = torch.randn(B, NUM_BETAS, generator=g)
    s = torch.rand(B, generator=g) + 0.5
    n_small = int(B * small_s_frac)
    if n_small:
        s[:n_small] = torch.rand(n_small, generator=g) * 0.09
    txy = torch.randn(B, 2, generator=g) * 0.2
    cam = torch.cat([s[:, None], txy], dim=1)
    f = torch.rand(B, generator=g) * 1200.0 + 300.0
    if B > 1:
        f[-1] = 1000.0
    K = torch.zeros(B, 3, 3)
    K[:, 0, 0] = f
    K[:, 1, 1] = f
    K[:, 0, 2] = img_res / 2
    K[:, 1, 2] = img_res / 2
    K[:, 2, 2] = 1.0
    return rotmat.contiguous(), betas.contiguous(), cam.contiguous(), K.contiguous()


def synthetic_pcl_inputs(B, seed=0, img_res=224, smin=56, smax=168, smooth=False):
    """img (B,3,R,R) fp32, bbox (B,4) int32 [x0,y0,x1,y1] fully inside the image, K (B,3,3)."""
    g = torch.Generator().manual_seed(2000 + seed)
    img = torch.randn(B, 3, img_res, img_res, generator=g)
    if smooth:
        k = torch.ones(1, 1, 9, 9) / 81.0
        img = torch.nn.functional.conv2d(
            img.reshape(B * 3, 1, img_res, img_res), k, padding=4
        ).reshape(B, 3, img_res, img_res)
    side = torch.randint(smin, smax + 1, (B,), generator=g)
    other = (side.double() * (0.6 + 0.4 * torch.rand(B, generator=g, dtype=torch.float64))).long().clamp(min=1)
    wide = torch.rand(B, generator=g) < 0.5
    w = torch.where(wide, side, other)
    h = torch.where(wide, other, side)
    x0 = (torch.rand(B, generator=g) * (img_res - 1 - w).float()).long()
    y0 = (torch.rand(B, generator=g) * (img_res - 1 - h).float()).long()
    bbox = torch.stack([x0, y0, x0 + w, y0 + h], dim=1).int()
    f = torch.rand(B, generator=g) * 1200.0 + 300.0
    K = torch.zeros(B, 3, 3)
    K[:, 0, 0] = f
    K[:, 1, 1] = f
    K[:, 0, 2] = img_res / 2
    K[:, 1, 2] = img_res / 2
    K[:, 2, 2] = 1.0
    return img.contiguous(), bbox.contiguous(), K.contiguous()


def synthetic_tube_mesh(seed=0):
    """A hand-sized closed-ended tube with MANO's own counts and topology class -- 778 vertices, 1538 faces, one open
    boundary ring of 16 vertices (the wrist; MANO: common/body_models.py:35-63) -- for the silhouette renderer, whose
    cost and coverage depend on the triangles being local (the random `faces` of `synthetic_mano_buffers` are not).
    48 rings of 16 vertices + a cap of 8 + 2 vertices.  Returns verts (778,3) fp32 in metres, centred, long axis = y,
    and faces (1538,3) int64."""
    g = torch.Generator().manual_seed(3000 + seed)
    rings, seg = 48, 16
    k = torch.arange(rings, dtype=torch.float64)
    u = k / (rings - 1)
    y = -0.085 + 0.15 * u
    rad = 0.036 * (1.0 - 0.45 * u) * (1.0 + 0.12 * torch.sin(3.0 * math.pi * u))
    th = torch.arange(seg, dtype=torch.float64) * (2.0 * math.pi / seg)
    bend = 0.02 * u * u
    ring = torch.stack([rad[:, None] * torch.cos(th)[None], y[:, None].expand(rings, seg), 0.45 * rad[:, None] * torch.sin(th)[None] + bend[:, None]], -1)
    th8 = th[0::2] + math.pi / seg
    mid = torch.stack([0.55 * rad[-1] * torch.cos(th8), torch.full((8,), float(y[-1]) + 0.006, dtype=torch.float64), 0.45 * 0.55 * rad[-1] * torch.sin(th8) + bend[-1]], -1)
    ctr = torch.tensor([[0.004, float(y[-1]) + 0.009, float(bend[-1])], [-0.004, float(y[-1]) + 0.009, float(bend[-1])]], dtype=torch.float64)
    verts = torch.cat([ring.reshape(-1, 3), mid, ctr], 0)
    verts = verts + 2e-4 * torch.randn(verts.shape, generator=g, dtype=torch.float64)
    faces = []
    for r in range(rings - 1):
        for s in range(seg):
            a, b = r * seg + s, r * seg + (s + 1) % seg
            faces += [[a, b, a + seg], [b, b + seg, a + seg]]
    o = lambda i: (rings - 1) * seg + i % seg   # noqa: E731  last ring
    m = lambda i: rings * seg + i % 8           # noqa: E731
    c0, c1 = rings * seg + 8, rings * seg + 9
    for i in range(8):
        faces += [[o(2 * i), o(2 * i + 1), m(i)], [o(2 * i + 1), o(2 * i + 2), m(i)], [o(2 * i + 2), m(i + 1), m(i)]]
    # th8[0..1] and th8[6..7] face +x (c0), th8[2..5] face -x (c1)
    faces += [[m(i), m(i + 1), c0] for i in (6, 7, 0)] + [[m(i), m(i + 1), c1] for i in (2, 3, 4)]
    faces += [[m(1), m(2), c1], [m(1), c1, c0], [m(5), m(6), c0], [m(5), c0, c1]]
    verts = verts - verts.mean(0, keepdim=True)
    faces = torch.tensor(faces, dtype=torch.int64)
    assert verts.shape == (NUM_VERTS, 3) and faces.shape == (NUM_FACES, 3)
    return verts.float().contiguous(), faces.contiguous()


def synthetic_silhouette_inputs(B, seed=0, img_res=224):
    """verts_cam (B,778,3): the tube mesh under a random rotation, 0.45-0.9 m in front of the camera, projected centre
    within the middle of the image; faces (1538,3); K (B,3,3) with f ~ U(500, 900)."""
    g = torch.Generator().manual_seed(4000 + seed)
    verts, faces = synthetic_tube_mesh(0)
    R = random_rotmats(B, g).reshape(B, 3, 3)
    f = torch.rand(B, generator=g) * 400.0 + 500.0
    z = torch.rand(B, generator=g) * 0.45 + 0.45
    cxy = (torch.rand(B, 2, generator=g) - 0.5) * 0.3 * img_res   # pixel offset of the hand centre
    t = torch.stack([cxy[:, 0] * z / f, cxy[:, 1] * z / f, z], -1)
    vc = torch.einsum("bij,vj->bvi", R, verts) + t[:, None]
    K = torch.zeros(B, 3, 3)
    K[:, 0, 0] = f
    K[:, 1, 1] = f
    K[:, 0, 2] = img_res / 2
    K[:, 1, 2] = img_res / 2
    K[:, 2, 2] = 1.0
    return vc.float().contiguous(), faces, K.contiguous()
