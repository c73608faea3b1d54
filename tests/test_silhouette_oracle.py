"""CPU checks of oracle/silhouette_oracle.py (the restatement of pytorch3d's soft silhouette as the reference configures it,
src/models/hands_light/renderer.py:124-209).  pytorch3d is absent: PARITY UNPINNED.  What can be pinned offline is pinned
here: the image convention the reference ends up with after `flip_transpose_canvas`, closed-form values of the blend on
hand-made triangles, the faces_per_pixel selection, and the analytic backward against central differences of the forward."""
import math

import numpy as np

from oracle import silhouette_oracle as so

S = 32
K1 = np.array([[[40.0, 0, 16], [0, 40.0, 16], [0, 0, 1]]])


def _tri(px):
    """camera-space triangle (z = 1) from three pixel-coordinate corners under K1"""
    px = np.asarray(px, np.float64)
    return np.concatenate([(px - 16.0) / 40.0, np.ones((3, 1))], 1)[None]


def test_image_convention_is_the_K_projection():
    # a triangle over pixel columns 2..9, rows 20..27 (u = column + 0.5, v = row + 0.5): after the reference's canvas flip
    # the mask lives at those rows/columns, not mirrored
    m = so.soft_silhouette(_tri([[2, 20], [10, 20], [2, 28]]), np.array([[0, 1, 2]]), K1, S, dtype=np.float64)[0, 0]
    rows, cols = np.nonzero(m > 0.5)
    assert rows.min() >= 20 and rows.max() <= 27 and cols.min() >= 2 and cols.max() <= 9
    assert m[21, 3] == 1.0 and m[5, 25] == 0.0


def test_closed_form_values_near_an_edge():
    # vertical edge x = 12.5 px exactly through the centres of column 12: d = 0 -> sigmoid(0) = 0.5;
    # column 13 lies 1 px = 2/S NDC outside: d^2 = (2/S)^2 > blur_radius -> no candidate -> 0
    sigma, blur = 1e-5, so.blur_radius()
    m = so.soft_silhouette(_tri([[2.5, 2], [12.5, 2], [12.5, 30]]), np.array([[0, 1, 2]]), K1, S, dtype=np.float64)[0, 0]
    assert (2.0 / S) ** 2 > blur
    assert abs(m[20, 12] - 0.5) < 1e-6 and m[20, 13] == 0.0
    # a pixel a known distance inside: edge x = 12.6 -> column 12 is 0.1 px inside; the other edges are far
    m = so.soft_silhouette(_tri([[2.5, 2], [12.6, 2], [12.6, 30]]), np.array([[0, 1, 2]]), K1, S, dtype=np.float64)[0, 0]
    d2 = (0.1 * 2.0 / S) ** 2
    assert d2 < blur
    assert abs(m[24, 12] - 1.0 / (1.0 + math.exp(-d2 / sigma))) < 1e-6 and 0.9 < m[24, 12] < 0.99
    # ... and outside by 0.1 px (edge x = 12.4): within the blur radius, probability sigmoid(-d^2/sigma)
    m = so.soft_silhouette(_tri([[2.5, 2], [12.4, 2], [12.4, 30]]), np.array([[0, 1, 2]]), K1, S, dtype=np.float64)[0, 0]
    assert abs(m[24, 12] - 1.0 / (1.0 + math.exp(d2 / sigma))) < 1e-6 and 0.01 < m[24, 12] < 0.1


def test_keeps_the_ten_nearest_faces():
    # twelve copies of the same near-miss triangle at depths 1.00 .. 1.11 (scaled so they project identically): every copy
    # has the same probability p at the probed pixel; only faces_per_pixel = 10 of them enter the product
    base = _tri([[2.5, 2], [12.4, 2], [12.4, 30]])[0]
    verts = np.concatenate([base * (1.0 + 0.01 * k) for k in range(12)])[None]
    faces = np.arange(36).reshape(12, 3)
    m, frags, _ = so.soft_silhouette(verts, faces, K1, S, dtype=np.float64, return_fragments=True)
    p = 1.0 / (1.0 + math.exp((0.1 * 2.0 / S) ** 2 / 1e-5))
    assert abs(m[0, 0, 24, 12] - (1.0 - (1.0 - p) ** 10)) < 1e-6 and m[0, 0, 24, 12] > 0.1
    kept = sorted(frags[0][0][24, 12].tolist())
    assert kept == list(range(10))           # the ten nearest, not the ten first or last
    m_rev = so.soft_silhouette(verts, faces[::-1].copy(), K1, S, dtype=np.float64, return_fragments=True)[1][0][0][24, 12]
    assert sorted(m_rev.tolist()) == list(range(2, 12))   # same ten depths when the face order is reversed


def test_backward_matches_central_differences():
    rng = np.random.default_rng(0)
    g = np.stack(np.meshgrid(np.linspace(-0.25, 0.25, 5), np.linspace(-0.25, 0.25, 5), indexing="ij"), -1).reshape(-1, 2)
    V = np.concatenate([g + rng.normal(0, 0.01, g.shape), 1 + rng.normal(0, 0.05, (25, 1))], 1)[None]
    faces = np.array([[a, a + 1, a + 5] for i in range(4) for a in [i * 5 + j for j in range(4)]] +
                     [[a + 1, a + 6, a + 5] for i in range(4) for a in [i * 5 + j for j in range(4)]])
    gm = rng.normal(size=(1, 1, S, S))
    gv = so.soft_silhouette_backward(V, faces, K1, gm, S)
    h = 1e-8
    fd = np.zeros_like(V)
    for v in range(25):
        for k in range(3):
            Vp, Vm = V.copy(), V.copy()
            Vp[0, v, k] += h
            Vm[0, v, k] -= h
            fd[0, v, k] = ((so.soft_silhouette(Vp, faces, K1, S, dtype=np.float64) - so.soft_silhouette(Vm, faces, K1, S, dtype=np.float64)) * gm).sum() / (2 * h)
    assert np.abs(gv).max() > 10.0
    assert np.abs(gv - fd).max() <= 1e-5 * np.abs(gv).max()


def test_render_loss_matches_reference_expression():
    # src/utils/loss_modules.py:146-152 on torch tensors == the numpy restatement
    import torch
    import torch.nn.functional as F

    rng = np.random.default_rng(1)
    p, t, v = rng.random((3, 1, 8, 8)), (rng.random((3, 1, 8, 8)) > 0.5).astype(np.float64), np.array([1.0, 0.0, 1.0])
    ref = F.l1_loss(torch.tensor(p), torch.tensor(t), reduction="none").view(3, -1) * torch.tensor(v)[..., None]
    assert np.allclose(so.render_loss(p, t, v), ref.numpy(), atol=0)
