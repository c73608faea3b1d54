"""CPU ORACLE — TEST INFRASTRUCTURE ONLY.  Never imported by the product (`hands_b200/`).

Soft-silhouette rendering of the MANO mesh: the consumer of `mano.v3d.cam.{r,l}` in the
reference, `src/models/hands_light/renderer.py:124-199` (`DiffRenderer` + `MANORenderer.forward`),
feeding `render_loss` (`src/utils/loss_modules.py:146-152`, `src/callbacks/loss/loss_arctic_sf.py:172-183`).

**Parity unpinned.**  The arithmetic lives in the third-party package `pytorch3d` (imported at
`renderer.py:8-9`; the reference pins no version), which is neither under `/root/reference` nor
installed here, and the reference holds no test or golden vector for it.  This file restates the
published algorithm of that package as the reference configures it:

* `PerspectiveCameras(focal, principal)` in NDC, R = I, T = 0 (`renderer.py:187-191`):
  `x = (fx'·X + px'·Z)/Z`, `y = (fy'·Y + py'·Z)/Z` with `K' = intrx_to_ndc · K`
  (`renderer.py:171-175,187-190`); the rasteriser keeps the view-space depth `Z` as `z`
  (`MeshRasterizer.transform`).
* `rasterize_meshes` (pytorch3d `csrc/rasterize_meshes`, `csrc/utils/geometry_utils.cuh`) with
  `blur_radius = log(1/1e-6 - 1)·sigma`, `faces_per_pixel = 10`, `perspective_correct = False`
  (`renderer.py:130-137`), `clip_barycentric_coords = True` (its default when `blur_radius > 0`),
  no back-face culling, no z clipping: per pixel centre, per face — bounding-box test widened by
  `sqrt(blur_radius)`, zero-area test (`kEpsilon = 1e-8`), barycentric coordinates, clipped
  barycentric depth `pz`, squared distance to the nearest edge segment, signed negative inside,
  rejected when outside and `dist >= blur_radius`; the `K` candidates of smallest `pz` are kept
  (a later face replaces the current farthest only when strictly nearer).
* `SoftSilhouetteShader` = `sigmoid_alpha_blend`: `mask = 1 - Π_k (1 - sigmoid(-dist_k / sigma))`.
* `flip_transpose_canvas` (`renderer.py:201-209`) undoes pytorch3d's +X-left/+Y-up image
  convention, so that `mask[b, 0, r, c]` is the coverage at the pixel centre `(c + 0.5, r + 0.5)`
  of the ordinary `K`-projected image; its NDC coordinate is `-1 + (2c + 1)/S` in fp32.

Only the distances carry gradient to the vertices (the shader reads nothing else): the backward
below is `PointTriangleDistanceBackward` on the nearest edge followed by the projection's Jacobian.
"""
import math

import numpy as np

K_EPS = 1e-8


def blur_radius(sigma=1e-5, dist_eps=1e-6):
    """renderer.py:128-133"""
    return math.log(1.0 / dist_eps - 1.0) * sigma


def ndc_intrinsics(K, img_res, dtype=np.float32):
    """renderer.py:171-175,187-190: K' = intrx_to_ndc @ K; focal = diag(K')[:2]; principal = K'[:2, 2]"""
    K = np.asarray(K, dtype)
    to_ndc = np.array([[2.0 / img_res, 0, -1], [0, 2.0 / img_res, -1], [0, 0, 1]], dtype)
    Kn = (to_ndc[None] @ K).astype(dtype)
    return Kn[:, 0, 0], Kn[:, 1, 1], Kn[:, 0, 2], Kn[:, 1, 2]


def project_ndc(verts, K, img_res, dtype=np.float32):
    """(B,V,3) camera-space -> (B,V,3) = (x_ndc, y_ndc, Z)   (PerspectiveCameras projection, homogeneous divide)"""
    v = np.asarray(verts, dtype)
    fx, fy, px, py = ndc_intrinsics(K, img_res, dtype)
    X, Y, Z = v[..., 0], v[..., 1], v[..., 2]
    x = ((X * fx[:, None]).astype(dtype) + (Z * px[:, None]).astype(dtype)).astype(dtype) / Z
    y = ((Y * fy[:, None]).astype(dtype) + (Z * py[:, None]).astype(dtype)).astype(dtype) / Z
    return np.stack([x.astype(dtype), y.astype(dtype), Z], -1)


def pix_to_ndc(S, dtype=np.float32):
    i = np.arange(S).astype(dtype)
    return (dtype(-1.0) + (dtype(2.0) * i + dtype(1.0)) / dtype(S)).astype(dtype)


def _edge(px, py, ax, ay, bx, by):
    return (px - ax) * (by - ay) - (py - ay) * (bx - ax)


def _seg_dist(px, py, ax, ay, bx, by):
    """squared distance from p to segment ab + the clamped parameter (PointLineDistanceForward)"""
    bax, bay = bx - ax, by - ay
    l2 = bax * bax + bay * bay
    if l2 <= K_EPS:
        dx, dy = px - bx, py - by
        return dx * dx + dy * dy, np.ones_like(px), True
    t = (bax * (px - ax) + bay * (py - ay)) / l2
    tt = np.clip(t, 0.0, 1.0)
    dx, dy = px - (ax + tt * bax), py - (ay + tt * bay)
    return dx * dx + dy * dy, tt, False


def rasterize_one(vn, faces, S, blur, Kf, dtype=np.float32):
    """One mesh.  vn (V,3) projected vertices, faces (F,3).  Returns per pixel the kept fragments:
    face index (S,S,Kf) int (-1 empty), signed squared distance (S,S,Kf), depth pz (S,S,Kf)."""
    pf = np.full((S, S, Kf), -1, np.int64)
    dist = np.full((S, S, Kf), -1.0, dtype)
    zb = np.full((S, S, Kf), -1.0, dtype)
    size = np.zeros((S, S), np.int64)
    cnt = np.zeros((S, S), np.int64)
    znext = np.full((S, S), np.inf)
    pn = pix_to_ndc(S, dtype)
    sb = dtype(math.sqrt(blur))
    blur = dtype(blur)
    eps = dtype(K_EPS)
    for f in range(faces.shape[0]):
        v0, v1, v2 = (vn[faces[f, k]].astype(dtype) for k in range(3))
        zmax = max(v0[2], v1[2], v2[2])
        area = _edge(v0[0], v0[1], v1[0], v1[1], v2[0], v2[1])
        if not np.isfinite(area) or zmax < 0 or (-eps <= area <= eps):
            continue
        xlo, xhi = min(v0[0], v1[0], v2[0]) - sb, max(v0[0], v1[0], v2[0]) + sb
        ylo, yhi = min(v0[1], v1[1], v2[1]) - sb, max(v0[1], v1[1], v2[1]) + sb
        cs = np.nonzero((pn >= xlo) & (pn <= xhi))[0]
        rs = np.nonzero((pn >= ylo) & (pn <= yhi))[0]
        if cs.size == 0 or rs.size == 0:
            continue
        py, px = np.meshgrid(pn[rs], pn[cs], indexing="ij")
        rr, cc = np.meshgrid(rs, cs, indexing="ij")
        barea = _edge(v2[0], v2[1], v0[0], v0[1], v1[0], v1[1]) + eps
        w0 = _edge(px, py, v1[0], v1[1], v2[0], v2[1]) / barea
        w1 = _edge(px, py, v2[0], v2[1], v0[0], v0[1]) / barea
        w2 = _edge(px, py, v0[0], v0[1], v1[0], v1[1]) / barea
        c0, c1, c2 = (np.clip(w, 0.0, 1.0) for w in (w0, w1, w2))
        ws = np.maximum(c0 + c1 + c2, dtype(1e-5))
        pz = (c0 / ws) * v0[2] + (c1 / ws) * v1[2] + (c2 / ws) * v2[2]
        d01, _, _ = _seg_dist(px, py, v0[0], v0[1], v1[0], v1[1])
        d02, _, _ = _seg_dist(px, py, v0[0], v0[1], v2[0], v2[1])
        d12, _, _ = _seg_dist(px, py, v1[0], v1[1], v2[0], v2[1])
        d = np.minimum(np.minimum(d01, d02), d12)
        inside = (w0 > 0) & (w1 > 0) & (w2 > 0)
        cand = (pz >= 0) & (inside | (d < blur))
        sd = np.where(inside, -d, d).astype(dtype)
        for r, c, z, s in zip(rr[cand], cc[cand], pz[cand].astype(dtype), sd[cand]):
            n = size[r, c]
            cnt[r, c] += 1
            if n < Kf:
                pf[r, c, n], dist[r, c, n], zb[r, c, n] = f, s, z
                size[r, c] = n + 1
            else:
                k = int(np.argmax(zb[r, c]))   # first of the farthest, as the running-max bookkeeping keeps it
                if z < zb[r, c, k]:
                    znext[r, c] = min(znext[r, c], zb[r, c, k])
                    pf[r, c, k], dist[r, c, k], zb[r, c, k] = f, s, z
                else:
                    znext[r, c] = min(znext[r, c], z)
    gap = np.where(cnt > Kf, znext - np.where(pf >= 0, zb, -np.inf).max(-1), np.inf)
    return pf, dist, zb, cnt, gap


def soft_silhouette(verts_cam, faces, K, img_res=224, sigma=1e-5, dist_eps=1e-6, faces_per_pixel=10, dtype=np.float32,
                    return_fragments=False):
    """MANORenderer.forward(...)['mask']  (renderer.py:178-199): (B,V,3),(F,3),(B,3,3) -> (B,1,S,S)"""
    S = int(img_res)
    vn = project_ndc(verts_cam, K, img_res, dtype)
    blur = blur_radius(sigma, dist_eps)
    B = vn.shape[0]
    mask = np.zeros((B, 1, S, S), dtype)
    frags = []
    for b in range(B):
        pf, dist, zb, cnt, gap = rasterize_one(vn[b], np.asarray(faces), S, blur, faces_per_pixel, dtype)
        with np.errstate(over="ignore"):
            prob = 1.0 / (1.0 + np.exp(dist.astype(np.float64) / sigma))   # sigmoid(-dist/sigma)
        prob = (prob * (pf >= 0)).astype(dtype)
        alpha = np.prod(1.0 - prob, axis=-1).astype(dtype)
        mask[b, 0] = 1.0 - alpha
        frags.append((pf, dist, zb, prob, alpha, cnt, gap))
    return (mask, frags, vn) if return_fragments else mask


def soft_silhouette_backward(verts_cam, faces, K, g_mask, img_res=224, sigma=1e-5, dist_eps=1e-6, faces_per_pixel=10,
                             dtype=np.float64):
    """d(sum(g_mask * mask)) / d verts_cam, the way autograd walks sigmoid_alpha_blend -> rasterize_meshes backward
    (distance term only) -> the camera projection.  Returns (B,V,3)."""
    S = int(img_res)
    faces = np.asarray(faces)
    verts = np.asarray(verts_cam, dtype)
    _, frags, vn = soft_silhouette(verts, faces, K, img_res, sigma, dist_eps, faces_per_pixel, dtype, return_fragments=True)
    fx, fy, px_, py_ = ndc_intrinsics(K, img_res, dtype)
    pn = pix_to_ndc(S, dtype)
    g_verts = np.zeros_like(verts)
    for b, (pf, dist, zb, prob, alpha, _cnt, _gap) in enumerate(frags):
        g_ndc = np.zeros((verts.shape[1], 2), dtype)
        gm = np.asarray(g_mask, dtype)[b, 0]
        rs, cs, ks = np.nonzero(pf >= 0)
        for r, c, k in zip(rs, cs, ks):
            p = prob[r, c, k]
            others = np.prod(np.delete(1.0 - prob[r, c], k))
            g_sd = gm[r, c] * others * (-p * (1.0 - p) / sigma)       # d mask / d signed dist
            if g_sd == 0.0:
                continue
            g_d = g_sd * (-1.0 if dist[r, c, k] < 0 else 1.0)          # signed -> absolute distance
            f = pf[r, c, k]
            ids = faces[f]
            v = vn[b][ids][:, :2]
            P = np.array([pn[c], pn[r]], dtype)
            d = [_seg_dist(P[0], P[1], v[i][0], v[i][1], v[j][0], v[j][1]) for i, j in ((0, 1), (0, 2), (1, 2))]
            e = (0, 1) if (d[0][0] <= d[1][0] and d[0][0] <= d[2][0]) else ((0, 2) if (d[1][0] <= d[0][0] and d[1][0] <= d[2][0]) else (1, 2))
            dd, tt, degenerate = d[{(0, 1): 0, (0, 2): 1, (1, 2): 2}[e]]
            a, bb = v[e[0]], v[e[1]]
            if degenerate:
                g_ndc[ids[e[1]]] += g_d * 2.0 * (bb - P)
                continue
            proj = a + tt * (bb - a)
            g_ndc[ids[e[0]]] += g_d * (1.0 - tt) * 2.0 * (proj - P)
            g_ndc[ids[e[1]]] += g_d * tt * 2.0 * (proj - P)
        X, Y, Z = verts[b, :, 0], verts[b, :, 1], verts[b, :, 2]
        g_verts[b, :, 0] = g_ndc[:, 0] * fx[b] / Z
        g_verts[b, :, 1] = g_ndc[:, 1] * fy[b] / Z
        g_verts[b, :, 2] = -(g_ndc[:, 0] * fx[b] * X + g_ndc[:, 1] * fy[b] * Y) / (Z * Z)
    return g_verts


def render_loss(pred_mask, gt_mask, is_valid):
    """src/utils/loss_modules.py:146-152 (return_mean=False): |pred - gt| per pixel, gated per sample -> (B, S*S)"""
    bz = pred_mask.shape[0]
    return np.abs(pred_mask - gt_mask).reshape(bz, -1) * np.asarray(is_valid).reshape(bz, 1)
