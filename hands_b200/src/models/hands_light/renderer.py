"""Drop-in for the reference's src/models/hands_light/renderer.py:124-199 (`DiffRenderer`, `MANORenderer`).

Same constructor argument (`args` with `img_res`), same `forward(mano_output, meta_info, is_right=True)` reading
`mano.v3d.cam.{r,l}` and `meta_info['intrinsics']`, same result dictionary (`image` all ones, `mask` (B,1,S,S));
pytorch3d's cameras, `Meshes`, rasteriser, shader and the canvas flip are two CUDA launches forward and two backward
(hands_b200/csrc/silhouette.cu).  The face tables default to the MANO layers' own (`build_mano_aa(...).faces`, the
alternative the reference names at renderer.py:155-156) instead of the un-shipped `default_mano_faces.pkl`.
"""
import math

import torch
import torch.nn as nn

from ....functional import SilhouetteHandle, SoftSilhouetteFunction


class DiffRenderer(nn.Module):
    """renderer.py:124-160: BlendParams(sigma=1e-5, gamma=1e-4), dist_eps=1e-6, faces_per_pixel=10."""

    SIGMA = 1e-5
    DIST_EPS = 1e-6

    def __init__(self, args):
        super().__init__()
        self.args = args
        self.img_res = int(args["img_res"] if isinstance(args, dict) else args.img_res)
        self.sigma = self.SIGMA
        self.blur_radius = math.log(1.0 / self.DIST_EPS - 1.0) * self.sigma
        self._handles = {}

    def _handle(self, faces, n_verts, device):
        """hb_sil* for this face table on `device`; rebuilt when the table is replaced or edited in place."""
        faces = torch.as_tensor(faces)
        key = (int(n_verts), device.type, device.index if device.index is not None else torch.cuda.current_device())
        stamp = (faces.data_ptr(), faces._version, tuple(faces.shape))
        hit = self._handles.get(key)
        if hit is None or hit[0] != stamp:
            hit = self._handles[key] = (stamp, SilhouetteHandle(faces, n_verts, device))
        return hit[1]

    # the handle cache holds ctypes pointers to device memory: never copied or pickled with the module
    def __getstate__(self):
        state = self.__dict__.copy()
        state["_handles"] = {}
        return state

    def forward(self, verts_cam, faces, K):
        """verts_cam (B,V,3) camera space, faces (F,3) shared by the batch, K (B,3,3) pixel intrinsics."""
        handle = self._handle(faces, verts_cam.shape[1], verts_cam.device)
        mask = SoftSilhouetteFunction.apply(handle, verts_cam, K, self.img_res, self.sigma, self.blur_radius)
        # SoftSilhouetteShader paints every pixel white (renderer.py:157: "all 1s ... only care about mask")
        return {"image": torch.ones(mask.shape[0], 3, self.img_res, self.img_res, dtype=mask.dtype, device=mask.device), "mask": mask}


class MANORenderer(nn.Module):
    def __init__(self, args, faces_r=None, faces_l=None, synthetic=False):
        super().__init__()
        self.args = args
        self.renderer = DiffRenderer(args)
        if faces_r is None or faces_l is None:
            from ....common.body_models import build_mano_aa

            faces_r = build_mano_aa(True, synthetic=synthetic).faces if faces_r is None else faces_r
            faces_l = build_mano_aa(False, synthetic=synthetic).faces if faces_l is None else faces_l
        self.register_buffer("mano_faces_r", torch.as_tensor(faces_r.astype("int64") if hasattr(faces_r, "astype") else faces_r).long(), persistent=False)
        self.register_buffer("mano_faces_l", torch.as_tensor(faces_l.astype("int64") if hasattr(faces_l, "astype") else faces_l).long(), persistent=False)

    def forward(self, mano_output, meta_info, is_right=True):
        vertices = mano_output["mano.v3d.cam.r" if is_right else "mano.v3d.cam.l"]
        faces = self.mano_faces_r if is_right else self.mano_faces_l
        K = meta_info["intrinsics"].to(vertices.device, torch.float32)
        return self.renderer(vertices, faces, K)
