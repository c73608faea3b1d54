"""Drop-in for the reference's common/body_models.py: `build_mano_aa` and the mesh index constants.

`build_mano_aa(is_rhand, create_transl=False, flat_hand=False)` returns an `nn.Module` whose
`forward(betas=, global_orient=, hand_pose=, transl=None)` yields `.vertices (B,778,3)` and
`.joints (B,21,3)` like `smplx.MANO(use_pca=False)` with the fingertip selector enabled
(SURVEY.md Appendix A), computed by libhands_b200.so.  Buffers are registered under the smplx names
so reference checkpoints load (`v_template, shapedirs, posedirs, J_regressor, lbs_weights, parents,
faces_tensor, pose_mean`).

Constants come from `$MANO_DIR/MANO_{RIGHT,LEFT}.pkl` when present (read lazily, not at import --
body_models.py:90 of the reference reads the env at import and fails without it); otherwise, when
`HANDS_B200_SYNTHETIC_MANO=1` or `synthetic=True`, from the seeded MANO-shaped generator used by the
tests and benchmarks (the licensed files cannot be shipped).
"""
import os
import pickle
from collections import namedtuple

import numpy as np
import torch
import torch.nn as nn

from ..functional import ManoHandle, ManoHeadFunction
from ..synthetic import TIP_IDS, synthetic_mano_buffers

ManoOutput = namedtuple("ManoOutput", ["vertices", "joints", "betas", "global_orient", "hand_pose", "full_pose"])

# wrist-sealing fan (reference body_models.py:35-58): ring order, extra centre vertex = 778
_SEAL_RING = (120, 108, 79, 78, 121, 214, 215, 279, 239, 234, 92, 38, 122, 118, 117, 119)
SEAL_FACES_R = [[_SEAL_RING[i], _SEAL_RING[(i + 1) % 16], 778] for i in range(16)]
CIRCLE_V_ID = np.array(_SEAL_RING[1:] + _SEAL_RING[:1], dtype=np.int64)


def seal_mano_mesh(v3d, faces, is_rhand):
    """v3d (B,778,3), faces (1538,3) -> (B,779,3), (1554,3)  (reference body_models.py:60-72)."""
    fan = torch.as_tensor(SEAL_FACES_R, dtype=torch.long, device=faces.device)
    if not is_rhand:
        fan = fan[:, [1, 0, 2]]  # flip the winding for the left hand
    centre = v3d[:, torch.as_tensor(CIRCLE_V_ID, device=v3d.device)].mean(dim=1, keepdim=True)
    return torch.cat((v3d, centre), dim=1), torch.cat((faces, fan), dim=0)


class MANODecimator:
    """Drop-in for the reference's MANODecimator (common/body_models.py:11-32): the 195x778 ARCTIC mesh down-sampler.
    `downsample(verts (B,778,3), is_right) -> (B,195,3)` = D @ verts.  A plain dense matrix product -- issued as ONE
    library GEMM (cuBLAS through torch.matmul over the (778, 3B) view) instead of the reference's `D.repeat(B,1,1)` + bmm,
    which materialises B copies of D (607 KB each).  Constants come from
    `$DATA_DIR/arctic/data/arctic_data/data/meta/mano_decimator_195.npy` like the reference (read lazily, once, not on every
    construction as src/models/generic/wrapper.py:81 does), or from `data={"D_right": ..., "D_left": ...}`."""

    def __init__(self, data=None):
        if data is None:
            path = f"{os.environ['DATA_DIR']}/arctic/data/arctic_data/data/meta/mano_decimator_195.npy"
            data = np.load(path, allow_pickle=True).item()
        self.data = {k: torch.as_tensor(np.asarray(v), dtype=torch.float32) for k, v in data.items() if "D" in k}

    def downsample(self, verts, is_right):
        flag = "right" if is_right else "left"
        D = self.data[f"D_{flag}"]
        if D.device != verts.device:
            D = self.data[f"D_{flag}"] = D.to(verts.device)
        B = verts.shape[0]
        flat = verts.permute(1, 0, 2).reshape(verts.shape[1], B * 3)           # (778, 3B)
        return torch.matmul(D, flat).reshape(D.shape[0], B, 3).permute(1, 0, 2).contiguous()


def _to_np(x):
    """MANO pickles hold chumpy arrays / scipy sparse matrices; reduce to a dense float array."""
    if hasattr(x, "todense"):
        x = x.todense()
    if hasattr(x, "r"):
        x = x.r
    return np.asarray(x)


def load_mano_pkl(path, flat_hand_mean=False):
    """Read an official MANO_{RIGHT,LEFT}.pkl into the smplx buffer layout."""
    with open(path, "rb") as fh:
        d = pickle.load(fh, encoding="latin1")
    shapedirs = _to_np(d["shapedirs"])[:, :, :10]
    posedirs = _to_np(d["posedirs"])
    hands_mean = np.zeros(45) if flat_hand_mean else _to_np(d["hands_mean"]).reshape(45)
    parents = _to_np(d["kintree_table"])[0].astype(np.int64).copy()
    parents[0] = -1
    t = lambda a: torch.as_tensor(np.ascontiguousarray(a), dtype=torch.float32)  # noqa: E731
    return {
        "v_template": t(_to_np(d["v_template"])),
        "shapedirs": t(shapedirs),
        "posedirs": t(posedirs.reshape(-1, posedirs.shape[-1]).T),
        "J_regressor": t(_to_np(d["J_regressor"])),
        "lbs_weights": t(_to_np(d["weights"])),
        "parents": torch.as_tensor(parents),
        "pose_mean": t(np.concatenate([np.zeros(3), hands_mean])),
        "faces": torch.as_tensor(_to_np(d["f"]).astype(np.int64)),
        "tip_ids": torch.tensor(TIP_IDS, dtype=torch.int64),
    }


class VertexJointSelector(nn.Module):
    """smplx's vertex_joint_selector: holds the finger-tip vertex ids (vertex_ids['mano']) as the buffer
    `extra_joints_idxs`; the gather itself happens inside the skinning kernel (bit-exact copy)."""

    def __init__(self, tip_ids):
        super().__init__()
        self.register_buffer("extra_joints_idxs", torch.as_tensor(tip_ids).long().clone())

    def forward(self, vertices, joints):
        return torch.cat([joints, vertices.index_select(1, self.extra_joints_idxs)], dim=1)


class MANOLayer(nn.Module):
    """B200-native stand-in for `smplx.MANO(model_path, use_pca=False, is_rhand=..., flat_hand_mean=...)`."""

    NUM_HAND_JOINTS = 15

    def __init__(self, buffers, is_rhand=True, create_transl=False):
        super().__init__()
        self.is_rhand = is_rhand
        # state_dict key set of smplx.MANO(use_pca=False) [smplx-recalled, SURVEY.md App. A]: buffers faces_tensor, v_template,
        # shapedirs, J_regressor, posedirs, parents, lbs_weights, hand_mean, pose_mean, vertex_joint_selector.extra_joints_idxs;
        # parameters betas, global_orient, hand_pose (+ transl).  The reference loads checkpoints strictly
        # (common/abstract_pl.py:42-44), so nothing else may be persistent here.
        for name in ("v_template", "shapedirs", "posedirs", "J_regressor", "lbs_weights", "pose_mean"):
            self.register_buffer(name, buffers[name].float().clone())
        self.register_buffer("hand_mean", buffers["pose_mean"][3:].float().clone())
        self.register_buffer("parents", buffers["parents"].long().clone())
        self.register_buffer("faces_tensor", buffers["faces"].long().clone())
        self.vertex_joint_selector = VertexJointSelector(buffers["tip_ids"])
        self.faces = buffers["faces"].cpu().numpy()
        # smplx keeps 1-row default parameters in parameters()/state_dict()
        self.betas = nn.Parameter(torch.zeros(1, 10))
        self.global_orient = nn.Parameter(torch.zeros(1, 3))
        self.hand_pose = nn.Parameter(torch.zeros(1, 45))
        if create_transl:
            self.transl = nn.Parameter(torch.zeros(1, 3))
        self._handles = {}

    @property
    def tip_ids(self):
        return self.vertex_joint_selector.extra_joints_idxs

    _CONST_NAMES = ("v_template", "shapedirs", "posedirs", "J_regressor", "lbs_weights", "parents", "pose_mean")

    def _const_buffers(self):
        bufs = {k: getattr(self, k) for k in self._CONST_NAMES}
        bufs["tip_ids"] = self.tip_ids
        return bufs

    def handle(self, device):
        """hb_mano* for `device`.  The device-side constant blob is rebuilt whenever a buffer it was made from has been
        replaced or edited in place (load_state_dict, .to(), manual edits): the cache key holds every buffer's
        (data_ptr, _version)."""
        if device.type != "cuda":
            raise RuntimeError("hands_b200 MANO layer has no CPU path; move the module and inputs to a CUDA device")
        bufs = self._const_buffers()
        stamp = tuple((t.data_ptr(), t._version) for t in bufs.values())
        key = (device.type, device.index if device.index is not None else torch.cuda.current_device())
        hit = self._handles.get(key)
        if hit is None or hit[0] != stamp:
            hit = (stamp, ManoHandle(bufs, device))
            self._handles[key] = hit
        return hit[1]

    # the handle cache holds ctypes pointers to device memory: never copied or pickled with the module
    def __getstate__(self):
        state = self.__dict__.copy()
        state["_handles"] = {}
        return state

    def _load_from_state_dict(self, state_dict, prefix, local_metadata, strict, missing_keys, unexpected_keys, error_msgs):
        """Strict loading of reference checkpoints (common/abstract_pl.py:42-44): smplx versions differ in which helper
        tensors they register (`body_pose`, `hand_components`, `hand_mean`), so keys under this module's prefix that do not
        exist on one side are neither 'missing' nor 'unexpected'."""
        m0, u0 = len(missing_keys), len(unexpected_keys)
        super()._load_from_state_dict(state_dict, prefix, local_metadata, strict, missing_keys, unexpected_keys, error_msgs)
        tolerated = {prefix + k for k in ("body_pose", "hand_components", "hand_mean", "transl")}
        missing_keys[m0:] = [k for k in missing_keys[m0:] if k not in tolerated]
        unexpected_keys[u0:] = [k for k in unexpected_keys[u0:] if k not in tolerated]
        self._handles = {}

    def forward(self, betas=None, global_orient=None, hand_pose=None, transl=None, return_verts=True, return_full_pose=False, **kwargs):
        ref = next(t for t in (betas, global_orient, hand_pose, self.v_template) if t is not None)
        B = max(t.shape[0] for t in (betas, global_orient, hand_pose) if t is not None) if any(t is not None for t in (betas, global_orient, hand_pose)) else 1
        betas = self.betas.expand(B, -1) if betas is None else betas
        global_orient = self.global_orient.expand(B, -1) if global_orient is None else global_orient
        hand_pose = self.hand_pose.expand(B, -1) if hand_pose is None else hand_pose
        if transl is None and hasattr(self, "transl"):
            transl = self.transl.expand(B, -1)
        if betas.shape[0] != B:
            betas = betas.expand(B, -1)
        pose = torch.cat([global_orient.reshape(B, 3), hand_pose.reshape(B, 45)], dim=1)
        h = self.handle(ref.device if ref.is_cuda else self.v_template.device)
        verts, _, joints, _, _, _ = ManoHeadFunction.apply(h, pose, betas, None, None, transl, None, 0.0, 0.0)
        full_pose = pose + self.pose_mean if return_full_pose else None
        return ManoOutput(verts, joints, betas, global_orient, hand_pose, full_pose)


def _buffers_for(is_rhand, flat_hand, synthetic):
    mano_dir = os.environ.get("MANO_DIR")
    if not synthetic and mano_dir:
        for cand in (mano_dir, os.path.join(mano_dir, "mano")):
            path = os.path.join(cand, "MANO_RIGHT.pkl" if is_rhand else "MANO_LEFT.pkl")
            if os.path.exists(path):
                return load_mano_pkl(path, flat_hand_mean=flat_hand)
    if synthetic or os.environ.get("HANDS_B200_SYNTHETIC_MANO") == "1":
        return synthetic_mano_buffers(is_rhand, flat_hand=flat_hand)
    raise FileNotFoundError(
        "MANO model files not found: set MANO_DIR to the directory holding MANO_RIGHT.pkl/MANO_LEFT.pkl, "
        "or set HANDS_B200_SYNTHETIC_MANO=1 for seeded MANO-shaped constants (tests/benchmarks)"
    )


def build_mano_aa(is_rhand, create_transl=False, flat_hand=False, synthetic=False):
    """Same signature as the reference's build_mano_aa (common/body_models.py:92-99)."""
    return MANOLayer(_buffers_for(is_rhand, flat_hand, synthetic), is_rhand=is_rhand, create_transl=create_transl)


def build_layers(device=None, synthetic=False):
    """Reference build_layers (body_models.py:75-88) minus the ARCTIC object tensors (out of scope)."""
    layers = {"right": build_mano_aa(True, synthetic=synthetic), "left": build_mano_aa(False, synthetic=synthetic)}
    if device is not None:
        layers = {k: v.to(device) for k, v in layers.items()}
    return layers
