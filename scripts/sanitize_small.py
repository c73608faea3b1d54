"""Small forward+backward calls of every kernel family for compute-sanitizer (memcheck / racecheck): a few crops in both
forward modes and with a uint8 source, boxes touching the image border, an empty box, a generic resolution; one MANO head
fwd+bwd per pose format; the loss kernels."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from hands_b200 import _lib  # noqa: E402
from hands_b200.losses import keypoint_losses, vector_loss  # noqa: E402
from hands_b200.pcl import perspective_crop  # noqa: E402
from hands_b200.src.nets.hand_heads.mano_head import MANOHead  # noqa: E402
from hands_b200.synthetic import synthetic_head_inputs, synthetic_pcl_inputs  # noqa: E402

dev = torch.device("cuda:0")
lib = _lib.load()
for res, cpi in ((224, 2), (96, 1), (50, 1)):
    B = 3
    n = B * cpi
    img, bbox, K = synthetic_pcl_inputs(n, seed=res, img_res=res, smin=res // 4, smax=3 * res // 4)
    img = img[:B].contiguous()
    bbox[0] = torch.tensor([0, 0, res // 3, res // 4])
    bbox[1] = torch.tensor([res - 1 - res // 3, res - 1 - res // 4, res - 1, res - 1])
    bbox[2] = torch.tensor([5, 5, 5, 5])
    for exact in (0, 1):
        lib.hb_pcl_set_exact(exact)
        x = img.to(dev).requires_grad_(True)
        crop, rot = perspective_crop(x, bbox.to(dev), K.to(dev), img_res=res, crops_per_img=cpi)
        crop.sum().backward()
    lib.hb_pcl_set_exact(0)
    if res % 16 == 0:
        u8 = torch.randint(0, 256, (B, 3, res, res), dtype=torch.uint8, device=dev)
        perspective_crop(u8, bbox.to(dev), K.to(dev), img_res=res, crops_per_img=cpi, mean=(0.485, 0.456, 0.406), std=(0.229, 0.224, 0.225))
head = MANOHead(True, 1000.0, 224.0, synthetic=True).to(dev)
for B in (5, 130):
    rotmat, betas, cam, K = [t.to(dev) for t in synthetic_head_inputs(B, seed=B)]
    r = rotmat.clone().requires_grad_(True)
    o = head(r, betas, cam, K)
    (o["v3d.cam.r"].sum() + o["j2d.norm.r"].sum()).backward()
    x6 = torch.randn(B, 16, 6, device=dev, requires_grad=True)
    o = head.forward_rot6d(x6, betas, cam, K, layout="cols")
    l3, l2, _ = keypoint_losses(o["j3d.cam.r"], o["j2d.norm.r"], torch.zeros(B, 21, 3, device=dev), torch.zeros(B, 21, 2, device=dev), torch.ones(B, 21, device=dev))
    lv = vector_loss(o["cam_t.wp.r"], torch.zeros(B, 3, device=dev), torch.ones(B, device=dev), None, pred2=cam)
    (l3 + l2 + lv).backward()
# soft-silhouette consumer: two image sizes (one no multiple of the tile), a degenerate and an off-screen face, mask loss
from hands_b200.losses import render_loss  # noqa: E402
from hands_b200.src.models.hands_light.renderer import MANORenderer  # noqa: E402
from hands_b200.synthetic import synthetic_silhouette_inputs  # noqa: E402

for S in (224, 52):
    vc, faces, K = synthetic_silhouette_inputs(2, seed=S, img_res=S)
    K[:, 0, 0] *= S / 224.0
    K[:, 1, 1] *= S / 224.0
    vc[0, 5] = vc[0, 6]              # zero-area faces
    vc[1, 700:] += 5.0               # part of the mesh far off screen
    rdr = MANORenderer({"img_res": S}, faces_r=faces.numpy(), faces_l=faces.numpy()).to(dev)
    v = vc.to(dev).requires_grad_(True)
    m = rdr({"mano.v3d.cam.r": v}, {"intrinsics": K.to(dev)}, is_right=True)["mask"]
    render_loss(m, torch.zeros_like(m), torch.ones(2, device=dev)).backward()
torch.cuda.synchronize()
print("sanitize_small: done")
