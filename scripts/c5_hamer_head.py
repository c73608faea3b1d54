"""BASELINE.json config C5: HaMeR-light MANO head training step with the full gradient all-reduce.

Per GPU: synthetic `token_out` (8192, 1024) -> decpose / decshape / deccam read-outs (src/models/hamer_light/mano_head.py:38-40,
111,725 parameters per hand side) + mean-parameter init -> 6D pose fused into the MANO head (both hand sides) -> key-point
loss terms (loss_arctic_sf.py:70-158 shape; hands_b200.losses) -> backward -> NCCL all-reduce of the parameter gradients
(+ optionally a 39,497,837-parameter stand-in for the HaMeR decoder head, --full-head) + ONE packed metric all-reduce.

    python scripts/c5_hamer_head.py [--batch 8192] [--steps 20] [--full-head]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 scripts/c5_hamer_head.py

The Linears are plain library GEMMs (torch / cuBLAS); everything downstream is this repo's kernels.  Side measurement for
DESIGN.md, not the bench line.
"""
import argparse
import json
import os
import sys

import torch
import torch.distributed as dist
import torch.nn as nn

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from hands_b200.distributed import PackedMetrics, allreduce_gradients  # noqa: E402
from hands_b200.losses import keypoint_losses, mrrpe_sums  # noqa: E402
from hands_b200.src.nets.hand_heads.mano_head import MANOHead  # noqa: E402


class Readouts(nn.Module):
    """decpose / decshape / deccam of hamer_light/mano_head.py:38-40 with the mean-parameter buffers (:54-60)."""

    def __init__(self, dim=1024):
        super().__init__()
        self.decpose, self.decshape, self.deccam = nn.Linear(dim, 96), nn.Linear(dim, 10), nn.Linear(dim, 3)
        for m in (self.decpose, self.decshape, self.deccam):
            nn.init.xavier_uniform_(m.weight, gain=0.01)
        eye6 = torch.tensor([1.0, 0.0, 0.0, 0.0, 1.0, 0.0]).repeat(16)
        self.register_buffer("init_hand_pose", eye6[None])
        self.register_buffer("init_betas", torch.zeros(1, 10))
        self.register_buffer("init_cam", torch.tensor([[0.9, 0.0, 0.0]]))

    def forward(self, token_out):
        return self.decpose(token_out) + self.init_hand_pose, self.decshape(token_out) + self.init_betas, self.deccam(token_out) + self.init_cam


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=8192)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--full-head", action="store_true", help="add a 39,497,837-parameter gradient to the all-reduce (158 MB)")
    args = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    res = run(dev, world, rank, args.batch, args.steps, args.warmup, args.full_head)
    if rank == 0:
        print(json.dumps(res))
    if world > 1:
        dist.destroy_process_group()


def run(dev, world, rank, batch=8192, steps=20, warmup=3, full_head=False):
    """One measurement (all ranks call it; the process group already exists when world > 1).  Returns the result dict."""
    args = argparse.Namespace(batch=batch, steps=steps, warmup=warmup, full_head=full_head)
    torch.manual_seed(0)  # same initial parameters on every rank (what DDP's broadcast gives)
    B = args.batch
    heads = {True: MANOHead(True, 1000.0, 224, synthetic=True).to(dev), False: MANOHead(False, 1000.0, 224, synthetic=True).to(dev)}
    ro = {True: Readouts().to(dev), False: Readouts().to(dev)}
    params = [p for m in ro.values() for p in m.parameters()]
    extra = None
    if args.full_head:
        extra = nn.Parameter(torch.zeros(39_497_837 - sum(p.numel() for p in params), device=dev))
        extra.grad = torch.zeros_like(extra)
        params.append(extra)
    g = torch.Generator(device=dev).manual_seed(100 + rank)   # each rank its own shard of the synthetic data
    token = torch.randn(B, 1024, generator=g, device=dev)
    K = torch.tensor([[1000.0, 0, 112], [0, 1000.0, 112], [0, 0, 1]], device=dev).expand(B, 3, 3).contiguous()
    gt = {s: dict(j3d=0.1 * torch.randn(B, 21, 3, generator=g, device=dev) + torch.tensor([0, 0, 0.6], device=dev),
                  j2d=0.3 * torch.randn(B, 21, 2, generator=g, device=dev), jv=torch.ones(B, 21, device=dev),
                  betas=torch.zeros(B, 10, device=dev)) for s in (True, False)}
    names = ["loss/kp3d/r", "loss/kp2d/r", "loss/kp3d/l", "loss/kp2d/l", "mpjpe/ra/r", "mpjpe/ra/l", "pix_err/r", "pix_err/l", "mrrpe/r/l"]
    metrics = PackedMetrics(names, dev)
    last = {}

    def step():
        for p in params:
            if p is not extra:
                p.grad = None
        total = 0.0
        out = {}
        for side in (True, False):
            pf = "r" if side else "l"
            pose6d, betas, cam = ro[side](token)
            o = heads[side].forward_rot6d(pose6d, betas, cam, K, layout="cols")
            t = gt[side]
            l3, l2, sums = keypoint_losses(o["j3d.cam." + pf], o["j2d.norm." + pf], t["j3d"], t["j2d"], t["jv"])
            total = total + 5.0 * l3 + 5.0 * l2 + 0.001 * ((betas - t["betas"]) ** 2).mean()
            out[pf] = o["j3d.cam." + pf]
            n = float(B * 21)
            metrics.add(f"loss/kp3d/{pf}", sums[0], n * 3)
            metrics.add(f"loss/kp2d/{pf}", sums[1], n * 2)
            metrics.add(f"mpjpe/ra/{pf}", sums[2], sums[3])
            metrics.add(f"pix_err/{pf}", sums[4], sums[5])
        m = mrrpe_sums(out["r"].detach(), out["l"].detach(), gt[True]["j3d"], gt[False]["j3d"])
        metrics.add("mrrpe/r/l", m[0], m[1])
        total.backward()
        allreduce_gradients(params)
        if world > 1:
            dist.all_reduce(metrics.buf)
        last["loss"] = total.detach()

    for _ in range(args.warmup):
        step()
        metrics.buf.zero_()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        step()
    e1.record()
    torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1) / args.steps], device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    grad_bytes = 4 * sum(p.numel() for p in params)
    return {"config": "C5 HaMeR-light MANO head + grad all-reduce", "n_gpus": world, "batch_per_gpu": B, "ms_per_step": float(ms),
            "hands_per_s": 2 * B * world / (float(ms) * 1e-3), "allreduce_bytes_per_step": grad_bytes if world > 1 else 0,
            "loss": float(last["loss"])}


if __name__ == "__main__":
    main()
