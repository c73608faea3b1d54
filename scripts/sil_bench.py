"""Time the soft-silhouette renderer (forward, backward) on synthetic hand-sized meshes.  python scripts/sil_bench.py [B]"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from hands_b200.functional import SilhouetteHandle, SoftSilhouetteFunction  # noqa: E402
from hands_b200.synthetic import synthetic_silhouette_inputs  # noqa: E402
from scripts.bench_legs import cuda_time  # noqa: E402


def run(B=1024, dev="cuda:0", iters=10):
    import math

    vc, faces, K = synthetic_silhouette_inputs(B, seed=0)
    h = SilhouetteHandle(faces, 778, dev)
    v = vc.to(dev).requires_grad_(True)
    K = K.to(dev)
    sigma, blur = 1e-5, math.log(1.0 / 1e-6 - 1.0) * 1e-5
    g = torch.randn(B, 1, 224, 224, device=dev)
    state = {}

    def fwd():
        state["m"] = SoftSilhouetteFunction.apply(h, v, K, 224, sigma, blur)

    def bwd():
        (state["g"],) = torch.autograd.grad(state["m"], v, g, retain_graph=True)

    t_f = cuda_time(fwd, iters, 3, dev) * 1e3
    t_b = cuda_time(bwd, iters, 3, dev) * 1e3
    cov = float((state["m"] > 0.5).float().mean())
    return {"meshes": B, "fwd_ms": t_f, "bwd_ms": t_b, "meshes_per_s": B / ((t_f + t_b) * 1e-3), "coverage": cov,
            "out_bytes_per_mesh": 224 * 224 * 4}


if __name__ == "__main__":
    print(json.dumps(run(int(sys.argv[1]) if len(sys.argv) > 1 else 1024)))
