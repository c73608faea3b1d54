"""Side measurements for DESIGN.md: BASELINE.json configs C2 (MANO head fwd+bwd, B=1024) and C3 (PCL fwd+bwd, 1024 crops)
on one B200, plus the same step replayed from a CUDA graph (launch-latency bound at these sizes).  Not the bench line."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from hands_b200.step import GeometryStep  # noqa: E402


def cuda_time(fn, reps=50, warm=5):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e-3 / reps


def graphed(fn):
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        for _ in range(3):
            fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g, stream=s):
        fn()
    return g.replay


def main():
    dev = torch.device("cuda:0")
    out = {}
    for B in (1024, 8192, 65536):
        st = GeometryStep(B, dev, with_pcl=False, hands_per_sample=1)

        def mano():
            st.mano_forward(0)
            st.mano_backward(0)

        t = cuda_time(mano)
        out[f"C2 MANO head fwd+bwd B={B}"] = {"ms": t * 1e3, "hands_per_s": B / t, "hbm_frac": B / t * 31068 / 6557.8e9}
        if B == 1024:
            tg = cuda_time(graphed(mano))
            out[f"C2 MANO head fwd+bwd B={B} (CUDA graph replay)"] = {"ms": tg * 1e3, "hands_per_s": B / tg, "hbm_frac": B / tg * 31068 / 6557.8e9}
        del st
    for n in (1024, 16384):
        st = GeometryStep(n, dev, with_mano=False, hands_per_sample=1)
        st.pcl_setup()

        def pcl():
            st.pcl_forward()
            st.pcl_backward()

        t = cuda_time(pcl, reps=20)
        by = n * (3 * 224 * 224 * 4 * 3 + 12 * st.mean_s2)
        out[f"C3 PCL fwd+bwd {n} crops"] = {"ms": t * 1e3, "crops_per_s": n / t, "hbm_frac": by / t / 6557.8e9}
        del st
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
