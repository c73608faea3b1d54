"""Per-kernel times of the PCL path on one B200: forward, backward-mid, backward-img, each timed alone over one
1024-image launch (2 crops per image), CUDA events.  Development aid; not the bench line."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from hands_b200.step import GeometryStep  # noqa: E402


def cuda_time(fn, reps=20, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / reps


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
    st = GeometryStep(n, torch.device("cuda:0"), with_mano=False)
    st.pcl_setup()
    st.pcl_forward()
    st.pcl_backward()
    out = {"images": n,
           "fwd_us": cuda_time(st.pcl_forward),
           "mid_us": cuda_time(lambda: st.pcl_backward_stage(1))}
    # transposed grid_sample: scatter form (default) and gather form, same workspace contents
    ref = None
    for name, on in (("img_scatter_us", 1), ("img_gather_us", 0)):
        prev = st.lib.hb_pcl_set_scatter(on)
        out[name] = cuda_time(lambda: st.pcl_backward_stage(2))
        g = st.g_img.clone()
        st.lib.hb_pcl_set_scatter(prev)
        if ref is None:
            ref = g
        else:
            out["scatter_vs_gather_rel"] = float((ref - g).abs().max() / g.abs().max())
    out["eligible_frac"] = float((st.params[:, 26].view(torch.int32) > 0).float().mean())
    out["phases_hist"] = torch.bincount(st.params[:, 26].view(torch.int32).cpu(), minlength=7).tolist()
    print(json.dumps(out))


if __name__ == "__main__":
    main()
