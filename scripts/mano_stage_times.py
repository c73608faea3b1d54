"""MANO head forward / backward of one hand side, B hands, timed alone (CUDA events).  Development aid; also the workload the
ncu captures of the mano_* kernels run (scripts/ncu_mano.sh)."""
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
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
    st = GeometryStep(B, torch.device("cuda:0"), with_pcl=False, hands_per_sample=1)
    out = {"hands": B, "fwd_us": cuda_time(lambda: st.mano_forward(0)), "bwd_us": cuda_time(lambda: st.mano_backward(0))}
    out["hands_per_s_fwd_bwd"] = B / ((out["fwd_us"] + out["bwd_us"]) * 1e-6)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
