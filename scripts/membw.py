"""HBM bandwidth of the three access mixes the PCL kernels have, with plain torch ops (development aid): write-only
(fill_), read-only (sum), copy.  The roofline denominator (MEASURED_PEAKS.json) is a copy; a write-dominated kernel such as
the crop forward is bounded by the write-only figure."""
import json

import torch


def t(fn, reps=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) * 1e-3)
    return best


def main():
    n = 1 << 30   # 4 GiB of fp32
    a = torch.empty(n, device="cuda")
    b = torch.empty(n, device="cuda")
    a.normal_()
    out = {"bytes": 4 * n}
    out["write_only_fill_GBs"] = 4 * n / t(lambda: b.fill_(1.0)) / 1e9
    out["write_only_zero_GBs"] = 4 * n / t(lambda: b.zero_()) / 1e9
    out["read_only_sum_GBs"] = 4 * n / t(lambda: a.sum()) / 1e9
    out["copy_GBs_read_plus_write"] = 8 * n / t(lambda: b.copy_(a)) / 1e9
    out["add_2r1w_GBs"] = 12 * n / t(lambda: torch.add(a, b, out=b)) / 1e9
    print(json.dumps(out))


if __name__ == "__main__":
    main()
