"""Tolerance bookkeeping for the GPU parity tests.

Several bounds are `max(stated tolerance, 3 x the fp32 oracle's own distance from the fp64 oracle)` -- the widening exists
for inputs where fp32 conditioning itself limits the reference (angles near pi).  Every such assertion goes through
`tol_check`, which records which branch passed; the session writes the tally to gpurun_out/tolerance_branches.json and
prints it, so a report can say how often the widened branch was the one that held."""
import json
import os

LOG = []


def tol_check(name, err, tol, own=None):
    err = float(err)
    bound = tol if own is None else max(tol, 3.0 * float(own))
    branch = "stated" if err <= tol else ("widened" if err <= bound else "FAIL")
    LOG.append({"name": name, "err": err, "tol": tol, "own": None if own is None else float(own), "branch": branch})
    assert err <= bound, (name, err, tol, own)


def dump(root):
    if not LOG:
        return None
    tally = {"stated": 0, "widened": 0, "FAIL": 0}
    for r in LOG:
        tally[r["branch"]] += 1
    out = {"tally": tally, "widened": [r for r in LOG if r["branch"] != "stated"], "n": len(LOG),
           "worst_vs_stated": sorted(LOG, key=lambda r: -r["err"] / r["tol"])[:10]}
    try:
        d = os.path.join(root, "gpurun_out")
        os.makedirs(d, exist_ok=True)
        with open(os.path.join(d, "tolerance_branches.json"), "w") as fh:
            json.dump(out, fh, indent=1)
    except OSError:
        pass
    return out
