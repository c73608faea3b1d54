import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu under gpurun)")


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN


def pytest_terminal_summary(terminalreporter):
    import _tol

    out = _tol.dump(ROOT)
    if out:
        terminalreporter.write_line(f"tolerance branches: {out['tally']} of {out['n']} checks (gpurun_out/tolerance_branches.json)")
        for r in out["widened"]:
            terminalreporter.write_line(f"  {r['branch']}: {r['name']} err {r['err']:.3e} tol {r['tol']:.0e} own {r['own']}")
