"""pytest wiring: markers, import paths, fixtures shared by the CPU and GPU suites."""
import glob
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "genetic-gaussian-splats_b200")
for p in (ROOT, PKG):
    if p not in sys.path:
        sys.path.insert(0, p)

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


# Counts the parity tests want in the driver's log whatever the verbosity: AABB edges flipped by
# a last-ulp difference, candidates whose result is undefined in the reference itself, ...
PARITY_NOTES = []


def note(msg: str) -> None:
    PARITY_NOTES.append(str(msg))


def pytest_terminal_summary(terminalreporter):
    if PARITY_NOTES:
        terminalreporter.section("parity notes")
        for line in PARITY_NOTES:
            terminalreporter.write_line(line)


def golden_names():
    # render / fitness cases (make_golden.py); mask_cases.npz (make_mask_golden.py) has its own tests
    return sorted(os.path.splitext(os.path.basename(p))[0]
                  for p in glob.glob(os.path.join(GOLDEN_DIR, "*.npz"))
                  if os.path.basename(p) not in ("mask_cases.npz", "breed_reference_stats.npz"))


def load_golden(name):
    with np.load(os.path.join(GOLDEN_DIR, name + ".npz")) as z:
        return {k: z[k] for k in z.files}


@pytest.fixture(params=golden_names())
def golden(request):
    g = load_golden(request.param)
    g["name"] = request.param
    return g
