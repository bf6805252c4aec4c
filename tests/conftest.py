import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)

GOLD = os.path.join(ROOT, "tests", "golden")


def _build_if_missing():
    """A fresh checkout has no built libraries (they are git-ignored): build them once, as __graft_entry__.build() does
    (nvcc cross-compiles sm_100a without a GPU; ~20 s)."""
    import subprocess
    lib = os.path.join(ROOT, "ft8_demodulator_b200", "libft8b200.so")
    orc = os.path.join(ROOT, "oracle", "libft8oracle.so")
    if not os.path.exists(lib):
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "ft8_demodulator_b200", "csrc"), "-j4"], stdout=subprocess.DEVNULL)
    if not os.path.exists(orc):
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle")], stdout=subprocess.DEVNULL)


def pytest_configure(config):
    _build_if_missing()
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")
    config.addinivalue_line("markers", "needs_reference: needs the reference tree at /root/reference (build container only)")


def pytest_collection_modifyitems(config, items):
    import reference_loader
    have_ref = reference_loader.available()
    skip_ref = pytest.mark.skip(reason="/root/reference not present (it does not travel to the GPU box)")
    for item in items:
        if "needs_reference" in item.keywords and not have_ref:
            item.add_marker(skip_ref)


@pytest.fixture(scope="session")
def orc():
    import ft8_oracle
    ft8_oracle.lib()
    return ft8_oracle


def load_golden(name):
    path = os.path.join(GOLD, name)
    if name.endswith(".json"):
        with open(path) as f:
            return json.load(f)
    return np.load(path, allow_pickle=False)


def golden_case(npz, tag):
    """Sub-dictionary of one traced decode ('<tag>__key' entries); plain bits are unpacked."""
    pre = tag + "__"
    d = {k[len(pre):]: npz[k] for k in npz.files if k.startswith(pre)}
    if "plain_packed" in d:
        d["plain"] = np.unpackbits(d["plain_packed"], axis=1)[:, :174]
    return d


@pytest.fixture(scope="session")
def c1_slot():
    from ft8_demodulator_b200 import synth
    return synth.c1_slot()
