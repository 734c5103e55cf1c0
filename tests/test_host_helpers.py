"""CPU: the small host helpers whose semantics this repo defines twice -- oracle/shim (what the verbatim reference is compiled against,
i.e. what every golden vector was made with) and include/pnol/UtilityFunctions.hpp (what the product's host classes call) -- agree:
sequential sums, linspace, first-extremum tie rule, mod / sign bit for bit; the inverse (LU there, Gauss-Jordan here) to 1e-12.
The reference takes them from an un-vendored library ("parity unpinned" at that boundary), so a silent difference
between the two definitions would move the optimisers' iterates without any kernel being wrong."""
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_shim_and_product_host_helpers_agree(tmp_path):
    cxx = shutil.which("g++")
    if cxx is None:
        pytest.skip("no g++ on this box")
    exe = str(tmp_path / "check_host_helpers")
    subprocess.check_call([cxx, "-std=c++17", "-O2", "-ffp-contract=off", "-w", "-I" + os.path.join(ROOT, "oracle", "shim"),
                           "-I" + os.path.join(ROOT, "include"), os.path.join(ROOT, "oracle", "check_host_helpers.cpp"),
                           os.path.join(ROOT, "oracle", "shim", "shim_impl.cpp"), "-o", exe, "-lpthread"])
    r = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stdout[-2000:]
    assert "host helpers agree" in r.stdout
