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


def test_three_operation_quotient_claim_in_exact_arithmetic():
    """include/pnol/device/exact_div.cuh, div_exact3_core: for a divisor d whose rounded reciprocal r = RN(1/d) satisfies
    |r d - 1| <= (15/32) 2^-53, q1 = RN(q0 + RN(x - q0 d) r) with q0 = RN(x r) is the correctly rounded x / d. Checked here with exact
    rationals (float(Fraction) rounds to nearest even, as the device's RN operations do; the residual x - q0 d is exact in an FMA):
    random and adversarial significands of d and x, and dividends placed next to a floating-point quotient or a midpoint. The GPU side of
    the same claim is pnol_selftest_fast_div (4e8 pairs) and the bit-exact Jacobian tests."""
    import math
    import random
    import struct
    from fractions import Fraction as Fr

    rnd = random.Random(20261019)

    def from_bits(b):
        return struct.unpack("<d", struct.pack("<Q", b))[0]

    def draw(e_lo, e_hi, mode):
        mant = rnd.getrandbits(52)
        if mode == 1:
            mant |= 0xFFFFFFFFFF000
        elif mode == 2:
            mant &= 0xFFF
        elif mode == 3:
            mant = 0xFFFFFFFFFFFFE - rnd.getrandbits(4)
        v = from_bits(((rnd.randint(e_lo, e_hi) + 1023) << 52) | mant)
        return v if rnd.random() < 0.5 else -v

    def three(x, d, r):
        q0 = float(Fr(x) * Fr(r))
        r0 = float(Fr(x) - Fr(q0) * Fr(d))
        return float(Fr(q0) + Fr(r0) * Fr(r))

    checked = qualifying = 0
    for i in range(12000):
        d = draw(-40, 20, i % 4)
        r = float(1 / Fr(d))
        if abs(Fr(r) * Fr(d) - 1) * 2 ** 53 > Fr(15, 32):
            continue
        qualifying += 1
        xs = [draw(-40, 40, (i // 4) % 4)]
        q = abs(draw(-20, 20, (i // 16) % 3))
        ulp = math.nextafter(q, math.inf) - q
        x_mid = float(Fr(d) * (Fr(q) + Fr(rnd.choice([-1, 0, 1, 1, 2]), 2) * Fr(ulp)))      # x / d next to q or to q + ulp / 2
        xs += [x_mid, math.nextafter(x_mid, math.inf), math.nextafter(x_mid, -math.inf)]
        for x in xs:
            assert three(x, d, r) == float(Fr(x) / Fr(d)), (x.hex(), d.hex())
            checked += 1
    assert qualifying > 4000 and checked == 4 * qualifying
    # the steps the examples use: 1e-6 and 1e-7 qualify, 1e-5 does not (and takes the five-operation quotient)
    for d, ok in ((1e-6, True), (1e-7, True), (1e-3, True), (1e-5, False), (1.3e-6, False)):
        assert (abs(Fr(float(1 / Fr(d))) * Fr(d) - 1) * 2 ** 53 <= Fr(15, 32)) == ok
