"""Generates tests/golden/examples_ref.json: the tail of what the VERBATIM reference's own example drivers print
(/root/reference/Source/Examples.cpp run through `oracle/_ref/pnol_ref_cli example name=<driver>`, rank 0 of 8 mini-MPI ranks, the
GA / simplex random stream = the shim's counter stream with seed 12345). tests/test_gpu_dropin_examples.py compares the same drivers
compiled against include/pnol and run on the B200 (oracle/_ref/pnol_examples_dropin) with these. Needs /root/reference:

    python tests/golden/make_examples_golden.py"""
import json
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle_lib as O  # noqa: E402

DRIVERS = ["testBFGS", "testBFGS_booth", "testBFGS_MPI", "testBFGSBnd", "testBFGSBndMPISW", "testBFGSBnd_MPI", "testLMExp", "testLMExpMPI",
           "testLMCubicLinearCoef", "testGA", "testGAParallel", "testSimplexSearch", "testHessian", "testCreateObject",
           "testGradientEvaluation", "testGradientApproxMultMPI", "testGradientApproxMultMPIRecur"]
POOL = 8
SEED = 12345
TAIL = 3000


def main():
    assert O.have_ref(), "build oracle/_ref first (make -C oracle ref)"
    out = {"pool_width": POOL, "seed": SEED, "tails": {}}
    for d in DRIVERS:
        env = dict(os.environ, PNOL_SHIM_NPROCS=str(POOL))
        r = subprocess.run([O.REF_CLI, "example", "name=" + d, "seed=%d" % SEED, "out=/tmp/pnol_examples_unused"], env=env, capture_output=True,
                           text=True, timeout=600)
        assert r.returncode == 0, (d, r.stderr[-500:])
        out["tails"][d] = r.stdout[-TAIL:]
        print("%-32s %6d chars" % (d, len(r.stdout)))
    with open(os.path.join(HERE, "examples_ref.json"), "w") as f:
        json.dump(out, f, indent=1)


if __name__ == "__main__":
    main()
