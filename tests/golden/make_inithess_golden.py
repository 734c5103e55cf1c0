"""Generates tests/golden/inithess_golden.npz from the VERBATIM reference (oracle/_ref/pnol_ref_cli): BFGS-family runs with
initHessFD = true started where the forward-difference Hessian is INDEFINITE. The reference inverts it with matrixInverse (LU)
and carries on with an indefinite D (Source/BFGS_with_linesearch.cpp:35-41, Source/BFGS_bnd_linesearch_MPI_SW.cpp:51-59); a
Cholesky-only inverse would stop there.

    python tests/golden/make_inithess_golden.py        # build container (needs oracle/_ref)
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, ROOT)
import oracle_lib as O  # noqa: E402

BF = dict(c1=1e-4, c2=0.9, dalpha=1e-6, alphaguess=1.0, maxiterls=1000, dxgrad=1e-7, dxhess=1e-3, xmindiff=1e-5, mingrad=1e-5)
SW = dict(c1=1e-4, c2=0.8, dalpha=1e-6, alphaguess=1.0, alphatol=1e-10, alphamult=2.0, maxiterls=50, bndtol=1e-5, dxgrad=1e-6, dxhess=1e-3,
          xmindiff=1e-5, mingrad=1e-5)
# Rosenbrock: d2f/dx_k^2 = 1200 x_k^2 - 400 x_{k+1} + 2 < 0 where x_{k+1} > 3 x_k^2
X0 = np.array([0.1, 1.0, 0.2, 1.2, 0.15])
CASES = {
    "bfgs_it1": ("bfgs", X0, 1, BF, {}),
    "bfgs_it3": ("bfgs", X0, 3, BF, {}),
    "bfgs_to_stop": ("bfgs", X0, 200, BF, {}),
    "bfgs_bnd_sw_it2": ("bfgs_bnd_sw", X0, 2, SW, dict(xlb=np.full(5, -2.0), xub=np.full(5, 2.0), nprocs=4)),
    "goldstein_it2": ("bfgs", np.array([1.2, 0.4]), 2, BF, {}),
}


def main():
    assert O.have_ref()
    G = {}
    for name, (variant, x0, iters, prm, extra) in CASES.items():
        obj = "goldstein" if name.startswith("goldstein") else "rosenbrock"
        arrays = dict(x=x0)
        nprocs = 1
        for k, v in extra.items():
            if k == "nprocs":
                nprocs = v
            else:
                arrays[k] = v
        n = x0.size
        B = O.ref_cli("hessian", arrays=dict(x=x0, dx=np.full(n, prm["dxhess"])), obj=obj)["B"].reshape(n, n)
        ev = np.linalg.eigvalsh(0.5 * (B + B.T))
        assert ev.min() < 0 < ev.max(), "the start point must have an indefinite FD Hessian"
        r = O.ref_cli(variant, arrays=arrays, obj=obj, maxiter=iters, inithess=1, nprocs=nprocs, **prm)
        # the reference's own sensitivity: every start coordinate moved by one ulp, up and down (one coordinate alone can sit in a
        # flat direction: coordinate 0 of X0 moves the iterates by 5e-17, coordinate 1 by 3e-8)
        twins = []
        for j in range(n):
            for d in (1.0, -1.0):
                xt = x0.copy()
                xt[j] = np.nextafter(xt[j], xt[j] + d)
                twins.append(O.ref_cli(variant, arrays=dict(arrays, x=xt), obj=obj, maxiter=iters, inithess=1, nprocs=nprocs, **prm)["X"])
        sens = np.array([np.linalg.norm(t - r["X"]) / np.linalg.norm(r["X"]) for t in twins])
        G[name + "/x0"], G[name + "/X"], G[name + "/f0"], G[name + "/fOpt"], G[name + "/B"] = x0, r["X"], r["f0"], r["fOpt"], B
        G[name + "/X_ulp"] = twins[int(np.argmax(sens))]
        G[name + "/sens_all"] = sens
        print(name, "eig(B) in [%.3g, %.3g]" % (ev.min(), ev.max()), "fOpt", r["fOpt"], "X", r["X"], "one-ulp twins move X by", sens)
    out = os.path.join(HERE, "inithess_golden.npz")
    np.savez_compressed(out, **G)
    print("wrote", out)


if __name__ == "__main__":
    main()
