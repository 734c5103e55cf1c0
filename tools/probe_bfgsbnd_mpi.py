"""Runs every BFGSBnd_MPI golden case on the device and prints how far the result sits from the verbatim reference's
(tests/golden/bfgsbnd_mpi_golden.npz). Usage on the GPU box: python tools/probe_bfgsbnd_mpi.py > gpurun_out/bfgsbnd_mpi_probe.json"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
from make_bfgsbnd_mpi_golden import PARAMS, cases  # noqa: E402
from parallelnonlinearoptimizationlibrary_b200 import capi, hostapi  # noqa: E402

G = np.load(os.path.join(ROOT, "tests", "golden", "bfgsbnd_mpi_golden.npz"))
ctx = capi.Context(0)
hostapi.attach(ctx)
out = {}
for name, (obj, x0, lb, ub, P, iters, extra, twin) in cases().items():
    k = PARAMS
    p = [k["c1"], k["c2"], k["alphamin"], k["maxalphamult"], k["alphaguess"], k["maxiterls"], k["dxgrad"], k["dxhess"], iters, k["xmindiff"],
         k["mingrad"], k["fsteptol"], extra.get("inithess", 0)]
    try:
        r = hostapi.bfgs("bfgsbnd_mpi", obj, x0, p, lb, ub, pool_width=P)
    except Exception as e:  # noqa: BLE001
        out[name] = {"error": str(e)}
        continue
    X, fOpt = G[name + "/X"], float(G[name + "/fOpt"][0])
    d = {"f0_equal": bool(r["f0"] == G[name + "/f0"][0]), "fOpt": r["fOpt"], "fOpt_ref": fOpt,
         "dX_rel": float(np.linalg.norm(r["X"] - X) / max(np.linalg.norm(X), 1e-300)), "dX_abs": float(np.max(np.abs(r["X"] - X))),
         "iterations": r["iterations"], "iterations_ref": int(G[name + "/iterations_done"])}
    if twin:
        d["twin_dX_rel"] = float(np.linalg.norm(G[name + "/X_ulp"] - X) / np.linalg.norm(X))
        d["twin_dfOpt_rel"] = float(abs(G[name + "/fOpt_ulp"][0] - fOpt) / abs(fOpt))
    out[name] = d
print(json.dumps(out, indent=1))
