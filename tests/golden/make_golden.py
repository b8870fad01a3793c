"""Generates tests/golden/oracle_small.json from the CPU oracle (the reference is R and cannot run
here, so the fixture pins the oracle against drift and gives the GPU tests a fixed case).
Run from the repo root:  python tests/golden/make_golden.py"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle  # noqa: E402
from helpers import design, markers  # noqa: E402

L, E, p, ite, burn, thin = 12, 3, 25, 30, 6, 2
X = markers(L, p, 123)
Y, names = design(L, E)
K = oracle.getK(Y, X, kernel="GK", model="MDs", rownames=names)
ne = [L] * E
eig = oracle.setDEC(K, ne, 1e-10)
for blocks in eig:                       # canonical sign (largest-|.| entry positive), as the device does
    for b in blocks:
        sg = np.sign(b["U"][np.argmax(np.abs(b["U"]), axis=0), np.arange(b["U"].shape[1])])
        b["U"] = b["U"] * sg
rng = np.random.default_rng(7)
y = rng.standard_normal(L * E)
y[[4, 17]] = np.nan
nrs = [sum(b["s"].shape[0] for b in bl) for bl in eig]
nz, nc = oracle.tape_sizes(L * E, nrs, ite, n_na=2)
z = rng.standard_normal(nz)
c = rng.chisquare(np.tile(np.array([nr + 3 for nr in nrs] + [L * E + 3], dtype=float), ite))
fit = oracle.BGGE(y, K, ne=ne, ite=ite, burn=burn, thin=thin, draws=oracle.TapeDraws(z, c), eig=eig)
out = {
    "X": X.tolist(), "Y": [list(r) for r in Y], "names": names, "ne": ne, "ite": ite, "burn": burn, "thin": thin,
    "y": [None if np.isnan(v) else v for v in y], "z": z.tolist(), "c": c.tolist(),
    "G": K["G"]["Kernel"].tolist(), "GE": K["GE"]["Kernel"].tolist(),
    "eig": [[{"s": b["s"].tolist(), "U": b["U"].tolist(), "pos": b["pos"]} for b in bl] for bl in eig],
    "yHat": fit["yHat"].tolist(), "varE": fit["varE"], "chain_mu": fit["chain"]["mu"].tolist(),
    "chain_varE": fit["chain"]["varE"].tolist(), "chain_varU_G": fit["chain"]["K"]["G"]["varU"].tolist(),
    "chain_varU_GE": fit["chain"]["K"]["GE"]["varU"].tolist(), "u_G": fit["K"]["G"]["u"].tolist(),
    "u_sd_GE": fit["K"]["GE"]["u.sd"].tolist(), "y_out": fit["y"].tolist(),
}
with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "oracle_small.json"), "w") as f:
    json.dump(out, f)
print("wrote oracle_small.json", os.path.getsize(os.path.join(os.path.dirname(os.path.abspath(__file__)), "oracle_small.json")))
