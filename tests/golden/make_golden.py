"""Generates tests/golden/*.npz from the ORACLE (oracle/agp_ref.py).  The reference itself cannot run in
this image (no Julia; it ships no fixtures of its own -- SURVEY.md s8c), so these are regression
fixtures of the oracle's outputs on fixed seeded inputs, not reference-generated vectors
("parity unpinned", see oracle/agp_ref.py header).  Run:  python tests/golden/make_golden.py"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from oracle import agp_ref as ref  # noqa: E402

here = os.path.dirname(os.path.abspath(__file__))


def dump(name, cfg, Xs, extra=None):
    lp = ref.logpdf(cfg["k"], cfg["mean"], cfg["noise"], cfg["X"], cfg["y"])
    post = ref.posterior(cfg["k"], cfg["mean"], cfg["noise"], cfg["X"], cfg["y"])
    m, v = ref.post_mean_and_var(post, Xs)
    out = dict(X=cfg["X"], y=cfg["y"], Xs=Xs, logpdf=lp, alpha=post["alpha"], mean_s=m, var_s=v,
               U_diag=np.diag(post["U"]).copy())
    out.update(extra or {})
    np.savez(os.path.join(here, name), **out)


c1 = ref.make_config("C1")
dump("c1.npz", c1, np.linspace(-0.5, 1.5, 21)[:, None])
c2 = ref.make_config("C2", n=600)
dump("c2_n600.npz", c2, np.random.default_rng(7).random((33, 8)))
c3 = ref.make_config("C3", n=500)
c3["Xs"] = c3["Xs"][:40]
dump("c3_n500_f32.npz", c3, c3["Xs"], dict(ard=c3["k"].ard))
c5 = ref.make_config("C5", n=3000, dtype=np.float64)
el = ref.elbo(c5["k"], c5["mean"], c5["noise"], c5["X"], c5["y"], c5["Z"], c5["jitter"])
dt = ref.dtc(c5["k"], c5["mean"], c5["noise"], c5["X"], c5["y"], c5["Z"], c5["jitter"])
np.savez(os.path.join(here, "c5_n3000_f64.npz"), X=c5["X"], y=c5["y"], Z=c5["Z"], elbo=el, dtc=dt)
print("golden fixtures written")
