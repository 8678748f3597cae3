"""Generates tests/golden/*.npz from the CPU oracle (the reference itself cannot be built or imported here:
Eigen / Boost / fmt / Pinocchio are absent, DESIGN.md "Oracle").  The fixtures therefore pin the oracle against
regressions and give the GPU tests committed vectors; they do NOT pin the oracle to the reference ("parity
unpinned" for the Pinocchio arithmetic).  Run:  python tests/golden/make_golden.py"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE)); sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import ddpb_loader  # noqa: E402

P = ddpb_loader.load()
import oracle_lib as ol  # noqa: E402

S = P.synthetic
CASES = {"c1": dict(T=24, iters=4), "c2": dict(T=12, iters=3), "c3": dict(T=8, iters=3), "c4": dict(T=8, iters=2), "c5": dict(T=6, iters=2)}


def main():
    for name, kw in CASES.items():
        w = S.workload(name, B=2, T=kw["T"])
        g = w.generate()
        out = dict(x0=g["x0"], us0=g["us"], targets=g["targets"], mult_jac=g["mult_jac"], ne=g["ne"], iters=kw["iters"], mu=w.mu)
        for b in range(w.B):
            o = ol.from_workload(w, g["targets"][b])
            xs0 = o.rollout(g["x0"][b], g["us"][b])
            d = o.compute_derivatives(xs0, g["us"][b])
            mv = np.zeros(o.ne_total)
            k, K, mu1, reg1, rs = o.backward_pass(d, xs0, mv, g["mult_jac"][b], 0.0, w.mu)
            r = o.solve(xs0, g["us"][b], g["mult_jac"][b], kw["iters"], w.mu)
            rq = o.solve(xs0, g["us"][b], g["mult_jac"][b], kw["iters"], w.mu, prec=1)   # __float128 oracle (accuracy reference)
            out[f"xsq_{b}"] = rq["xs"]
            out.update({f"xs0_{b}": xs0, f"k_{b}": k, f"K_{b}": K, f"mu1_{b}": mu1, f"reg1_{b}": reg1, f"xs_{b}": r["xs"], f"us_{b}": r["us"],
                        f"log_{b}": r["log"]})
            for f in ("f_val", "fx", "fu", "eq_val", "eq_x", "eq_u", "fxx", "fux", "eq_xx"):
                out[f"{f}_{b}"] = d[f]
        np.savez_compressed(os.path.join(HERE, f"{name}.npz"), **out)
        print("wrote", name)


if __name__ == "__main__":
    main()
