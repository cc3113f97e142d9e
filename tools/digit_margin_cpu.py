"""CPU study (NumPy emulation of the int8 digit split, tools/proto_ozaki.py) of the margin of the default digit counts:
Cholesky part 8 digits, inverse part + K^-1 7 digits. Gradient error of one evaluation against the oracle for noise levels
from 1e-1 to 1e-5 (condition number of Ky grows as variance * N / noise) and inverse-part digits 8 / 7 / 6 / 5.
The tolerance is 1e-6 relative on gradients, 1e-8 absolute on the log marginal likelihood.
    python tools/digit_margin_cpu.py [N] [NB] [D] [lengthscale factor]     (lengthscale = factor * sqrt(D); D = 2 with factor 1.5
                                                                          gives smooth, badly conditioned covariance matrices)"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import gpy_oracle as o  # noqa: E402
from proto_ozaki import evaluate, ozaki_nt  # noqa: E402


def main():
    N = int(sys.argv[1]) if len(sys.argv) > 1 else 768
    NB = int(sys.argv[2]) if len(sys.argv) > 2 else 128
    D = int(sys.argv[3]) if len(sys.argv) > 3 else 8
    fac = float(sys.argv[4]) if len(sys.argv) > 4 else 1.0
    X, Y = o.synthetic(N, D)
    var, ls, _ = o.theta_bench(D, True)
    ls = ls * fac
    print("N=%d NB=%d D=%d RBF ARD, variance %.1f, lengthscale %g * sqrt(D); entries: |dLML| / max relative gradient error" % (N, NB, D, var, fac))
    print("%-9s %-10s %-22s %-22s %-22s %-22s %-22s" % ("noise", "cond(Ky)", "fp64 blocked sweep", "8 / 8 digits", "8 / 7 (default)",
                                                       "8 / 6", "8 / 5"))
    for noise in (1e-1, 1e-2, 1e-3, 1e-4, 1e-5):
        lml0, g0, res = o.eval_lml_grad(X, Y, "rbf", True, var, ls, noise)
        ev = np.linalg.eigvalsh(res["K"] + (noise + 1e-8) * np.eye(N))
        cells = []
        lml, g = evaluate(X, Y, var, ls, noise, NB, lambda A, B: A @ B.T)
        cells.append("%.1e / %.1e" % (abs(lml - lml0), np.max(np.abs(g - g0) / np.abs(g0))))
        for SU in (8, 7, 6, 5):
            lml, g = evaluate(X, Y, var, ls, noise, NB, lambda A, B: ozaki_nt(A, B, 8), lambda A, B, SU=SU: ozaki_nt(A, B, SU))
            cells.append("%.1e / %.1e" % (abs(lml - lml0), np.max(np.abs(g - g0) / np.abs(g0))))
        print("%-9.0e %-10.1e %-22s %-22s %-22s %-22s %-22s" % ((noise, ev[-1] / ev[0]) + tuple(cells)), flush=True)


if __name__ == "__main__":
    main()
