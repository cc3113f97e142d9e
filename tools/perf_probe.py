"""Performance probe (not a test): per-phase device times of the exact-GP evaluation at the benchmark sizes."""
import sys, os, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from gpy_b200 import _ffi
from oracle import gpy_oracle as o

def main():
    sizes = [int(s) for s in sys.argv[1].split(",")] if len(sys.argv) > 1 else [4096, 16384]
    nbs = [int(s) for s in sys.argv[2].split(",")] if len(sys.argv) > 2 else [0]
    reps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
    D = 8
    eng = _ffi.Engine(0)
    for N in sizes:
        X, Y = o.synthetic(N, D)
        var, ls, noise = o.theta_bench(D, True)
        for nb in nbs:
            eng.set_option("nb", nb)
            eng.N = 0
            # force reallocation with the new block size
            eng.set_data(X[: N - 1], Y[: N - 1]) if False else None
            eng.set_data(X, Y)
            for r in range(reps):
                t0 = time.time()
                lml, g, jit = eng.exact_eval("rbf", True, var, ls, noise)
                wall = (time.time() - t0) * 1e3
                st = eng.stats()
            flops = float(N) ** 3
            print("N=%d nb=%d lml=%.9f total %.2f ms (wall %.2f) kbuild %.3f sweep %.2f [update %.2f ms = %.2f TF/s over %d launches] solve %.3f lauum %.2f ms = %.2f TF/s | N^3/total = %.2f TF/s | %.2f evals/s | launches %d" % (
                N, nb, lml, st["total_ms"], wall, st["kbuild_ms"], st["sweep_ms"], st["update_ms"], st["update_flops"] / st["update_ms"] * 1e-9 if st["update_ms"] else 0,
                st["update_launches"], st["solve_ms"], st["lauum_ms"], st["lauum_flops"] / st["lauum_ms"] * 1e-9, flops / st["total_ms"] * 1e-9, 1e3 / st["total_ms"], st["launches"]), flush=True)
            print("   grad", np.array2string(g, precision=10), "kbuild GB/s %.0f" % (st["kbuild_bytes"] / st["kbuild_ms"] * 1e-6), flush=True)

if __name__ == "__main__":
    main()
