"""In-situ kernel timeline of ONE evaluation (CUPTI through torch.profiler; there is no nsys in the image):

    python tools/timeline.py [N] [out.txt]

Prints, for the evaluation: every kernel's stream, start and duration, the idle gaps of the main stream (and what the side
stream ran meanwhile), and per-kernel totals with concurrency preserved (unlike an ncu launch list, which serialises)."""
import json
import os
import sys
import tempfile

import numpy as np
import torch
from torch.profiler import ProfilerActivity, profile

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gpy_b200 import _ffi  # noqa: E402


def main():
    N = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
    out = open(sys.argv[2], "w") if len(sys.argv) > 2 else sys.stdout
    opts = {}
    for a in sys.argv[3:]:
        k, v = a.split("=")
        opts[k] = int(v)
    D = 8
    rng = np.random.default_rng(0)
    X = rng.uniform(-3, 3, (N, D))
    Y = np.sin(X).sum(1, keepdims=True) / np.sqrt(D) + 0.1 * rng.standard_normal((N, 1))
    th = ("rbf", True, 1.0, np.full(D, np.sqrt(D)), 0.01)
    torch.cuda.init()
    e = _ffi.Engine(0)
    for k, v in opts.items():
        e.set_option(k, v)
    e.set_data(X, Y)
    for _ in range(3):
        e.exact_eval(*th)
    with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
        e.exact_eval(*th)
        torch.cuda.synchronize()
    st = e.stats()
    path = os.path.join(tempfile.mkdtemp(), "trace.json")
    prof.export_chrome_trace(path)
    ev = [x for x in json.load(open(path))["traceEvents"] if x.get("cat") in ("kernel", "gpu_memcpy", "gpu_memset")]
    ev.sort(key=lambda x: x["ts"])
    t0 = ev[0]["ts"]
    streams = {}
    for x in ev:
        streams.setdefault(x["args"].get("stream"), []).append(x)
    main_stream = max(streams, key=lambda s: sum(x["dur"] for x in streams[s] if "oz_gemm" in x["name"] or "gemm_update" in x["name"]))
    print("N=%d options %s : engine total %.2f ms, trace span %.2f ms, %d device activities, streams %s (main %s)" % (
        N, opts, st["total_ms"], (ev[-1]["ts"] + ev[-1]["dur"] - t0) / 1e3, len(ev), sorted(streams), main_stream), file=out)
    tot = {}
    for x in ev:
        n = x["name"].split("(")[0].replace("void ", "").replace("gpx::", "")
        a = tot.setdefault(n, [0, 0.0])
        a[0] += 1
        a[1] += x["dur"]
    print("\nper-kernel totals inside the evaluation (in situ, overlapping kernels both count):", file=out)
    for n, (c, d) in sorted(tot.items(), key=lambda kv: -kv[1][1]):
        print("  %-44s n=%5d  total %9.3f ms  avg %8.1f us" % (n[:44], c, d / 1e3, d / c), file=out)
    # main-stream gaps
    print("\nmain-stream idle gaps > 30 us (what the other streams ran in the gap):", file=out)
    ms = streams[main_stream]
    gap_total = 0.0
    for a, b in zip(ms[:-1], ms[1:]):
        g0, g1 = a["ts"] + a["dur"], b["ts"]
        if g1 - g0 > 30:
            gap_total += g1 - g0
            inside = {}
            for s, lst in streams.items():
                if s == main_stream:
                    continue
                for x in lst:
                    lo, hi = max(x["ts"], g0), min(x["ts"] + x["dur"], g1)
                    if hi > lo:
                        n = x["name"].split("(")[0].replace("void ", "").replace("gpx::", "")[:24]
                        inside[n] = inside.get(n, 0.0) + (hi - lo)
            print("  at %8.3f ms: %7.1f us before %-22s | %s" % ((g0 - t0) / 1e3, g1 - g0, b["name"].split("(")[0].replace("gpx::", "")[:22],
                  ", ".join("%s %.0f" % kv for kv in sorted(inside.items(), key=lambda kv: -kv[1]))), file=out)
    print("  sum of main-stream gaps: %.2f ms" % (gap_total / 1e3), file=out)
    print("\nfull list (ms from the first activity; stream; duration us; name):", file=out)
    for x in ev:
        print("  %9.3f  s%-4s %9.1f  %s" % ((x["ts"] - t0) / 1e3, x["args"].get("stream"), x["dur"],
                                           x["name"].split("(")[0].replace("gpx::", "")[:60]), file=out)


if __name__ == "__main__":
    main()
