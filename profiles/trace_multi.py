"""Phase times of one epx_multi_analyze call per device thread (EPX_TRACE=1): python profiles/trace_multi.py [n_gpus]"""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["EPX_TRACE"] = "1"
from epimethex_b200 import _native, synth
W = int(sys.argv[1]) if len(sys.argv) > 1 else 2
G, n, P = 20531, 473, 485577
ann = synth.annotation(P, G)
x = synth.expression(G, n)
xp = _native.pinned_empty(x.shape); xp[...] = x
y = _native.pinned_empty((P, n)); synth.methylation(P, n, ann, out=y)
idx = synth.probe_index(ann, 0, G)
ms = _native.MultiSession(list(range(W)))
ms.set_methylation(y)
for i in range(3):
    t0 = time.perf_counter()
    ms.analyze(xp, idx.row_ptr, idx.probe_idx, _native.make_options(), full=False, pinned_results=True)
    print("call", i, round((time.perf_counter() - t0) * 1e3, 3), "ms", file=sys.stderr)
ms.close()
