"""Repeats the SAME training-mode forward+backward (same seed) and compares the LSTM outputs run to run: any difference is a race.
   python tools/debug_fwd_race.py <workload> [runs]"""
import ctypes as C, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from imagecaptionlearn_py_b200 import _cabi, core
name = sys.argv[1] if len(sys.argv) > 1 else "card2048"
runs = int(sys.argv[2]) if len(sys.argv) > 2 else 12
wl = bench.WORKLOADS[name]
bts = bench.make_batches(wl, 20171201)
bench.build_graph(wl)
sess = core.Session(max_seq_len=bench.T_PAD)
sess.ensure()
L = _cabi.lib()
S = sum(len(bt["seq_lengths"]) for bt in bts)
lens = np.concatenate([bt["seq_lengths"] for bt in bts])
first = None
for r in range(runs):
    sess.base_seed, sess.run_counter = 9, 0
    sess.run(_cabi.OP_GRADS, [dict(bt) for bt in bts], 0.5, 0.5, True)
    outs = []
    for d in range(2):
        out = np.empty((S, sess.max_seq_len, wl["H"]), np.float32)
        _cabi.check(L.icl_get_lstm_outputs(sess.handle, d, _cabi.np_ptr(out)))
        outs.append(out)
    if first is None:
        first = outs
        print(name, "S", S, "H", wl["H"], "finite", all(np.isfinite(o).all() for o in outs))
        continue
    for d in range(2):
        diff = np.abs(outs[d] - first[d])
        bad = np.argwhere(diff > 0)
        if len(bad):
            order = np.argsort(-lens, kind="stable"); rank = np.empty(S, int); rank[order] = np.arange(S)
            print("run %d dir %d: %d elements differ (max %.3g); seq ranks %s; t %s; units %s" % (
                r, d, len(bad), diff.max(), sorted(set(rank[bad[:, 0]] // 128))[:12], sorted(set(bad[:, 1]))[:12], sorted(set(bad[:, 2] // 4 * 4))[:16]))
print("done")
