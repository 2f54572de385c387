"""Per-CTA summary of the k_rec_fwd16 clock64 trace at one step: who waits for whom.
   python tools/trace_fwd16_summary.py <step> <cta> [<cta> ...]"""
import ctypes as C, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ.setdefault("ICL_TRACE_CS", "0")
import bench
from imagecaptionlearn_py_b200 import _cabi, core
wl = bench.WORKLOADS[os.environ.get("WL", "card2048")]
bt = bench.make_batch(wl, 20171201)
core.reset_default_graph(); core.set_random_seeds()
with core.variable_scope("bidirectional_lstm"):
    core.setup_bidirectional_lstm(wl["H"], wl["data_norm"], n_embedding_width=300)
core.setup_core_architecture(wl["task"], "first_last_mention", wl["B"], wl["start"], wl["depth"], False, "relu", wl["C"], wl["F"])
core.add_train_op(core.get_collection("loss")[0], 1e-3, 1e-8, 5.0)
sess = core.Session(max_seq_len=50); sess.ensure()
L = _cabi.lib()
ka = []
b = sess.build_batch([bt], True, ka); sess._bind_stream()
_cabi.check(L.icl_upload(sess.handle, C.byref(b)))
for i in range(3):
    _cabi.check(L.icl_run_resident(sess.handle, _cabi.OP_TRAIN, 0.5, 0.5, 5 + i))
K = int(sys.argv[1])
print("step %d; per tile of the CTA: producer flag wait / load issue, mma first k-block -> commit, cell warp accumulator wait / cells, store warp boxes ready -> published" % K)
for cta in [int(x) for x in sys.argv[2:]]:
    _cabi.check(L.icl_rec_trace(sess.handle, cta, None))
    _cabi.check(L.icl_run_resident(sess.handle, _cabi.OP_TRAIN, 0.5, 0.5, 99))
    buf = np.zeros((4, 2048, 4), np.int64)
    _cabi.check(L.icl_rec_trace(sess.handle, cta, _cabi.np_ptr(buf)))
    def evs(r):
        e = buf[r][buf[r][:, 0] >= 0]
        return e[e[:, 1] == K]
    us = lambda c: c / 1965.0
    def spans(r, a, b_):
        e = evs(r); out = []
        for t in sorted(set(e[:, 2])):
            x = e[e[:, 2] == t]
            ta, tb = x[x[:, 0] == a][:, 3], x[x[:, 0] == b_][:, 3]
            if len(ta) and len(tb): out.append(us(tb[0] - ta[0]))
        return out
    p = evs(0)
    period = us(p[:, 3].max() - p[:, 3].min()) / max(1, len(set(p[:, 2]))) if len(p) else float("nan")
    f = lambda v: " ".join("%.1f" % x for x in v)
    print("cta %3d: ~%.2f us/tile | flag wait [%s] issue [%s] | mma [%s] | acc wait [%s] cells [%s] | publish [%s]" % (
        cta, period, f(spans(0, 0, 1)), f(spans(0, 1, 2)), f(spans(1, 1, 2)), f(spans(3, 1, 2)), f(spans(3, 2, 4)), f(spans(2, 0, 1))))
