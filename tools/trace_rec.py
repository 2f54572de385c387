"""Bring-up: run the card2048 bench batch once with the persistent-kernel trace on and print a per-(step,tile) timeline."""
import os
os.environ.setdefault('ICL_PHASE_EVENTS', '1')      # these tools read icl_phase_ms
import ctypes as C, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from imagecaptionlearn_py_b200 import _cabi, core
wl = bench.WORKLOADS["card2048"]
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
op = int(sys.argv[2]) if len(sys.argv) > 2 else _cabi.OP_TRAIN
for i in range(3):
    _cabi.check(L.icl_run_resident(sess.handle, op, 0.5, 0.5, 5 + i))
cta = int(sys.argv[1]) if len(sys.argv) > 1 else 0
_cabi.check(L.icl_rec_trace(sess.handle, cta, None))
_cabi.check(L.icl_run_resident(sess.handle, op, 0.5, 0.5, 99))
buf = np.zeros((4, 2048, 4), np.int64)
_cabi.check(L.icl_rec_trace(sess.handle, cta, _cabi.np_ptr(buf)))
ph = (C.c_float * 8)(); L.icl_phase_ms(sess.handle, ph); print("phases", list(ph))
names = ["producer", "mma", "loader", "epilogue"]
t0 = min(buf[r][buf[r][:, 0] >= 0][:, 3].min() for r in range(4) if (buf[r][:, 0] >= 0).any())
for r in range(4):
    ev = buf[r][buf[r][:, 0] >= 0]
    print("==", names[r], len(ev), "events")
    for e, k, t, c in ev[:int(os.environ.get("NEV", "60"))]:
        print("   ev%d k=%d t=%d  %8.2f us" % (e, k, t, (c - t0) / 1965.0))
    if len(ev):
        print("   last: %.2f us" % ((ev[-1, 3] - t0) / 1965.0))
