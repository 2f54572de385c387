"""Bring-up: per-step timeline of one CTA of k_bptt_nsplit (ICL_BPTT_NSPLIT=0: k_bptt_cluster) on the card2048 bench batch (clock64 events of cell warp 0)."""
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
for i in range(3):
    _cabi.check(L.icl_run_resident(sess.handle, _cabi.OP_TRAIN, 0.5, 0.5, 5 + i))
cta = int(sys.argv[1]) if len(sys.argv) > 1 else 0
os.environ["ICL_PERSISTENT"] = "1"
_cabi.check(L.icl_rec_trace(sess.handle, cta, None))
_cabi.check(L.icl_run_resident(sess.handle, _cabi.OP_TRAIN, 0.5, 0.5, 99))
buf = np.zeros((4, 2048, 4), np.int64)
_cabi.check(L.icl_rec_trace(sess.handle, cta, _cabi.np_ptr(buf)))
ph = (C.c_float * 8)(); L.icl_phase_ms(sess.handle, ph); print("phases", list(ph))
names = ["producer", "mma", "epilogue(w2)"]
ev = buf[2][buf[2][:, 0] >= 0]
t0 = ev[0, 3]
labels = ({0: "step start", 1: "tmem_full", 2: "staged", 3: "bounds done", 4: "cells done", 5: "cluster barrier"} if os.environ.get("ICL_BPTT_NSPLIT", "1") != "0"
          else {0: "step start", 1: "tmem_full", 2: "parked", 3: "S1", 4: "final done", 5: "S23"})
last = t0
for e, k, t, c in ev[:int(os.environ.get("NEV", "80"))]:
    print("  k=%2d %-11s %8.2f us  (+%.2f)" % (k, labels.get(int(e), str(e)), (c - t0) / 1965.0, (c - last) / 1965.0))
    last = c
print("total %.2f us over %d events" % ((ev[-1, 3] - t0) / 1965.0, len(ev)))
