"""rec_fwd phase time of card2048 in predict mode (no gate / c stores) vs training mode, per K2 slicing."""
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
for op, name in ((_cabi.OP_PREDICT, "predict"), (_cabi.OP_TRAIN, "train")):
    acc = np.zeros(8)
    for i in range(25):
        _cabi.check(L.icl_run_resident(sess.handle, op, 0.5, 0.5, 5 + i))
        ph = (C.c_float * 8)(); L.icl_phase_ms(sess.handle, ph)
        if i >= 5: acc += np.array(list(ph))
    print(name, "ICL_RF_U=%s" % os.environ.get("ICL_RF_U", "-"), [round(x / 20, 4) for x in acc])
