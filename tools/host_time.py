"""Bring-up: host issue time vs device time of one resident train step; e2e breakdown (upload / run / fetch)."""
import os
os.environ.setdefault('ICL_PHASE_EVENTS', '1')
import ctypes as C, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench, torch
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
outs = (_cabi.HeadOut * _cabi.MAX_HEADS)()
pr = np.empty((wl["B"], wl["C"]), np.float32); pd = np.empty((wl["B"],), np.int64)
outs[0].proba = pr.ctypes.data_as(C.POINTER(C.c_float)); outs[0].pred = pd.ctypes.data_as(C.POINTER(C.c_int64))
for it in range(6):
    torch.cuda.synchronize()
    t0 = time.perf_counter(); _cabi.check(L.icl_upload(sess.handle, C.byref(b))); t1 = time.perf_counter()
    torch.cuda.synchronize(); t1b = time.perf_counter()
    _cabi.check(L.icl_run_resident(sess.handle, _cabi.OP_TRAIN, 0.5, 0.5, 5 + it)); t2 = time.perf_counter()
    torch.cuda.synchronize(); t3 = time.perf_counter()
    _cabi.check(L.icl_fetch(sess.handle, outs)); t4 = time.perf_counter()
    ms = C.c_float(); L.icl_last_step_ms(sess.handle, C.byref(ms))
    print("upload host %.2f ms (+%.2f to finish copies) | run: host issue %.2f ms, until done %.2f ms, device %.2f ms | fetch %.2f ms"
          % (1e3 * (t1 - t0), 1e3 * (t1b - t1), 1e3 * (t2 - t1b), 1e3 * (t3 - t1b), ms.value, 1e3 * (t4 - t3)))
t0 = time.perf_counter()
for i in range(5):
    kal = []
    bb = sess.build_batch([bt], True, kal)
t1 = time.perf_counter()
print("python build_batch %.2f ms" % (1e3 * (t1 - t0) / 5))
