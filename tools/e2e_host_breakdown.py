"""Where the end-to-end step's host time goes (card2048 by default): Python batch marshalling (Session.build_batch), icl_upload (pack
into pinned memory + H2D enqueue), icl_run_resident (kernel enqueue), icl_poll_stats, against the device time of the step.
    python tools/e2e_host_breakdown.py [workload]"""
import ctypes as C, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")
import numpy as np
import bench, torch
from imagecaptionlearn_py_b200 import _cabi, core

name = sys.argv[1] if len(sys.argv) > 1 else "card2048"
wl = bench.WORKLOADS[name]
rot = bench.make_rotation(wl, 20171201, 4)
train_op = bench.build_graph(wl)
sess = core.Session(max_seq_len=bench.T_PAD); sess.ensure()
L = _cabi.lib()
prev = (_cabi.HeadOut * _cabi.MAX_HEADS)()
sess._bind_stream()
acc = np.zeros(4)
N = 40
for it in range(N + 5):
    ka = []
    t0 = time.perf_counter(); b = sess.build_batch(rot[it % 4], True, ka); t1 = time.perf_counter()
    _cabi.check(L.icl_upload(sess.handle, C.byref(b))); t2 = time.perf_counter()
    _cabi.check(L.icl_run_resident(sess.handle, _cabi.OP_TRAIN, 0.5, 0.5, 5 + it)); t3 = time.perf_counter()
    _cabi.check(L.icl_poll_stats(sess.handle, prev)); t4 = time.perf_counter()
    if it >= 5:
        acc += [t1 - t0, t2 - t1, t3 - t2, t4 - t3]
torch.cuda.synchronize()
print("%s host ms per step: build_batch %.3f | icl_upload %.3f | icl_run_resident (enqueue) %.3f | icl_poll_stats %.3f | sum %.3f"
      % ((name,) + tuple(1e3 * acc / N) + (1e3 * acc.sum() / N,)))
for rep in range(2):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for it in range(N):
        bench.run_train(core, sess, train_op, wl, rot[it % 4])
    torch.cuda.synchronize()
    print("run_op loop: %.3f ms per step" % (1e3 * (time.perf_counter() - t0) / N))
# device only
ka2 = []
b = sess.build_batch(rot[0], True, ka2); _cabi.check(L.icl_upload(sess.handle, C.byref(b))); torch.cuda.synchronize()
t0 = time.perf_counter()
for it in range(N):
    _cabi.check(L.icl_run_resident(sess.handle, _cabi.OP_TRAIN, 0.5, 0.5, 99 + it))
torch.cuda.synchronize()
print("resident steps back to back: %.3f ms per step (host enqueue + device)" % (1e3 * (time.perf_counter() - t0) / N))
