"""Bring-up: host timeline of the pipelined train step (upload / enqueue / poll) and raw pinned H2D bandwidth."""
import os
os.environ.setdefault('ICL_PHASE_EVENTS', '1')      # these tools read icl_phase_ms
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
prev = (_cabi.HeadOut * _cabi.MAX_HEADS)()
for it in range(8):
    t0 = time.perf_counter(); _cabi.check(L.icl_upload(sess.handle, C.byref(b))); t1 = time.perf_counter()
    _cabi.check(L.icl_run_resident(sess.handle, _cabi.OP_TRAIN, 0.5, 0.5, 5 + it)); t2 = time.perf_counter()
    _cabi.check(L.icl_poll_stats(sess.handle, prev)); t3 = time.perf_counter()
    print("upload %.2f ms | enqueue %.2f ms | poll(prev) %.2f ms | total %.2f" % (1e3 * (t1 - t0), 1e3 * (t2 - t1), 1e3 * (t3 - t2), 1e3 * (t3 - t0)))
tl = (C.c_float * 8)()
L.icl_debug_timeline.argtypes = [C.c_void_p, C.POINTER(C.c_float)]
_cabi.check(L.icl_debug_timeline(sess.handle, tl))
print("timeline (ms): older copy %.2f-%.2f compute %.2f-%.2f | newer copy %.2f-%.2f compute %.2f-%.2f" % tuple(tl))
torch.cuda.synchronize()
t0 = time.perf_counter()
for it in range(10):
    core.run_op(sess, core.get_collection("train_op")[0], [bt], 0.5, 0.5, "first_last_mention", [wl["task"]], [""], True)
torch.cuda.synchronize()
print("run_op loop: %.2f ms/step" % (1e3 * (time.perf_counter() - t0) / 10))
x = torch.empty(32 << 20, dtype=torch.uint8).pin_memory(); y = torch.empty(32 << 20, dtype=torch.uint8, device="cuda")
for n in (1 << 20, 4 << 20, 32 << 20):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(10):
        y[:n].copy_(x[:n], non_blocking=True)
    torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / 10
    print("pinned H2D %d MiB: %.1f GB/s" % (n >> 20, n / dt / 1e9))
# which phase of a step stretches while the NEXT batch is being packed / copied?
ph = (C.c_float * 8)()
for overlap in (0, 1, 0, 1):
    _cabi.check(L.icl_upload(sess.handle, C.byref(b))); torch.cuda.synchronize()
    _cabi.check(L.icl_run_resident(sess.handle, _cabi.OP_TRAIN, 0.5, 0.5, 77))
    if overlap:
        _cabi.check(L.icl_upload(sess.handle, C.byref(b)))
    _cabi.check(L.icl_phase_ms(sess.handle, ph))
    print("overlap=%d phases %s sum %.2f" % (overlap, " ".join("%s=%.3f" % (n, v) for n, v in zip(_cabi.PHASES, ph)), sum(ph)))
side = torch.cuda.Stream()
big = torch.empty(64 << 20, dtype=torch.uint8).pin_memory(); dbig = torch.empty(64 << 20, dtype=torch.uint8, device="cuda")
import threading
def burn(stop):
    a = np.ones(8 << 20, np.float32); c = np.empty_like(a)
    while not stop.is_set():
        np.copyto(c, a)
for mode in ("none", "h2d-only", "cpu-memcpy-only", "none"):
    _cabi.check(L.icl_upload(sess.handle, C.byref(b))); torch.cuda.synchronize()
    stop = threading.Event(); th = []
    if mode == "cpu-memcpy-only":
        th = [threading.Thread(target=burn, args=(stop,)) for _ in range(8)]
        [t.start() for t in th]; time.sleep(0.05)
    _cabi.check(L.icl_run_resident(sess.handle, _cabi.OP_TRAIN, 0.5, 0.5, 77))
    if mode == "h2d-only":
        with torch.cuda.stream(side):
            dbig.copy_(big, non_blocking=True)
    _cabi.check(L.icl_phase_ms(sess.handle, ph))
    stop.set(); [t.join() for t in th]
    torch.cuda.synchronize()
    print("%s: phases %s sum %.2f" % (mode, " ".join("%s=%.3f" % (n, v) for n, v in zip(_cabi.PHASES, ph)), sum(ph)))
