"""A/B tool: ms per `run_op(train_op)` step with host float32 sentence tensors (the `e2e` leg of bench.py), for env-var sweeps
(ICL_PACK_NT, ICL_HOST_THREADS):  ICL_PACK_NT=1 python tools/e2e_loop.py"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench, torch
from imagecaptionlearn_py_b200 import core
wl = bench.WORKLOADS["card2048"]
bt = bench.make_batch(wl, 20171201)
if os.environ.get("E2E_F64"):                      # what the reference feeds: np.zeros() float64 sentence tensors
    bt["sentences"] = bt["sentences"].astype("float64")
core.reset_default_graph(); core.set_random_seeds()
with core.variable_scope("bidirectional_lstm"):
    core.setup_bidirectional_lstm(wl["H"], wl["data_norm"], n_embedding_width=300)
core.setup_core_architecture(wl["task"], "first_last_mention", wl["B"], wl["start"], wl["depth"], False, "relu", wl["C"], wl["F"])
core.add_train_op(core.get_collection("loss")[0], 1e-3, 1e-8, 5.0)
sess = core.Session(max_seq_len=50); sess.ensure()
op = core.get_collection("train_op")[0]
res = []
for rep in range(4):
    for _ in range(3):
        core.run_op(sess, op, [bt], 0.5, 0.5, "first_last_mention", [wl["task"]], [""], True)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(20):
        core.run_op(sess, op, [bt], 0.5, 0.5, "first_last_mention", [wl["task"]], [""], True)
    torch.cuda.synchronize(); res.append(1e3 * (time.perf_counter() - t0) / 20)
print("ICL_WIRE_FP16=%s ICL_PACK_NT=%s ICL_HOST_THREADS=%s %s: %s ms/step" % (os.environ.get("ICL_WIRE_FP16", "-"), os.environ.get("ICL_PACK_NT", "-"), os.environ.get("ICL_HOST_THREADS", "-"), bt["sentences"].dtype,
                                                        " ".join("%.3f" % r for r in res)))
