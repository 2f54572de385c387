"""torchrun --nproc-per-node N tools/check_nvls.py : the in-switch all-reduce (icl_nvls_allreduce) against NCCL on the same gradients."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
import bench
from imagecaptionlearn_py_b200 import _cabi, core
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
wl = bench.WORKLOADS[os.environ.get("WL", "nonvis512")]
bts = bench.make_batches(wl, 20171201 + 1000 * rank)
bench.build_graph(wl)
sess = core.Session(max_seq_len=bench.T_PAD, device=local, dist=True)
sess.ensure()
L = _cabi.lib()
dist.broadcast(sess.param_tensor(), 0)
ka = []
b = sess.build_batch(bts, True, ka); sess._bind_stream()
_cabi.check(L.icl_upload(sess.handle, C.byref(b)))
worst = 0.0
for it in range(4):
    _cabi.check(L.icl_run_resident(sess.handle, _cabi.OP_GRADS, 0.5, 0.5, 100 + it))
    if it == 0:
        sess.allreduce_grads()            # first call sets the path up (and adopts the symmetric buffer): redo the step on it
        _cabi.check(L.icl_run_resident(sess.handle, _cabi.OP_GRADS, 0.5, 0.5, 100 + it))
    ref = sess.grad_tensor().clone()
    dist.all_reduce(ref)
    sess.allreduce_grads()
    got = sess.grad_tensor()
    torch.cuda.synchronize()
    err = float((got - ref).abs().max() / ref.abs().max())
    worst = max(worst, err)
    _cabi.check(L.icl_apply_update(sess.handle))
if rank == 0:
    print("nvls path:", sess._nvls is not None, "| max |nvls - nccl| / max |nccl| over 4 steps: %.3g" % worst, "| world", world)
assert worst < 1e-6
dist.barrier(); dist.destroy_process_group()
