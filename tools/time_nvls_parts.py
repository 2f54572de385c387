"""torchrun --nproc-per-node N tools/time_nvls_parts.py : CUDA-event time of the pieces of the in-switch all-reduce."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
import bench
from imagecaptionlearn_py_b200 import _cabi, core
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
wl = bench.WORKLOADS["card2048"]
bts = bench.make_batches(wl, 20171201)
bench.build_graph(wl)
sess = core.Session(max_seq_len=bench.T_PAD, device=local, dist=True)
sess.ensure()
L = _cabi.lib()
ka = []
b = sess.build_batch(bts, True, ka); sess._bind_stream()
_cabi.check(L.icl_upload(sess.handle, C.byref(b)))
_cabi.check(L.icl_run_resident(sess.handle, _cabi.OP_GRADS, 0.5, 0.5, 1))
sess.allreduce_grads()
t, hdl = sess._nvls
mc = C.c_void_p(int(hdl.multicast_ptr))
def timed(fn, n=50):
    for _ in range(5): fn()
    dist.barrier(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return 1e3 * e0.elapsed_time(e1) / n
res = dict(
    barrier=timed(lambda: hdl.barrier(channel=0)),
    kernel=timed(lambda: _cabi.check(L.icl_nvls_allreduce(sess.handle, mc, rank, world))),
    all3=timed(lambda: (hdl.barrier(channel=0), _cabi.check(L.icl_nvls_allreduce(sess.handle, mc, rank, world)), hdl.barrier(channel=1))),
    nccl=timed(lambda: dist.all_reduce(t)),
    update=timed(lambda: _cabi.check(L.icl_apply_update(sess.handle))))
if rank == 0:
    print("world %d, us per call:" % world, {k: round(v, 1) for k, v in res.items()}, "signal pad bytes", hdl.signal_pad_size)
dist.barrier(); dist.destroy_process_group()
