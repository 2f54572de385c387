"""Device-timed phases of named bench.py workloads (one GPU): python tools/bench_configs.py multitask512 affinity512 [--steps N]
Prints one line per workload: ms per step, end-to-end ms, launches per step, CUDA-event time of every kernel group."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")


def main():
    import torch
    import bench
    from imagecaptionlearn_py_b200 import _cabi
    steps = int(sys.argv[sys.argv.index("--steps") + 1]) if "--steps" in sys.argv else 20
    names = [a for i, a in enumerate(sys.argv[1:], 1) if not a.startswith("--") and sys.argv[i - 1] != "--steps"]
    ctx = bench.Ctx()
    ctx.rank, ctx.world, ctx.local, ctx.dist = 0, 1, 0, None
    torch.cuda.set_device(0)
    ctx.gemm_mode = _cabi.GEMM_TCGEN05_TF32
    ctx.flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    for name in names or list(bench.BY_CONFIG):
        r = bench.measure(name, steps, 5, ctx)
        print(json.dumps(dict(workload=name, ms_per_step=round(r["ms_per_step"], 4), b2b_ms=round(r["b2b_ms"], 4), with_phase_timers_ms=round(r["ph_step_ms"], 4),
                              e2e_ms=round(r["e2e"]["ms_per_step"], 4), launches_per_step=r["launches"] / steps,
                              phases_ms={n: round(float(v), 4) for n, v in zip(_cabi.PHASES, r["ph_ms"])})), flush=True)


if __name__ == "__main__":
    main()
