#!/usr/bin/env python
"""Benchmark of the BiLSTM + mention-span-head hot path (BASELINE.json metric: BiLSTM train captions/sec; mention-box pairs/sec).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--workload card2048|nonvis512|...] [--no-by-config]

One "step" = one `train_op` pass (forward + BPTT + clip + TF-Adam, dropout 0.5/0.5) of the hot path over one synthetic
F30kE-shaped batch.  The headline workload is BASELINE.json configs[1] (icl_core_lstm cardinality head, H=300, batch 2048
captions on one B200).  N>1: one process per GPU (torchrun), each rank trains on its own batch of the same size (weak
scaling) and the flat gradient buffer is all-reduced (SUM: the loss is a sum over examples, nn_utils/core.py:267) over NCCL
before the fused clip+Adam update.

JSON line: `value` = captions/s with the batch resident in HBM (icl_run_resident, CUDA events, max over ranks); `e2e` = the same
step through the reference-facing call `run_op(sess, train_op, [batch_tensors], ...)` with host NumPy buffers (H2D of the batch
and D2H of loss/accuracy inside the timed region, a rotation of distinct host batches); `e2e_variants` = that call with the
reference's own float64 feed, with `load_batch` inside the timed region, and with the corpus cache; `roofline` = the dominant
kernel group's algorithmic work / its CUDA-event time; `by_config` = the other BASELINE configs (C1 nonvis512, C3 intra / cross,
C4 affinity, C5 multitask) measured the same way in short runs; `cpu_baseline` = the NumPy oracle (a port of the reference's TF
graph; TensorFlow 1.x cannot run here) timed on a bounded sample of the same workload.
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")      # before the CUDA context exists: see imagecaptionlearn_py_b200/__init__.py
ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def _head(task, B, start, depth, C_, F, scope=""):
    return dict(task=task, B=B, start=start, depth=depth, C=C_, F=F, scope=scope)


WORKLOADS = {
    # BASELINE.json configs; hyper-parameters: the CLI defaults (icl_core_lstm.py:265-321) / config/lstm_{intra,cross}_params.config:96 /
    # icl_multitask_lstm.py:560-583 (H=200, 512-2, five heads in TASKS order :22)
    "nonvis512": dict(cfg="configs[0]", H=300, data_norm=False, heads=[_head("nonvis", 512, 512, 2, 2, 256)]),
    "card2048": dict(cfg="configs[1]", H=300, data_norm=False, heads=[_head("card", 2048, 512, 2, 12, 256)]),
    "rel_intra512": dict(cfg="configs[2]", H=200, data_norm=True, heads=[_head("rel_intra", 512, 1024, 3, 4, 512)]),
    "rel_cross512": dict(cfg="configs[2]", H=200, data_norm=True, heads=[_head("rel_cross", 512, 1024, 3, 4, 512)]),
    "affinity512": dict(cfg="configs[3]", H=300, data_norm=False, heads=[_head("affinity", 512, 512, 2, 2, 256)]),
    "multitask512": dict(cfg="configs[4]", H=200, data_norm=False, joint="simple_joint",
                         heads=[_head("rel_intra", 512, 512, 2, 4, 512, "rel_intra"), _head("rel_cross", 512, 512, 2, 4, 512, "rel_cross"),
                                _head("nonvis", 512, 512, 2, 2, 256, "nonvis"), _head("affinity", 512, 512, 2, 2, 256, "affinity"),
                                _head("card", 512, 512, 2, 12, 256, "card")]),
    "card2048_h200": dict(cfg="(bring-up)", H=200, data_norm=False, heads=[_head("card", 2048, 512, 2, 12, 256)]),
}
BY_CONFIG = ("nonvis512", "rel_intra512", "rel_cross512", "affinity512", "multitask512")
E, T_PAD, KEEP_IN, KEEP = 300, 50, 0.5, 0.5
LR, ADAM_EPS, CLIP = 1e-3, 1e-8, 5.0
BOX_W = 4096
ENC = "first_last_mention"

# legacy single-head view used by tests/test_gpu_full_size.py
for _n, _w in WORKLOADS.items():
    _h = _w["heads"][0]
    _w.update(task=_h["task"], B=_h["B"], start=_h["start"], depth=_h["depth"], C=_h["C"], F=_h["F"])


def _corpus_for(head, seed):
    from imagecaptionlearn_py_b200 import synth
    task, B = head["task"], head["B"]
    n_img = {"affinity": max(4, B // 200), "rel_cross": max(6, B // 60)}.get(task, max(8, B // 10))
    corpus = synth.make_corpus(n_img, seed=seed, with_boxes=(task == "affinity"))
    dd = synth.make_data_dict(corpus, task, F=head["F"])
    dd["max_seq_len"] = T_PAD                      # the reference pads to the dataset-global maximum (data.py:375)
    return dd


def _ids_for(head, dd, seed):
    from imagecaptionlearn_py_b200 import data as nn_data
    from imagecaptionlearn_py_b200 import synth
    task, B = head["task"], head["B"]
    ids = synth.example_ids(dd, task)
    rng = np.random.Generator(np.random.PCG64(seed + 1))
    if task == "affinity":
        ids = nn_data.shuffle_mention_box_pairs(ids, rng)
    else:
        ids = list(np.asarray(ids, dtype=object)[rng.permutation(len(ids))])
    if len(ids) < B:
        raise RuntimeError("synthetic corpus too small: %d ids for batch %d" % (len(ids), B))
    return ids


def make_batch(wl, seed, packed=False, dedup=False, head=0):
    """One reference-shaped batch_tensors dict (nn_utils/data.py:349-528) from the synthetic corpus, for head `head` of `wl`."""
    from imagecaptionlearn_py_b200 import data as nn_data
    h = wl["heads"][head]
    dd = _corpus_for(h, seed)
    ids = _ids_for(h, dd, seed)
    return nn_data.load_batch(ids[:h["B"]], dd, h["task"], h["C"], packed=packed, dedup=dedup)


def make_batches(wl, seed, packed=False):
    """The list run_op takes: one batch_tensors dict per head, in head order."""
    return [make_batch(wl, seed + 17 * i, packed=packed, head=i) for i in range(len(wl["heads"]))]


def make_rotation(wl, seed, n_rot, packed=False):
    """n_rot DISTINCT host batch lists over ONE corpus per head (one token / box table, as in a training run): rotation r takes the
    id list rolled by r * len / n_rot.  Rotation 0 equals make_batches(wl, seed, packed)."""
    from imagecaptionlearn_py_b200 import data as nn_data
    per_head = []
    for i, h in enumerate(wl["heads"]):
        dd = _corpus_for(h, seed + 17 * i)
        per_head.append((h, dd, _ids_for(h, dd, seed + 17 * i)))
    rot = []
    for r in range(n_rot):
        bl = []
        for h, dd, ids in per_head:
            k = (r * len(ids)) // n_rot
            rolled = ids[k:] + ids[:k]
            bl.append(nn_data.load_batch(rolled[:h["B"]], dd, h["task"], h["C"], packed=packed))
        rot.append(bl)
    return rot


def build_graph(wl):
    from imagecaptionlearn_py_b200 import core
    core.reset_default_graph()
    core.set_random_seeds()
    with core.variable_scope("bidirectional_lstm"):
        core.setup_bidirectional_lstm(wl["H"], wl["data_norm"], n_embedding_width=E)
    for h in wl["heads"]:
        args = (h["task"], ENC, h["B"], h["start"], h["depth"], False, "relu", h["C"], h["F"])
        kw = dict(box_embedding_width=BOX_W) if h["task"] == "affinity" else {}
        if h["scope"]:
            with core.variable_scope(h["scope"]):
                core.setup_core_architecture(*args, **kw)
        else:
            core.setup_core_architecture(*args, **kw)
    loss = core.setup_joint_loss(wl["joint"]) if wl.get("joint") else core.get_collection("loss")[0]
    core.add_train_op(loss, LR, ADAM_EPS, CLIP)
    return core.get_collection("train_op")[0]


def run_train(core, sess, train_op, wl, bts):
    core.run_op(sess, train_op, bts, KEEP_IN, KEEP, ENC, [h["task"] for h in wl["heads"]], [h["scope"] for h in wl["heads"]], True)


def flops_per_token(H, train=True):
    """valid-token FLOPs of the BiLSTM, both directions."""
    fwd = 4 * (E + H) * 4 * H
    return fwd * 2 + 4 * H * 4 * H if train else fwd


def head_dims(wl, h):
    from imagecaptionlearn_py_b200 import core
    H = wl["H"]
    d0 = (8 * H if h["task"].startswith("rel") else 4 * H) + h["F"] + (BOX_W if h["task"] == "affinity" else 0)
    return [d0] + core.get_widths(h["start"], h["depth"]) + [h["C"]]


def sample_clocks(stop, out):
    q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
    dev = os.environ.get("LOCAL_RANK", "0")
    try:        # NVML in-process: a sample every 5 ms (an nvidia-smi process per sample yields one or two per timed region)
        import pynvml as N
        N.nvmlInit()
        h = N.nvmlDeviceGetHandleByIndex(int(dev))
        reasons_fn = getattr(N, "nvmlDeviceGetCurrentClocksEventReasons", None) or N.nvmlDeviceGetCurrentClocksThrottleReasons
        mx = N.nvmlDeviceGetMaxClockInfo(h, N.NVML_CLOCK_SM)
        bits = (0x8, 0x40, 0x20, 0x4)          # hw_slowdown, hw_thermal_slowdown, sw_thermal_slowdown, sw_power_cap
        while not stop.is_set():
            r = int(reasons_fn(h))
            out.append([str(N.nvmlDeviceGetClockInfo(h, N.NVML_CLOCK_SM)), str(mx)] + ["Active" if r & b else "Not Active" for b in bits])
            stop.wait(0.005)
        return
    except Exception:
        pass
    while not stop.is_set():
        try:
            r = subprocess.run(["nvidia-smi", "-i", dev, "--query-gpu=" + q, "--format=csv,noheader,nounits"],
                               capture_output=True, text=True, timeout=5)
            f = [x.strip() for x in r.stdout.strip().split(",")]
            if len(f) >= 6:
                out.append(f)
        except Exception:
            pass
        stop.wait(0.2)


def clocks_summary(samples):
    if not samples:
        return dict(sm_mhz=None, sm_max_mhz=None, reasons=["unsampled"])
    sm = sorted(float(s[0]) for s in samples)
    names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
    reasons = [n for i, n in enumerate(names) if any(s[2 + i].lower().startswith("active") for s in samples)]
    return dict(sm_mhz=sm[len(sm) // 2], sm_max_mhz=float(samples[0][1]), reasons=reasons, samples=len(samples))


# ----------------------------------------------------------------------------------------------- reference arm (CPU)
def host_threads():
    """All the host cores for the CPU arm: torchrun exports OMP_NUM_THREADS=1, which would pin the BLAS behind NumPy to one
    thread -- lift the limit explicitly and report what is actually in effect."""
    n = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    try:
        import threadpoolctl
        threadpoolctl.threadpool_limits(limits=n)
        n = max([p.get("num_threads", 1) for p in threadpoolctl.threadpool_info()] or [n])
    except Exception:
        pass
    return int(n)


def oracle_train_step(params, cfg, bt, state, keep_in, keep, rng):
    from oracle import icl_oracle as O
    S, T = bt["sentences"].shape[:2]
    H = cfg["H"]
    B = len(bt["labels"])
    bern = lambda shape, k: (rng.random(shape) < k).astype(np.float32)
    masks = dict(in_fw=bern((S, T, E), keep_in), in_bw=bern((S, T, E), keep_in), out_fw=bern((S, T, H), keep),
                 out_bw=bern((S, T, H), keep), heads=[[bern((B, w), keep) for w in cfg["heads"][0]["widths"]]])
    f = O.model_forward(params, cfg, bt["sentences"], bt["seq_lengths"], [bt], keep_in, keep, masks)
    g = O.model_backward(params, cfg, f, [bt])
    O.clip_and_adam(params, g, state, LR, ADAM_EPS, CLIP)
    return float(f["loss"])


def cpu_sample(wl, sample_B, steps, warmup, target_s=None):
    """Time the NumPy oracle (port of the reference TF graph) on the first `sample_B` examples of the (single-head) workload.  With
    `target_s` the number of timed steps is chosen from the last warm-up step so that the sample is about that many seconds."""
    from oracle import icl_oracle as O
    from imagecaptionlearn_py_b200 import core
    cores = host_threads()
    h = dict(wl["heads"][0], B=sample_B)
    small = dict(wl, heads=[h])
    bt = make_batch(small, 20171201)
    tl = int(bt["seq_lengths"].max())
    bt["sentences"] = bt["sentences"][:, :tl]              # dynamic_rnn stops at the batch maximum (sequence_length)
    widths = core.get_widths(h["start"], h["depth"])
    H = wl["H"]
    box_w = bt["box_embeddings"].shape[1] if "box_embeddings" in bt else 0
    cfg = dict(H=H, data_norm=wl["data_norm"],
               heads=[dict(task=h["task"], scope="", encoding_scheme=ENC, n_layers=len(widths),
                           widths=widths, activation="relu", weighted_classes=False,
                           in_width=O.head_in_width(h["task"], ENC, H, h["F"], box_w), n_classes=h["C"])])
    rng = np.random.default_rng(7)
    params = O.init_params(rng, cfg, E, np.float32)
    state = {}
    t1 = 0.0
    for _ in range(warmup):
        t0 = time.perf_counter()
        oracle_train_step(params, cfg, bt, state, KEEP_IN, KEEP, rng)
        t1 = time.perf_counter() - t0
    if target_s and t1 > 0:
        steps = int(max(steps, min(60, round(target_s / t1))))
    t0 = time.perf_counter()
    for _ in range(steps):
        oracle_train_step(params, cfg, bt, state, KEEP_IN, KEEP, rng)
    dt = time.perf_counter() - t0
    return dict(value=sample_B * steps / dt, unit="captions/s", cores=int(cores), kind="port",
                sample="%d train steps of the NumPy oracle (fp32, BLAS threads=%d) on %d captions of the same workload"
                       % (steps, cores, sample_B), ms_per_step=1e3 * dt / steps, steps=steps, batch=sample_B)


def run_reference(args, wl):
    """The reference arm: the NumPy restatement of the reference's TF CPU graph on ALL host cores, on the same workload and the same
    batch per step as the B200 arm (rank 0 only; the other ranks exit)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    B = wl["heads"][0]["B"]
    r = cpu_sample(wl, B, args.steps, args.warmup)
    line = dict(metric="bilstm_train_captions_per_sec", value=r["value"], unit="captions/s", n_gpus=args.gpus, steps=args.steps,
                warmup=args.warmup, ms_per_step=r["ms_per_step"], higher_is_better=True, scaling="weak", vs_baseline=None,
                dtype="f32", data="synthetic", impl="reference",
                config=dict(workload=args.workload, baseline_config=wl["cfg"], task=wl["heads"][0]["task"], batch_per_step=B,
                            batch_per_gpu=B, lstm_hidden=wl["H"], embed=E, padded_T=T_PAD, host_threads=r["cores"],
                            note="TensorFlow 1.x / Python 2 are not installable here: the reference's TF CPU graph is "
                                 "restated in NumPy (oracle/icl_oracle.py) and timed on the host cores; one host runs one batch per "
                                 "step whatever --gpus says (the reference has no multi-GPU path)"),
                cpu_baseline=dict(value=r["value"], unit="captions/s", cores=r["cores"], kind="port", sample=r["sample"]),
                e2e=dict(value=r["value"], unit="captions/s", h2d_bytes_per_step=0, d2h_bytes_per_step=0), gpu_launches=0)
    _JSON_OUT.write(json.dumps(line) + "\n"); _JSON_OUT.flush()


# ----------------------------------------------------------------------------------------------- our arm (B200)
class Ctx(object):
    pass


def peaks():
    p = {}
    try:
        p = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    bf16 = p.get("bf16_tflops_sustained", 1400.0)            # fallback: the profiling guide's sustained figure
    return dict(bf16=bf16, tf32=bf16 / 2.0, hbm=p.get("hbm_gbs", 6650.0),        # kind::tf32 issues at half the kind::f16 rate
                source=("MEASURED_PEAKS.json: hbm_gbs; bf16_tflops_sustained (kind::f16), / 2 for kind::tf32" if p
                        else "fallback 6650 GB/s, 1400 (f16) / 700 (tf32) TFLOP/s"))


def phase_work(wl, n_tok):
    """ALGORITHMIC work per step of each kernel group: FLOPs of the contraction and the bytes that must cross
    HBM once, per valid token and direction x 2 directions; heads: per example.  phase -> (flops, bytes, kernel, launches, tensor peak key)"""
    H = wl["H"]
    tok2 = 2.0 * n_tok
    w = {"proj_gemm": (tok2 * 2 * E * 4 * H, tok2 * (E * 2 + 4 * H * 4), "k_gemm_tcgen05<0, 0, 256, 2, 1>", 2, "bf16"),    # fp16 x in, fp32 out
         "rec_fwd": (tok2 * 2 * H * 4 * H, tok2 * (4 * H + 4 * H + 3 * H) * 4, "k_rec_fwd16", 1, "bf16"),       # Zx in; gates, c, h, TF32(h) out
         "rec_bwd": (tok2 * 2 * H * 4 * H, tok2 * (4 * H + 3 * H + 4 * H) * 4, "k_bptt_cluster", 1, "tf32"),    # gates, c, c_prev, dH in; dZ out
         "wgrad": (tok2 * 2 * (E + H) * 4 * H, tok2 * (E + H + 4 * H) * 4, "k_gemm_tcgen05<1, 1, 256, 4, 0>", 2, "tf32")}
    fl = by = 0.0
    for h in wl["heads"]:
        d = head_dims(wl, h)
        mm = sum(d[k] * d[k + 1] for k in range(len(d) - 1))
        fl += 2.0 * h["B"] * mm                                      # 2 * sum_k w_k w_{k+1} per example (forward)
        # gather + concat (read the sources, write batch_input), weights once, activations once
        by += h["B"] * (2.0 * d[0] + sum(d[1:])) * 4 + mm * 4
    w["heads_fwd"] = (fl, by, "k_gather_concat + k_gemm_tcgen05 x L + k_softmax_ce", len(wl["heads"]), "tf32")
    w["heads_bwd"] = (2 * fl, 2 * by, "k_softmax_bwd + k_gemm_tcgen05 (dz chain, dW) + k_scatter_spans", len(wl["heads"]), "tf32")
    return w


def roofline_of(wl, n_tok, ph_ms, pk, traffic=None):
    from imagecaptionlearn_py_b200 import _cabi
    work = phase_work(wl, n_tok)
    dom = max(work, key=lambda k: ph_ms[_cabi.PHASES.index(k)])
    t_ms = float(ph_ms[_cabi.PHASES.index(dom)])
    flops, nbytes, kernel, n_launch, tp = work[dom]
    t_tensor, t_hbm = flops / (pk[tp] * 1e12), nbytes / (pk["hbm"] * 1e9)
    if t_hbm >= t_tensor:
        roof = dict(kernel=kernel, phase=dom, bound="hbm", unit="GB/s", peak=pk["hbm"], achieved=nbytes / (t_ms * 1e-3) / 1e9)
    else:
        roof = dict(kernel=kernel, phase=dom, bound="tensor", unit="TFLOP/s", peak=pk[tp], achieved=flops / (t_ms * 1e-3) / 1e12)
    roof.update(frac=roof["achieved"] / roof["peak"], traffic=traffic, launches_per_step=n_launch, ms_per_launch=t_ms / n_launch,
                algorithmic_bytes_per_launch=nbytes / n_launch, algorithmic_flops_per_launch=flops / n_launch,
                tensor_frac=flops / (t_ms * 1e-3) / 1e12 / pk[tp], hbm_frac=nbytes / (t_ms * 1e-3) / 1e9 / pk["hbm"], peak_source=pk["source"],
                note="time = CUDA events around the kernel group on its launch stream (icl_phase_ms), averaged over the timed steps; the "
                     "algorithmic bytes of the recurrences are this repo's accounting (SURVEY 8d gives none): gates, c, h / dH in and out once")
    by_phase = {}
    for k, wk in work.items():
        ms_k = float(ph_ms[_cabi.PHASES.index(k)])
        if ms_k > 0:
            by_phase[k] = dict(kernel=wk[2], ms=ms_k, hbm_frac=wk[1] / (ms_k * 1e-3) / 1e9 / pk["hbm"],
                               tensor_frac=wk[0] / (ms_k * 1e-3) / 1e12 / pk[wk[4]], tensor_peak_tflops=pk[wk[4]],
                               bound="hbm" if wk[1] / (pk["hbm"] * 1e9) >= wk[0] / (pk[wk[4]] * 1e12) else "tensor")
    return roof, by_phase


def measure(name, steps, warmup, ctx, clocks=None, e2e_variants=False):
    """One workload on this rank's GPU: device-timed resident steps (CUDA events, L2 flushed between steps, max over ranks) and the
    end-to-end run_op loop with host buffers.  Returns a dict (rank 0 fills the JSON from it)."""
    import torch
    from imagecaptionlearn_py_b200 import _cabi, core
    from imagecaptionlearn_py_b200 import data as nn_data
    wl = WORKLOADS[name]
    dist, rank, world, local = ctx.dist, ctx.rank, ctx.world, ctx.local
    seed0 = 20171201 + 1000 * (0 if os.environ.get("ICL_BENCH_SAME_BATCH") else rank)     # (every rank the same batch: isolates rank imbalance)
    bts = make_batches(wl, seed0)
    train_op = build_graph(wl)
    sess = core.Session(max_seq_len=T_PAD, device=local, dist=bool(dist), gemm_mode=ctx.gemm_mode)
    sess.ensure()
    L = _cabi.lib()
    if dist:                                   # identical initial weights on every rank
        dist.broadcast(sess.param_tensor(), 0)
    keepalive = []
    b = sess.build_batch(bts, True, keepalive)
    sess._bind_stream()
    _cabi.check(L.icl_upload(sess.handle, C.byref(b)))
    ns, nt, tm = C.c_int64(), C.c_int64(), C.c_int32()
    _cabi.check(L.icl_batch_stats(sess.handle, C.byref(ns), C.byref(nt), C.byref(tm)))
    n_seqs, n_tok, t_max = ns.value, nt.value, tm.value
    n_examples = sum(h["B"] for h in wl["heads"])
    # affinity heads: how layer 1 of the resident batch runs (factorised over the batch's distinct mentions / boxes)
    factor = {}
    for hi, h in enumerate(wl["heads"]):
        if h["task"] == "affinity":
            f_, nm_, nb_ = C.c_int32(), C.c_int32(), C.c_int32()
            _cabi.check(L.icl_head_factor_stats(sess.handle, hi, C.byref(f_), C.byref(nm_), C.byref(nb_), None))
            d0m, w1 = 4 * wl["H"] + h["F"], h["start"]
            factor[hi] = dict(formulation="factorised: z1 = U[mention] + V[box] + b1, one GEMM row per distinct mention / box" if f_.value
                              else "concatenated [mention | box] rows", pairs=h["B"], distinct_mentions=nm_.value if f_.value else h["B"],
                              distinct_boxes=nb_.value if f_.value else h["B"],
                              layer1_fwd_flops_concatenated=2 * h["B"] * (d0m + BOX_W) * w1,
                              layer1_fwd_flops_executed=2 * ((nm_.value * d0m + nb_.value * BOX_W) if f_.value else h["B"] * (d0m + BOX_W)) * w1,
                              note="training batches are reference-shaped (one caption copy per pair, nn_utils/data.py:397-403): the "
                                   "mentions do not repeat, the 4096 box columns collapse to the distinct boxes")

    def resident_step(i):
        seed = 1000 + i
        if dist:
            _cabi.check(L.icl_run_resident(sess.handle, _cabi.OP_GRADS, KEEP_IN, KEEP, seed))
            sess.allreduce_grads()
            _cabi.check(L.icl_apply_update(sess.handle))
        else:
            _cabi.check(L.icl_run_resident(sess.handle, _cabi.OP_TRAIN, KEEP_IN, KEEP, seed))
        _cabi.check(L.icl_join_side_work(sess.handle))      # the step's side-stream work (fp16 repack of the updated LSTM weights) is timed with it

    def max_over_ranks(x):
        if not dist:
            return x
        t = torch.tensor([x], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    for i in range(warmup):
        resident_step(i)
    torch.cuda.synchronize()
    n0 = C.c_int64()
    L.icl_kernel_launches(sess.handle, C.byref(n0))
    stop = threading.Event()
    th = None
    if clocks is not None and rank == 0:
        th = threading.Thread(target=sample_clocks, args=(stop, clocks), daemon=True)
        th.start()
    if dist:
        dist.barrier()
    torch.cuda.synchronize()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    _cabi.check(L.icl_set_phase_timing(sess.handle, 0))      # the headline steps run without the library's sixteen per-phase timing events
    wall0 = time.perf_counter()
    for i in range(steps):
        ctx.flush.zero_()                       # L2 flush between timed iterations (untimed)
        ev[i][0].record()
        resident_step(warmup + i)
        ev[i][1].record()
    torch.cuda.synchronize()
    if dist:
        dist.barrier()
    wall = time.perf_counter() - wall0
    stop.set()          # the clocks are sampled during the device-timed region only
    total_ms = sum(a.elapsed_time(b_) for a, b_ in ev)
    n1 = C.c_int64()
    L.icl_kernel_launches(sess.handle, C.byref(n1))       # kernels launched inside the device-timed region
    # per-phase times: a SEPARATE pass of the same steps with the library's timing events on (a timing event serialises the stream:
    # these steps run ~50 us longer than the headline ones, `ms_per_step_with_phase_timers`)
    _cabi.check(L.icl_set_phase_timing(sess.handle, 1))
    phases = np.zeros(_cabi.N_PHASES)
    pev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    for i in range(steps):
        ctx.flush.zero_()
        pev[i][0].record()
        resident_step(warmup + steps + i)
        pev[i][1].record()
        ph = (C.c_float * _cabi.N_PHASES)()
        _cabi.check(L.icl_phase_ms(sess.handle, ph))     # synchronises; CUDA-event time of each kernel group of this step
        phases += np.array(list(ph))
    torch.cuda.synchronize()
    ph_step_ms = max_over_ranks(sum(a.elapsed_time(b_) for a, b_ in pev)) / steps
    _cabi.check(L.icl_set_phase_timing(sess.handle, 0))
    # the same steps back to back (no flush, one event pair): what a training loop sees; the step's working set (~2 GB of activations at
    # card2048) is many times the 126 MB L2 by itself
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    if dist:
        dist.barrier()
    torch.cuda.synchronize()
    e0.record()
    for i in range(steps):
        seed = 5000 + i
        if dist:
            _cabi.check(L.icl_run_resident(sess.handle, _cabi.OP_GRADS, KEEP_IN, KEEP, seed))
            sess.allreduce_grads()
            _cabi.check(L.icl_apply_update(sess.handle))
        else:
            _cabi.check(L.icl_run_resident(sess.handle, _cabi.OP_TRAIN, KEEP_IN, KEEP, seed))
    _cabi.check(L.icl_join_side_work(sess.handle))
    e1.record()
    torch.cuda.synchronize()
    b2b_ms = e0.elapsed_time(e1) / steps

    ms_per_step = max_over_ranks(total_ms) / steps
    out = dict(wl=wl, n_seqs=n_seqs, n_tok=n_tok, t_max=t_max, n_examples=n_examples, ms_per_step=ms_per_step, ph_ms=phases / steps,
               launches=n1.value - n0.value, wall_ms=1e3 * wall / steps, value=world * n_seqs / (ms_per_step * 1e-3),
               b2b_ms=max_over_ranks(b2b_ms), factor=factor, ph_step_ms=ph_step_ms)

    # ---- end-to-end through the reference-facing API with host buffers: a rotation of N_ROT distinct host batches (a real
    # training loop never re-feeds a cache-warm buffer), as many timed steps as the device-timed leg
    N_ROT = 4
    rot = make_rotation(wl, seed0, N_ROT)

    def e2e_loop(batches_of, n, before=None):
        for i in range(2):
            run_train(core, sess, train_op, wl, batches_of(i))
        torch.cuda.synchronize()
        if dist:
            dist.barrier()
        t0 = time.perf_counter()
        for i in range(n):
            run_train(core, sess, train_op, wl, batches_of(i))
        torch.cuda.synchronize()
        s = max_over_ranks(time.perf_counter() - t0)
        h2d, d2h = C.c_int64(), C.c_int64()
        L.icl_copy_bytes(sess.handle, C.byref(h2d), C.byref(d2h))
        return dict(value=world * n_seqs * n / s, unit="captions/s", ms_per_step=1e3 * s / n, steps=n, h2d_bytes_per_step=h2d.value,
                    d2h_bytes_per_step=d2h.value, distinct_host_batches=N_ROT)
    out["e2e"] = e2e_loop(lambda i: rot[i % N_ROT], steps)
    out["e2e"]["api"] = ("core.run_op(sess, train_op, [batch_tensors], ...) with host float32 NumPy buffers ('sentences' [S,50,300]); "
                         "pipelined: the batch is packed into pinned memory and copied on a copy stream into the idle input set while the "
                         "previous step computes; loss/accuracy of the previous step are read back (D2H) every step")
    if e2e_variants:
        var = {}
        rot64 = [[dict(bt, sentences=bt["sentences"].astype(np.float64)) for bt in bl] for bl in rot]   # what nn_utils/data.py:375 builds
        var["float64_feed"] = e2e_loop(lambda i: rot64[i % N_ROT], steps)
        var["float64_feed"]["api"] = "same call, 'sentences' as float64 (np.zeros default of the reference's load_batch, nn_utils/data.py:375)"
        del rot64
        # host batch build inside the timed region: data.load_batch (vectorised) on fresh ids every step + run_op
        h0 = wl["heads"][0]
        if len(wl["heads"]) == 1:
            dd = _corpus_for(h0, seed0)
            ids = _ids_for(h0, dd, seed0)
            nb = max(1, len(ids) // h0["B"])
            for packed, key in ((False, "with_load_batch"), ("rows", "with_load_batch_corpus_cache")):
                def fresh(i, packed=packed):
                    lo = (i % nb) * h0["B"]
                    return [nn_data.load_batch(ids[lo:lo + h0["B"]], dd, h0["task"], h0["C"], packed=packed)]
                var[key] = e2e_loop(fresh, steps)
                var[key]["api"] = ("data.load_batch(ids, data_dict, task, n_classes%s) + run_op(train_op) per step, both inside the timed region"
                                   % (", packed='rows'" if packed else ""))
        rows = make_rotation(wl, seed0, N_ROT, packed="rows")
        var["corpus_cache"] = e2e_loop(lambda i: rows[i % N_ROT], steps)
        var["corpus_cache"]["api"] = ("same run_op call; load_batch(packed='rows'): int32 token rows into the device-resident token table "
                                      "(icl_set_token_table) instead of the padded [S,T,300] host tensor")
        out["e2e_variants"] = var
    sess.close()
    return out


def summarise(name, r, pk, world, traffic=None):
    """The by_config entry / the body of the headline line for one measured workload."""
    wl = r["wl"]
    roof, by_phase = roofline_of(wl, r["n_tok"], r["ph_ms"], pk, traffic)
    from imagecaptionlearn_py_b200 import _cabi
    d = dict(baseline_config=wl["cfg"], tasks=[h["task"] for h in wl["heads"]], batch_per_gpu=[h["B"] for h in wl["heads"]],
             lstm_hidden=wl["H"], sequences_per_step_per_gpu=r["n_seqs"], tokens_per_step_per_gpu=r["n_tok"],
             value=r["value"], unit="captions/s", ms_per_step=r["ms_per_step"],
             examples_per_sec=world * r["n_examples"] / (r["ms_per_step"] * 1e-3),
             e2e=dict(value=r["e2e"]["value"], unit="captions/s", ms_per_step=r["e2e"]["ms_per_step"],
                      examples_per_sec=world * r["n_examples"] / (r["e2e"]["ms_per_step"] * 1e-3),
                      h2d_bytes_per_step=r["e2e"]["h2d_bytes_per_step"], d2h_bytes_per_step=r["e2e"]["d2h_bytes_per_step"]),
             phases_ms={n: float(v) for n, v in zip(_cabi.PHASES, r["ph_ms"])}, ms_per_step_with_phase_timers=r["ph_step_ms"],
             roofline=dict(kernel=roof["kernel"], phase=roof["phase"], bound=roof["bound"], frac=roof["frac"], achieved=roof["achieved"],
                           peak=roof["peak"], unit=roof["unit"]),
             gpu_launches_per_step=r["launches"] / max(1, r.get("steps", 1)))
    if r.get("factor"):
        d["affinity_layer1"] = {wl["heads"][hi]["scope"] or "affinity": v for hi, v in r["factor"].items()}
    if wl["heads"][0]["task"] == "affinity" and len(wl["heads"]) == 1:
        d["unit_pairs"] = "mention-box pairs/s = examples_per_sec"
    return d, roof, by_phase


def host_memory_bandwidth(ctx, threads=4, mb=256, reps=4):
    """STREAM-style copy on the host, on every rank at the same time with the thread count a rank's packing pool uses: what the box's
    memory system gives the ranks when all of them pack.  Returns GB/s (read + written bytes) of this rank and the sum over ranks."""
    import torch
    old = torch.get_num_threads()
    torch.set_num_threads(threads)
    a = torch.empty(mb << 18, dtype=torch.float32).fill_(1.0)
    b = torch.empty_like(a)
    b.copy_(a)
    if ctx.dist:
        ctx.dist.barrier()
    t0 = time.perf_counter()
    for _ in range(reps):
        b.copy_(a)
    dt = time.perf_counter() - t0
    torch.set_num_threads(old)
    mine = 2.0 * a.numel() * 4 * reps / dt / 1e9
    total = mine
    if ctx.dist:
        t = torch.tensor([mine], device="cuda", dtype=torch.float64)
        ctx.dist.all_reduce(t)
        total = float(t.item())
    return dict(per_rank_gbs=mine, all_ranks_gbs=total, threads_per_rank=threads, how="torch CPU copy of %d MiB x %d on every rank at once" % (mb, reps))


def affinity_predict_pairs_per_sec(local, steps=10):
    """BASELINE.json's second metric on the PREDICT side: pairs/s end to end through get_pred_scores_mcc's batch path (keep 1.0, every
    distinct caption of a batch encoded once, token rows + box rows into the device-resident tables)."""
    import torch
    from imagecaptionlearn_py_b200 import core
    wl = WORKLOADS["affinity512"]
    h = wl["heads"][0]
    build_graph(wl)
    proba_op = core.get_collection("predicted_proba")[0]
    sess = core.Session(max_seq_len=T_PAD, device=local)
    sess.ensure()
    bt_pred = make_batch(wl, 20171201, packed="rows", dedup=True)
    fn = lambda: core.run_op(sess, proba_op, [bt_pred], 1.0, 1.0, ENC, [h["task"]], [""], False)
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(steps):
        fn()
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / steps
    sess.close()
    return dict(predict_e2e=h["B"] / dt, unit="pairs/s", predict_e2e_ms_per_step=1e3 * dt, predict_distinct_captions=int(len(bt_pred["seq_lengths"])))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="card2048", choices=sorted(WORKLOADS))
    ap.add_argument("--gemm", default="tf32", choices=["tf32", "simt"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-by-config", action="store_true")
    ap.add_argument("--by-config-steps", type=int, default=8)
    args = ap.parse_args()
    wl = WORKLOADS[args.workload]
    # stdout carries the ONE JSON line and nothing else: whatever a library writes to fd 1 during the run (NCCL prints its version
    # there when the box sets NCCL_DEBUG) goes to stderr instead
    global _JSON_OUT
    sys.stdout.flush()
    _JSON_OUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    if args.impl == "reference":
        return run_reference(args, wl)
    args.warmup = max(args.warmup, 3)

    import torch
    from imagecaptionlearn_py_b200 import _cabi
    ctx = Ctx()
    ctx.rank, ctx.world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    ctx.local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a B200: the hot path has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(ctx.local)
    ctx.dist = None
    if ctx.world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", ctx.local))
        ctx.dist = dist
    ctx.gemm_mode = _cabi.GEMM_TCGEN05_TF32 if args.gemm == "tf32" else _cabi.GEMM_SIMT_FP32
    ctx.flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")      # > 126 MB L2
    rank, world = ctx.rank, ctx.world

    clk = []
    main_r = measure(args.workload, args.steps, args.warmup, ctx, clocks=clk, e2e_variants=True)
    main_r["steps"] = args.steps
    host_bw = host_memory_bandwidth(ctx)
    others = {}
    if not args.no_by_config:
        for name in BY_CONFIG:
            if name != args.workload:
                others[name] = measure(name, args.by_config_steps, 3, ctx)
                others[name]["steps"] = args.by_config_steps

    if rank == 0:
        pk = peaks()
        traffic = None
        try:       # DRAM bytes of the dominant kernel from the committed `ncu --set full` capture of this workload (profiles/)
            if args.workload == "card2048":
                for fn in ("r2f_ncu_recurrences.json", "r2_ncu_full_summary.json", "r2a_ncu_recurrences.json", "r1h_ncu_full_summary.json"):
                    path = os.path.join(ROOT, "profiles", fn)
                    if os.path.exists(path):
                        ents = json.load(open(path))
                        work = phase_work(wl, main_r["n_tok"])
                        dom = max(work, key=lambda k: main_r["ph_ms"][_cabi.PHASES.index(k)])
                        seen = {}
                        for e in ents:
                            if e["kernel"].startswith(work[dom][2]) and e["kernel"] not in seen:
                                seen[e["kernel"]] = e["dram_bytes"]
                        if seen:
                            traffic = sum(seen.values())
                            break
        except Exception:
            pass
        body, roof, by_phase = summarise(args.workload, main_r, pk, world, traffic)
        if traffic is not None:
            roof["traffic_source"] = "profiles/: dram__bytes_read.sum + dram__bytes_write.sum of the committed ncu --set full capture (N=1, card2048)"
        H = wl["H"]
        step_flops = flops_per_token(H) * main_r["n_tok"]
        ms_per_step = main_r["ms_per_step"]
        line = dict(metric="bilstm_train_captions_per_sec", value=main_r["value"], unit="captions/s", n_gpus=world, steps=args.steps,
                    warmup=args.warmup, ms_per_step=ms_per_step, higher_is_better=True, scaling="weak", vs_baseline=None,
                    dtype="tf32" if args.gemm == "tf32" else "f32", data="synthetic",
                    config=dict(workload=args.workload, baseline_config=wl["cfg"], task=wl["heads"][0]["task"], batch_per_gpu=wl["heads"][0]["B"],
                                global_batch=wl["heads"][0]["B"] * world, lstm_hidden=H, embed=E, padded_T=T_PAD, t_max=main_r["t_max"],
                                tokens_per_step_per_gpu=main_r["n_tok"], keep_prob=[KEEP_IN, KEEP], clip_norm=CLIP,
                                parallelism="dp%d" % world, l2="flushed between timed steps (256 MiB memset, untimed)",
                                tokens_per_sec=world * main_r["n_tok"] / (ms_per_step * 1e-3),
                                bilstm_tflops=step_flops / (ms_per_step * 1e-3) / 1e12,
                                wall_ms_per_step_incl_flush=main_r["wall_ms"],
                                back_to_back_ms_per_step=main_r["b2b_ms"],
                                ms_per_step_with_phase_timers=main_r["ph_step_ms"],
                                phase_timers="phases_ms / roofline come from a separate pass of the same steps with the library's per-phase "
                                             "timing events on (icl_set_phase_timing): a timing event serialises the stream, so those steps "
                                             "are longer than the headline ones"),
                    phases_ms=body["phases_ms"], roofline=roof, roofline_by_phase=by_phase,
                    e2e=main_r["e2e"], e2e_variants=main_r.get("e2e_variants"),
                    gpu_launches=main_r["launches"], clocks=clocks_summary(clk))
        # the host-tensor e2e path per rank and step: reads the valid rows of 'sentences' (4 B/elem), writes the packed pinned mirror
        # (fp16 wire: 2 B/elem), and the DMA engine reads that mirror again -- against what the box's memory system gives all ranks
        e = main_r["e2e"]
        elems = main_r["n_tok"] * E
        traffic = elems * 4 + 2 * (e["h2d_bytes_per_step"])
        line["e2e_host_limit"] = dict(host_bytes_per_rank_step=int(traffic), all_ranks_gbs_needed_at_device_rate=world * traffic / (ms_per_step * 1e-3) / 1e9,
                                      all_ranks_gbs_achieved=world * traffic / (e["ms_per_step"] * 1e-3) / 1e9, host_copy_bandwidth=host_bw,
                                      note="host sentence tensors: valid rows read (fp32) + packed mirror written + mirror read by the DMA engine; "
                                           "when achieved ~ host_copy_bandwidth.all_ranks_gbs the host memory system is the limiter (the corpus-cache "
                                           "path sends 4-byte token rows instead)")
        if "corpus_cache" in (main_r.get("e2e_variants") or {}):
            line["e2e_resident_corpus"] = main_r["e2e_variants"]["corpus_cache"]
        if others:
            line["by_config"] = {}
            for name, r in others.items():
                line["by_config"][name] = summarise(name, r, pk, world)[0]
            if "affinity512" in others:
                a = line["by_config"]["affinity512"]
                line["mention_box_pairs_per_sec"] = dict(workload="affinity512", baseline_config=WORKLOADS["affinity512"]["cfg"], unit="pairs/s",
                                                         pairs_per_step_per_gpu=512, train_resident=a["examples_per_sec"],
                                                         train_resident_ms_per_step=a["ms_per_step"], train_e2e=a["e2e"]["examples_per_sec"],
                                                         train_e2e_ms_per_step=a["e2e"]["ms_per_step"])
                if world == 1:
                    line["mention_box_pairs_per_sec"].update(affinity_predict_pairs_per_sec(ctx.local))
        if world == 1 and not args.no_cpu_baseline:        # rank 0 at N=1 only: the other ranks would idle behind it
            line["cpu_baseline"] = {k: v for k, v in cpu_sample(wl, min(wl["heads"][0]["B"], 512), 3, 2, target_s=12.0).items()
                                    if k in ("value", "unit", "cores", "kind", "sample")}
        _JSON_OUT.write(json.dumps(line) + "\n"); _JSON_OUT.flush()
    if ctx.dist:
        ctx.dist.barrier()
        ctx.dist.destroy_process_group()


if __name__ == "__main__":
    main()
