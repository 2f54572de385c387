"""ctypes binding of libicl_b200.so (C-ABI declared in include/icl_b200.h).

There is no CPU fallback: if the shared library is missing, or no sm_100 device is present when a model is
created, this raises.  Build the library with `python -c "import __graft_entry__ as g; g.build()"` or
`make -C imagecaptionlearn_py_b200/csrc`.
"""
import ctypes as C
import os

import numpy as np

MAX_HEADS, MAX_LAYERS, N_INDEX = 8, 8, 12
TASKS = {"nonvis": 0, "card": 1, "rel_intra": 2, "rel_cross": 3, "affinity": 4}
ENCODINGS = {"first_last_mention": 0, "first_last_sentence": 1}
ACTIVATIONS = {None: 0, "none": 0, "sigmoid": 1, "tanh": 2, "relu": 3, "leaky_relu": 4}
INDEX_ORDER = ("first_i_bw", "first_i_fw", "last_i_fw", "last_i_bw", "sent_last_i_fw", "sent_first_i_bw",
               "first_j_bw", "last_j_fw", "first_j_fw", "last_j_bw", "sent_last_j_fw", "sent_first_j_bw")
OP_PREDICT, OP_GRADS, OP_TRAIN = 0, 1, 2
GEMM_TCGEN05_TF32, GEMM_SIMT_FP32 = 0, 1
_DT = {np.dtype(np.float32): 0, np.dtype(np.float64): 1, np.dtype(np.int32): 2, np.dtype(np.int64): 3}

LIB_PATH = os.environ.get("ICL_B200_LIB") or os.path.join(os.path.dirname(os.path.abspath(__file__)), "libicl_b200.so")


class HeadConfig(C.Structure):
    _fields_ = [("task", C.c_int32), ("encoding", C.c_int32), ("batch_size", C.c_int32), ("n_classes", C.c_int32),
                ("n_feats", C.c_int32), ("box_width", C.c_int32), ("n_box_feats", C.c_int32), ("n_hidden", C.c_int32),
                ("widths", C.c_int32 * MAX_LAYERS), ("activation", C.c_int32), ("weighted_classes", C.c_int32),
                ("scope", C.c_char * 32)]


class Config(C.Structure):
    _fields_ = [("embed_width", C.c_int32), ("lstm_hidden", C.c_int32), ("data_norm", C.c_int32),
                ("max_seqs", C.c_int32), ("max_seq_len", C.c_int32), ("n_heads", C.c_int32),
                ("heads", HeadConfig * MAX_HEADS), ("learn_rate", C.c_float), ("adam_epsilon", C.c_float),
                ("clip_norm", C.c_float), ("beta1", C.c_float), ("beta2", C.c_float), ("device", C.c_int32),
                ("gemm_mode", C.c_int32)]


class HeadBatch(C.Structure):
    _fields_ = [("idx", C.c_void_p * N_INDEX), ("idx_dtype", C.c_int32),
                ("feats", C.c_void_p), ("feats_dtype", C.c_int32),
                ("box", C.c_void_p), ("box_dtype", C.c_int32),
                ("bfeats", C.c_void_p), ("bfeats_dtype", C.c_int32),
                ("labels", C.c_void_p), ("labels_dtype", C.c_int32),
                ("sent_offset", C.c_int32), ("box_rows", C.c_void_p), ("inactive", C.c_int32),
                ("sentences", C.c_void_p), ("n_seqs", C.c_int32), ("padded_T", C.c_int32)]


class Batch(C.Structure):
    _fields_ = [("sentences", C.c_void_p), ("sent_dtype", C.c_int32), ("sent_packed", C.c_int32),
                ("token_rows", C.c_void_p), ("seq_lengths", C.c_void_p), ("len_dtype", C.c_int32), ("n_seqs", C.c_int32), ("padded_T", C.c_int32),
                ("seq_gid_offset", C.c_int64), ("ex_gid_offset", C.c_int64), ("n_heads", C.c_int32),
                ("heads", HeadBatch * MAX_HEADS)]


class HeadOut(C.Structure):
    _fields_ = [("proba", C.POINTER(C.c_float)), ("pred", C.POINTER(C.c_int64)), ("loss", C.c_float),
                ("accuracy", C.c_float)]


# every symbol include/icl_b200.h declares (a CPU test checks the library exports all of them)
SYMBOLS = ["icl_last_error", "icl_version", "icl_create", "icl_destroy", "icl_set_stream", "icl_param_count",
           "icl_param_info", "icl_get_tensor", "icl_set_tensor", "icl_get_step", "icl_set_step", "icl_run",
           "icl_upload", "icl_run_resident", "icl_fetch", "icl_grad_buffer", "icl_param_buffer", "icl_apply_update",
           "icl_sync", "icl_get_lstm_outputs", "icl_get_batch_input", "icl_get_activation", "icl_rec_trace", "icl_debug_mask", "icl_gemm",
           "icl_kernel_launches", "icl_last_step_ms", "icl_phase_ms", "icl_copy_bytes", "icl_batch_stats", "icl_train_async",
           "icl_poll_stats", "icl_set_optimizer_slot", "icl_optimizer_slots",
           "icl_set_loss_weights", "icl_apply_update_ex", "icl_set_token_table", "icl_set_box_table", "icl_grad_split", "icl_wait_head_grads", "icl_grad_split_lstm", "icl_wait_fw_lstm_grads", "icl_join_side_work", "icl_debug_poison_recurrence", "icl_adopt_grad_buffer", "icl_nvls_allreduce", "icl_crc32c", "icl_pack_rows", "icl_head_factor_stats", "icl_set_phase_timing", "icl_group_rows"]
N_PHASES = 8
PHASES = ("prep", "proj_gemm", "rec_fwd", "heads_fwd", "heads_bwd", "rec_bwd", "wgrad", "update")

_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError("libicl_b200.so not built (%s); run __graft_entry__.build() -- there is no CPU fallback"
                               % LIB_PATH)
        L = C.CDLL(LIB_PATH)
        L.icl_last_error.restype = C.c_char_p
        L.icl_create.argtypes = [C.POINTER(Config), C.POINTER(C.c_void_p)]
        L.icl_destroy.argtypes = [C.c_void_p]
        L.icl_destroy.restype = None
        L.icl_set_stream.argtypes = [C.c_void_p, C.c_void_p]
        L.icl_param_count.argtypes = [C.c_void_p]
        L.icl_param_info.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_char_p), C.POINTER(C.c_int32),
                                     C.POINTER(C.c_int32), C.POINTER(C.c_int64)]
        L.icl_get_tensor.argtypes = [C.c_void_p, C.c_int, C.c_char_p, C.c_void_p]
        L.icl_set_tensor.argtypes = [C.c_void_p, C.c_int, C.c_char_p, C.c_void_p]
        L.icl_get_step.argtypes = [C.c_void_p, C.POINTER(C.c_int64)]
        L.icl_set_step.argtypes = [C.c_void_p, C.c_int64]
        L.icl_run.argtypes = [C.c_void_p, C.c_int, C.POINTER(Batch), C.c_float, C.c_float, C.c_uint64, C.POINTER(HeadOut)]
        L.icl_upload.argtypes = [C.c_void_p, C.POINTER(Batch)]
        L.icl_run_resident.argtypes = [C.c_void_p, C.c_int, C.c_float, C.c_float, C.c_uint64]
        L.icl_fetch.argtypes = [C.c_void_p, C.POINTER(HeadOut)]
        L.icl_train_async.argtypes = [C.c_void_p, C.POINTER(Batch), C.c_float, C.c_float, C.c_uint64, C.POINTER(HeadOut)]
        L.icl_poll_stats.argtypes = [C.c_void_p, C.POINTER(HeadOut)]
        L.icl_set_optimizer_slot.argtypes = [C.c_void_p, C.c_int]
        L.icl_optimizer_slots.argtypes = [C.c_void_p]
        L.icl_crc32c.argtypes = [C.c_void_p, C.c_uint64, C.c_uint32]
        L.icl_crc32c.restype = C.c_uint32
        L.icl_grad_split.argtypes = [C.c_void_p, C.POINTER(C.c_int64)]
        L.icl_wait_head_grads.argtypes = [C.c_void_p, C.c_void_p]
        L.icl_grad_split_lstm.argtypes = [C.c_void_p, C.POINTER(C.c_int64)]
        L.icl_wait_fw_lstm_grads.argtypes = [C.c_void_p, C.c_void_p]
        L.icl_join_side_work.argtypes = [C.c_void_p]
        L.icl_debug_poison_recurrence.argtypes = [C.c_void_p]
        L.icl_adopt_grad_buffer.argtypes = [C.c_void_p, C.c_void_p, C.c_int64]
        L.icl_nvls_allreduce.argtypes = [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32]
        L.icl_set_box_table.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_int32]
        L.icl_set_token_table.argtypes = [C.c_void_p, C.c_void_p, C.c_int64]
        L.icl_set_loss_weights.argtypes = [C.c_void_p, C.POINTER(C.c_float)]
        L.icl_apply_update_ex.argtypes = [C.c_void_p, C.c_double, C.POINTER(C.c_float)]
        L.icl_grad_buffer.argtypes = [C.c_void_p, C.POINTER(C.c_void_p), C.POINTER(C.c_int64)]
        L.icl_param_buffer.argtypes = [C.c_void_p, C.POINTER(C.c_void_p), C.POINTER(C.c_int64)]
        L.icl_apply_update.argtypes = [C.c_void_p]
        L.icl_sync.argtypes = [C.c_void_p]
        L.icl_get_lstm_outputs.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
        L.icl_get_batch_input.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
        L.icl_get_activation.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p]
        L.icl_rec_trace.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
        L.icl_pack_rows.argtypes = [C.c_void_p, C.c_int32, C.c_int64, C.c_void_p, C.c_int32]
        L.icl_head_factor_stats.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_int32), C.POINTER(C.c_int32), C.POINTER(C.c_int32),
                                            C.POINTER(C.c_int64)]
        L.icl_debug_mask.argtypes = [C.c_void_p, C.c_uint64, C.c_uint32, C.c_int64, C.c_int64, C.c_float, C.c_void_p]
        L.icl_gemm.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p,
                               C.c_void_p, C.c_int]
        L.icl_phase_ms.argtypes = [C.c_void_p, C.POINTER(C.c_float)]
        L.icl_set_phase_timing.argtypes = [C.c_void_p, C.c_int]
        L.icl_group_rows.argtypes = [C.c_void_p, C.c_int64, C.c_int64, C.c_void_p, C.POINTER(C.c_int32)]
        L.icl_copy_bytes.argtypes = [C.c_void_p, C.POINTER(C.c_int64), C.POINTER(C.c_int64)]
        L.icl_batch_stats.argtypes = [C.c_void_p, C.POINTER(C.c_int64), C.POINTER(C.c_int64), C.POINTER(C.c_int32)]
        L.icl_kernel_launches.argtypes = [C.c_void_p, C.POINTER(C.c_int64)]
        L.icl_last_step_ms.argtypes = [C.c_void_p, C.POINTER(C.c_float)]
        _lib = L
    return _lib


def check(rc):
    if rc != 0:
        raise RuntimeError("icl_b200: " + lib().icl_last_error().decode("utf-8", "replace"))


def np_ptr(a):
    return a.ctypes.data_as(C.c_void_p)


def dtype_code(a):
    return _DT[a.dtype]


def as_supported(a):
    """Return a C-contiguous array in one of the dtypes the C-ABI reads natively (no value change)."""
    a = np.asarray(a)
    if a.dtype not in _DT:
        a = a.astype(np.float64 if a.dtype.kind == "f" else np.int64)
    return np.ascontiguousarray(a)
