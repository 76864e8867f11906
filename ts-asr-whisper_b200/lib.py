"""ctypes binding of libdicow_b200.so (see include/dicow_b200.h).

The library is the only compute backend: if it is missing or the device is not sm_100 every op raises --
there is no eager / CPU fallback.
"""
from __future__ import annotations

import ctypes as C
import os
import threading

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libdicow_b200.so")

# dicow_epilogue_t
EPI_BIAS_BF16 = 0
EPI_BIAS_GELU_BF16 = 1
EPI_RESIDUAL_F32 = 2
EPI_BIAS_F32 = 3
EPI_GELU_FDDT_POS_F32 = 4
EPI_ACCUM_F32 = 5
EPI_GELU_SAVE_BF16 = 6
EPI_DGELU_BF16 = 7
GEMM_A_T = 4  # dicow_gemm_args_t.flags: A given transposed (At[k][m])
GEMM_W_T = 8  # W given transposed (Wt[k][n])


class DicowError(RuntimeError):
    pass


class GemmArgs(C.Structure):
    _fields_ = [
        ("struct_size", C.c_size_t),
        ("A", C.c_void_p), ("lda", C.c_int64), ("a_batch_stride", C.c_int64),
        ("A2", C.c_void_p), ("lda2", C.c_int64), ("a2_batch_stride", C.c_int64),
        ("K1", C.c_int32),
        ("W", C.c_void_p), ("ldw", C.c_int64),
        ("nb", C.c_int32), ("Mb", C.c_int32), ("N", C.c_int32), ("K", C.c_int32),
        ("bias", C.c_void_p),
        ("out", C.c_void_p), ("ldo", C.c_int64), ("out_batch_stride", C.c_int64),
        ("epilogue", C.c_int32),
        ("resid", C.c_void_p), ("ldr", C.c_int64), ("resid_batch_stride", C.c_int64),
        ("gate", C.c_void_p),
        ("stno", C.c_void_p), ("stno_batch_stride", C.c_int64),
        ("fddt_w", C.c_void_p), ("fddt_b", C.c_void_p), ("pos", C.c_void_p), ("flags", C.c_int32), ("splits", C.c_int32), ("aux_bf16", C.c_void_p),
    ]


class FddtLnArgs(C.Structure):
    _fields_ = [
        ("struct_size", C.c_size_t),
        ("x", C.c_void_p),
        ("rows", C.c_int32), ("d", C.c_int32), ("T", C.c_int32),
        ("stno", C.c_void_p), ("stno_batch_stride", C.c_int64),
        ("fddt_w", C.c_void_p), ("fddt_b", C.c_void_p),
        ("gamma", C.c_void_p), ("beta", C.c_void_p),
        ("eps", C.c_float),
        ("ln_out_bf16", C.c_void_p), ("ln_out_f32", C.c_void_p), ("x_out_bf16", C.c_void_p),
        ("delta1_bf16", C.c_void_p), ("delta2_bf16", C.c_void_p), ("store_x", C.c_int32), ("flags", C.c_int32),
        ("x_out", C.c_void_p),
    ]


class AttentionArgs(C.Structure):
    _fields_ = [
        ("struct_size", C.c_size_t),
        ("Q", C.c_void_p), ("K", C.c_void_p), ("V", C.c_void_p), ("out", C.c_void_p),
        ("B", C.c_int32), ("H", C.c_int32), ("Tq", C.c_int32), ("Tk", C.c_int32),
        ("q_row_stride", C.c_int64), ("q_batch_stride", C.c_int64),
        ("kv_row_stride", C.c_int64), ("kv_batch_stride", C.c_int64),
        ("o_row_stride", C.c_int64), ("o_batch_stride", C.c_int64),
        ("causal", C.c_int32), ("variant", C.c_int32), ("lse", C.c_void_p),
    ]


class AttentionBwdArgs(C.Structure):
    _fields_ = [
        ("struct_size", C.c_size_t),
        ("Q", C.c_void_p), ("K", C.c_void_p), ("V", C.c_void_p), ("O", C.c_void_p), ("dO", C.c_void_p),
        ("lse", C.c_void_p), ("dQ", C.c_void_p), ("dK", C.c_void_p), ("dV", C.c_void_p),
        ("B", C.c_int32), ("H", C.c_int32), ("Tq", C.c_int32), ("Tk", C.c_int32),
        ("q_row_stride", C.c_int64), ("q_batch_stride", C.c_int64),
        ("kv_row_stride", C.c_int64), ("kv_batch_stride", C.c_int64),
        ("o_row_stride", C.c_int64), ("o_batch_stride", C.c_int64),
        ("dq_row_stride", C.c_int64), ("dq_batch_stride", C.c_int64),
        ("dkv_row_stride", C.c_int64), ("dkv_batch_stride", C.c_int64),
        ("causal", C.c_int32), ("workspace", C.c_void_p), ("workspace_floats", C.c_int64),
    ]


class LnBwdArgs(C.Structure):
    _fields_ = [
        ("struct_size", C.c_size_t),
        ("x", C.c_void_p), ("delta1_bf16", C.c_void_p), ("delta2_bf16", C.c_void_p),
        ("rows", C.c_int32), ("d", C.c_int32), ("T", C.c_int32),
        ("stno", C.c_void_p), ("stno_batch_stride", C.c_int64), ("fddt_w", C.c_void_p), ("fddt_b", C.c_void_p),
        ("gamma", C.c_void_p), ("eps", C.c_float), ("dy_bf16", C.c_void_p), ("g_in", C.c_void_p),
        ("g_out", C.c_void_p), ("g_out_bf16", C.c_void_p), ("dgamma", C.c_void_p), ("dbeta", C.c_void_p),
        ("dfddt_w", C.c_void_p), ("dfddt_b", C.c_void_p), ("g_colsum", C.c_void_p),
    ]


class CtcBwdArgs(C.Structure):
    _fields_ = [
        ("struct_size", C.c_size_t),
        ("logits", C.c_void_p), ("lse", C.c_void_p), ("B", C.c_int32), ("T", C.c_int32), ("V1", C.c_int32),
        ("labels", C.c_void_p), ("Lmax", C.c_int32), ("reduction_mean", C.c_int32), ("loss_scale", C.c_float),
        ("workspace", C.c_void_p), ("dlogits_bf16", C.c_void_p), ("ldd", C.c_int64),
        ("scale_dev", C.c_void_p), ("out_f32", C.c_int32), ("ld", C.c_int64),
    ]


class SoftlabelCeBwdArgs(C.Structure):
    _fields_ = [
        ("struct_size", C.c_size_t),
        ("logits", C.c_void_p), ("ld", C.c_int64), ("rows", C.c_int32), ("V", C.c_int32),
        ("labels", C.c_void_p), ("upp_labels", C.c_void_p), ("ts_begin", C.c_int32), ("n_ts", C.c_int32),
        ("smoothing", C.c_void_p), ("soft_mode", C.c_int32), ("scale", C.c_float),
        ("dlogits_bf16", C.c_void_p), ("ldd", C.c_int64), ("scale_dev", C.c_void_p),
    ]


class LogmelArgs(C.Structure):
    _fields_ = [
        ("struct_size", C.c_size_t),
        ("audio", C.c_void_p), ("audio_batch_stride", C.c_int64),
        ("B", C.c_int32), ("n_pad", C.c_int64),
        ("lengths", C.c_void_p), ("mel_filters", C.c_void_p), ("n_mels", C.c_int32),
        ("out", C.c_void_p), ("attention_mask", C.c_void_p), ("workspace", C.c_void_p),
    ]


class GemmSkinnyArgs(C.Structure):
    _fields_ = [
        ("struct_size", C.c_size_t),
        ("A", C.c_void_p), ("lda", C.c_int64), ("W", C.c_void_p), ("ldw", C.c_int64),
        ("M", C.c_int32), ("N", C.c_int32), ("K", C.c_int32),
        ("bias", C.c_void_p), ("out", C.c_void_p), ("ldo", C.c_int64), ("epilogue", C.c_int32),
        ("resid", C.c_void_p), ("ldr", C.c_int64), ("pos", C.c_void_p), ("pos_stride", C.c_int64),
    ]


class DecodeLinearArgs(C.Structure):
    _fields_ = [
        ("struct_size", C.c_size_t),
        ("x", C.c_void_p), ("ldx", C.c_int64), ("gamma", C.c_void_p), ("beta", C.c_void_p), ("eps", C.c_float),
        ("A", C.c_void_p), ("lda", C.c_int64), ("W", C.c_void_p), ("ldw", C.c_int64),
        ("M", C.c_int32), ("N", C.c_int32), ("K", C.c_int32),
        ("bias", C.c_void_p), ("out", C.c_void_p), ("ldo", C.c_int64), ("epilogue", C.c_int32),
        ("resid", C.c_void_p), ("ldr", C.c_int64), ("n_split", C.c_int32), ("out2", C.c_void_p), ("ldo2", C.c_int64),
        ("pos", C.c_void_p), ("pos_stride", C.c_int64),
    ]


class DecodeLayerArgs(C.Structure):  # one entry of the device-resident layer table of dicow_decode_layers
    _fields_ = [(n, C.c_void_p) for n in (
        "ln1_g", "ln1_b", "ln2_g", "ln2_b", "ln3_g", "ln3_b", "wqkv", "bqkv", "wo_self", "bo_self", "wq_cross", "bq_cross",
        "wo_cross", "bo_cross", "w1", "b1", "w2", "b2", "self_kv", "cross_kv")]


class DecodeLayersArgs(C.Structure):
    _fields_ = [
        ("struct_size", C.c_size_t),
        ("B", C.c_int32), ("d", C.c_int32), ("H", C.c_int32), ("ffn", C.c_int32), ("L", C.c_int32), ("T", C.c_int32),
        ("S_max", C.c_int32), ("vocab", C.c_int32),
        ("layers", C.c_void_p), ("ids", C.c_void_p), ("ids_row_stride", C.c_int64),
        ("embed_tokens", C.c_void_p), ("embed_positions", C.c_void_p), ("pos", C.c_void_p),
        ("x", C.c_void_p), ("q", C.c_void_p), ("ctx", C.c_void_p), ("hidden", C.c_void_p), ("barrier", C.c_void_p),
        ("attn_workspace", C.c_void_p), ("eps", C.c_float), ("flags", C.c_int32),
    ]


class DecodeAttentionArgs(C.Structure):
    _fields_ = [
        ("struct_size", C.c_size_t),
        ("Q", C.c_void_p), ("q_batch_stride", C.c_int64), ("K", C.c_void_p), ("V", C.c_void_p),
        ("kv_row_stride", C.c_int64), ("kv_batch_stride", C.c_int64),
        ("out", C.c_void_p), ("o_batch_stride", C.c_int64),
        ("B", C.c_int32), ("H", C.c_int32), ("Tk", C.c_int32), ("pos", C.c_void_p), ("kv_head_stride", C.c_int64),
        ("kv_batch_div", C.c_int32), ("ancestry", C.c_void_p), ("ancestry_stride", C.c_int64),
    ]


class LogitsRulesArgs(C.Structure):
    _fields_ = [
        ("struct_size", C.c_size_t),
        ("logits", C.c_void_p), ("ld", C.c_int64), ("B", C.c_int32), ("V", C.c_int32),
        ("ids", C.c_void_p), ("ids_row_stride", C.c_int64), ("pos", C.c_void_p), ("cur_len", C.c_int32),
        ("begin_index", C.c_int32), ("eos", C.c_int32), ("pad", C.c_int32), ("no_timestamps", C.c_int32),
        ("ts_begin", C.c_int32), ("max_initial_timestamp_index", C.c_int32), ("timestamp_rules", C.c_int32),
        ("suppress_bitmap", C.c_void_p), ("unfinished", C.c_void_p), ("processed_scores", C.c_void_p),
        ("no_select", C.c_int32),
    ]


class AugmentArgs(C.Structure):
    _fields_ = [
        ("struct_size", C.c_size_t),
        ("stno", C.c_void_p), ("B", C.c_int32), ("C", C.c_int32), ("Ts", C.c_int32),
        ("seg", C.c_void_p), ("seg_soft", C.c_void_p), ("n_seg", C.c_int32),
        ("noise_rows", C.c_void_p), ("noise", C.c_void_p), ("n_noise", C.c_int32),
        ("spec", C.c_int32), ("feats", C.c_void_p), ("feats_out", C.c_void_p), ("stno_out", C.c_void_p),
        ("M", C.c_int32), ("Tf", C.c_int32), ("factor", C.c_int32), ("center", C.c_int32), ("warped", C.c_int32),
        ("freq_masks", C.c_void_p), ("n_freq_masks", C.c_int32), ("time_masks", C.c_void_p), ("n_time_masks", C.c_int32),
        ("mask_channels", C.c_int32),
    ]


class CtcJointArgs(C.Structure):
    _fields_ = [
        ("struct_size", C.c_size_t),
        ("ids", C.c_void_p), ("ids_row_stride", C.c_int64), ("pos", C.c_void_p), ("cur_len", C.c_int32),
        ("B", C.c_int32), ("V", C.c_int32), ("T", C.c_int32), ("V1", C.c_int32), ("K", C.c_int32),
        ("bos", C.c_int32), ("eos", C.c_int32), ("pad", C.c_int32), ("blank", C.c_int32), ("first_timestamp", C.c_int32),
        ("prefix_len", C.c_int32), ("ctc_weight", C.c_float),
        ("ctc_logp", C.c_void_p), ("processed_scores", C.c_void_p), ("workspace_i32", C.c_void_p),
        ("workspace_f32", C.c_void_p), ("states", C.c_void_p), ("r_prev", C.c_void_p), ("score_prev", C.c_void_p),
        ("unfinished", C.c_void_p), ("raw_logits", C.c_void_p), ("score_only", C.c_int32),
    ]


class BeamStepArgs(C.Structure):
    _fields_ = [
        ("struct_size", C.c_size_t),
        ("U", C.c_int32), ("NB", C.c_int32), ("V", C.c_int32), ("K", C.c_int32),
        ("processed_scores", C.c_void_p), ("joint_workspace_i32", C.c_void_p), ("joint_workspace_f32", C.c_void_p),
        ("ctc_weight", C.c_float), ("ctc_states", C.c_void_p), ("ctc_r_prev", C.c_void_p), ("ctc_score_prev", C.c_void_p),
        ("ctc_r_tmp", C.c_void_p), ("T", C.c_int32),
        ("run_score", C.c_void_p), ("fin_score", C.c_void_p), ("fin_flag", C.c_void_p), ("unsat", C.c_void_p),
        ("ids", C.c_void_p), ("fin_ids", C.c_void_p), ("ids_tmp", C.c_void_p), ("ids_row_stride", C.c_int64),
        ("ancestry", C.c_void_p), ("ancestry_tmp", C.c_void_p), ("ancestry_stride", C.c_int64),
        ("pos", C.c_void_p), ("eos", C.c_int32), ("pad", C.c_int32), ("first_timestamp", C.c_int32),
        ("max_length", C.c_int32), ("prompt_len", C.c_int32), ("length_penalty", C.c_float), ("early_stopping", C.c_int32),
        ("scratch_i32", C.c_void_p), ("scratch_f32", C.c_void_p), ("flags", C.c_void_p),
    ]


class SoftlabelCeArgs(C.Structure):
    _fields_ = [
        ("struct_size", C.c_size_t),
        ("logits", C.c_void_p), ("ld", C.c_int64), ("rows", C.c_int32), ("V", C.c_int32),
        ("labels", C.c_void_p), ("upp_labels", C.c_void_p), ("ts_begin", C.c_int32), ("n_ts", C.c_int32),
        ("smoothing", C.c_void_p), ("soft_mode", C.c_int32), ("workspace", C.c_void_p), ("loss", C.c_void_p),
    ]


class AdamwTensorArgs(C.Structure):  # dicow_adamw_tensor_args_t: one entry of the device-resident tensor table
    _fields_ = [
        ("p", C.c_void_p), ("g", C.c_void_p), ("m", C.c_void_p), ("v", C.c_void_p), ("n", C.c_int64),
        ("lr", C.c_float), ("weight_decay", C.c_float), ("bias_correction1", C.c_float), ("bias_correction2_sqrt", C.c_float),
    ]


class CtcLossArgs(C.Structure):
    _fields_ = [
        ("struct_size", C.c_size_t),
        ("logits", C.c_void_p), ("B", C.c_int32), ("T", C.c_int32), ("V1", C.c_int32),
        ("labels", C.c_void_p), ("Lmax", C.c_int32), ("reduction_mean", C.c_int32),
        ("workspace", C.c_void_p), ("loss", C.c_void_p), ("ld", C.c_int64),
    ]


_lock = threading.Lock()
_lib = None
_handles: dict[int, C.c_void_p] = {}


ABI_VERSION = 2  # what the struct mirrors in this file describe (dicow_abi_version() of the loaded library must match)


def load_library() -> C.CDLL:
    """dlopen the shared library and declare prototypes (no GPU needed)."""
    global _lib
    if _lib is not None:  # fast path: no lock once loaded (called on every launch)
        return _lib
    with _lock:
        if _lib is not None:
            return _lib
        if not os.path.exists(LIB_PATH):
            raise DicowError(
                f"{LIB_PATH} is not built. Run `python __graft_entry__.py build` (nvcc, sm_100a). "
                "There is no fallback path.")
        lib = C.CDLL(LIB_PATH)
        _declare(lib)
        if lib.dicow_abi_version() != ABI_VERSION:
            raise DicowError(f"{LIB_PATH} has ABI version {lib.dicow_abi_version()}, the Python bindings expect {ABI_VERSION}: "
                             "rebuild with `python __graft_entry__.py build`")
        _lib = lib
        return lib


def _declare(lib: C.CDLL) -> None:
    vp = C.c_void_p
    lib.dicow_create.argtypes = [C.c_int, C.POINTER(vp)]
    lib.dicow_create.restype = C.c_int
    lib.dicow_destroy.argtypes = [vp]
    lib.dicow_destroy.restype = C.c_int
    lib.dicow_last_error.argtypes = [vp]
    lib.dicow_last_error.restype = C.c_char_p
    lib.dicow_check.argtypes = [vp]
    lib.dicow_set_sm_budget.argtypes = [vp, C.c_int]
    lib.dicow_adamw_chunk_elems.argtypes = []
    lib.dicow_adamw_step.argtypes = [vp, vp, vp, C.c_int, C.c_float, C.c_float, C.c_float, vp]
    lib.dicow_check.restype = C.c_int
    lib.dicow_abi_version.argtypes = []
    lib.dicow_abi_version.restype = C.c_int
    lib.dicow_gemm_bf16.argtypes = [vp, C.POINTER(GemmArgs), vp]
    lib.dicow_gemm_bf16.restype = C.c_int
    lib.dicow_fddt_layernorm.argtypes = [vp, C.POINTER(FddtLnArgs), vp]
    lib.dicow_attention_bf16.argtypes = [vp, C.POINTER(AttentionArgs), vp]
    lib.dicow_features_to_channels_last.argtypes = [vp, vp, vp, C.c_int, C.c_int, C.c_int, vp]
    lib.dicow_fddt_full_combine.argtypes = [vp, vp, C.c_int64, vp, C.c_int64, C.c_int, C.c_int, C.c_int, vp, vp, vp]
    lib.dicow_fddt_full_scatter.argtypes = [vp, vp, vp, C.c_int64, C.c_int, C.c_int, C.c_int, vp, C.c_int64, vp]
    lib.dicow_stno_mask.argtypes = [vp, vp, C.c_int64, C.c_int, C.c_int64, C.c_int, C.c_int, C.c_int64, vp, C.c_int64, C.c_int64, vp]
    lib.dicow_augment_batch.argtypes = [vp, C.POINTER(AugmentArgs), vp]
    lib.dicow_zero_pad_rows.argtypes = [vp, vp, C.c_int, C.c_int, C.c_int, vp]
    lib.dicow_cast_f32_bf16.argtypes = [vp, vp, vp, C.c_int64, vp]
    lib.dicow_debug_set_attention_profile.argtypes = [vp, vp]
    lib.dicow_logmel.argtypes = [vp, C.POINTER(LogmelArgs), vp]
    lib.dicow_attention_bwd_bf16.argtypes = [vp, C.POINTER(AttentionBwdArgs), vp]
    lib.dicow_layernorm_fddt_bwd.argtypes = [vp, C.POINTER(LnBwdArgs), vp]
    lib.dicow_colsum.argtypes = [vp, vp, C.c_int, C.c_int64, C.c_int, C.c_int, vp, C.c_float, vp]
    lib.dicow_conv1d_col2im.argtypes = [vp, vp, vp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int64, C.c_int64, vp]
    lib.dicow_ctc_loss_bwd.argtypes = [vp, C.POINTER(CtcBwdArgs), vp]
    lib.dicow_dgelu_mul.argtypes = [vp, vp, C.c_int, C.c_int64, vp, C.c_int64, vp, C.c_int64, C.c_int, C.c_int, vp]
    lib.dicow_gate_bwd.argtypes = [vp, vp, C.c_int64, vp, C.c_int64, vp, vp, C.c_int64, C.c_int, C.c_int, vp, vp]
    lib.dicow_cast_f32_bf16_2d.argtypes = [vp, vp, C.c_int64, vp, C.c_int64, C.c_int, C.c_int, C.c_int, vp]
    lib.dicow_embedding_bwd.argtypes = [vp, vp, vp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, vp, vp, vp]
    lib.dicow_softlabel_ce_bwd.argtypes = [vp, C.POINTER(SoftlabelCeBwdArgs), vp]
    lib.dicow_gemm_skinny_bf16.argtypes = [vp, C.POINTER(GemmSkinnyArgs), vp]
    lib.dicow_decode_linear.argtypes = [vp, C.POINTER(DecodeLinearArgs), vp]
    lib.dicow_decode_attention_bf16.argtypes = [vp, C.POINTER(DecodeAttentionArgs), vp]
    lib.dicow_kv_to_head_major.argtypes = [vp, vp, vp, C.c_int, C.c_int, C.c_int, vp]
    lib.dicow_embed_tokens.argtypes = [vp, vp, C.c_int64, vp, vp, vp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, vp, vp]
    lib.dicow_advance.argtypes = [vp, vp, C.c_int, vp]
    lib.dicow_logits_rules_argmax.argtypes = [vp, C.POINTER(LogitsRulesArgs), vp]
    lib.dicow_ctc_joint_step.argtypes = [vp, C.POINTER(CtcJointArgs), vp]
    lib.dicow_beam_step.argtypes = [vp, C.POINTER(BeamStepArgs), vp]
    lib.dicow_log_softmax_rows.argtypes = [vp, vp, vp, C.c_int64, C.c_int, vp]
    lib.dicow_softlabel_ce.argtypes = [vp, C.POINTER(SoftlabelCeArgs), vp]
    lib.dicow_ctc_loss.argtypes = [vp, C.POINTER(CtcLossArgs), vp]
    for name in EXPORTED_SYMBOLS:
        fn = getattr(lib, name)  # raises AttributeError if the library does not export a declared symbol
        if name != "dicow_last_error":
            fn.restype = C.c_int


# every entry point include/dicow_b200.h declares (tests/test_abi.py checks the header against this list)
EXPORTED_SYMBOLS = [
    "dicow_create", "dicow_destroy", "dicow_last_error", "dicow_check", "dicow_abi_version",
    "dicow_gemm_bf16", "dicow_fddt_layernorm", "dicow_attention_bf16", "dicow_features_to_channels_last",
    "dicow_zero_pad_rows", "dicow_cast_f32_bf16", "dicow_debug_set_attention_profile", "dicow_logmel",
    "dicow_gemm_skinny_bf16", "dicow_decode_attention_bf16", "dicow_embed_tokens", "dicow_advance",
    "dicow_logits_rules_argmax", "dicow_softlabel_ce", "dicow_ctc_loss", "dicow_attention_bwd_bf16",
    "dicow_layernorm_fddt_bwd", "dicow_colsum", "dicow_conv1d_col2im", "dicow_ctc_loss_bwd", "dicow_softlabel_ce_bwd",
    "dicow_dgelu_mul", "dicow_embedding_bwd", "dicow_cast_f32_bf16_2d", "dicow_gate_bwd", "dicow_decode_linear", "dicow_kv_to_head_major", "dicow_ctc_joint_step",
    "dicow_log_softmax_rows", "dicow_beam_step", "dicow_fddt_full_combine", "dicow_stno_mask", "dicow_augment_batch", "dicow_fddt_full_scatter",
    "dicow_decode_layers", "dicow_set_sm_budget", "dicow_adamw_step", "dicow_adamw_chunk_elems",
]


def handle(device: int) -> C.c_void_p:
    """Per-device library handle (created lazily)."""
    lib = load_library()
    h = _handles.get(device)
    if h is None:
        out = C.c_void_p()
        rc = lib.dicow_create(int(device), C.byref(out))
        if rc != 0:
            raise DicowError(
                f"dicow_create(device={device}) failed with status {rc} "
                "(2 = CUDA error / no device, 4 = not an sm_100 GPU). No fallback path exists.")
        _handles[device] = out
        h = out
    return h


def check(rc: int, h: C.c_void_p, what: str) -> None:
    if rc != 0:
        msg = load_library().dicow_last_error(h)
        raise DicowError(f"{what} failed (status {rc}): {msg.decode() if msg else ''}")
