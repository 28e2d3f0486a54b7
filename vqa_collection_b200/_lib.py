"""ctypes binding of the C ABI in include/vqa_b200.h.

The library is the product: there is NO Python/CPU fallback.  If
``lib/libvqa_b200.so`` is missing, ``load()`` raises and tells the caller to
build it (``python -m vqa_collection_b200.build`` or ``__graft_entry__.build()``).
"""
import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "lib", "libvqa_b200.so")

VQA_F32, VQA_BF16, VQA_F16X2 = 0, 1, 2
ABI_VERSION = 11

c_void_p, c_int, c_float, c_size_t = C.c_void_p, C.c_int, C.c_float, C.c_size_t


class LinearArgs(C.Structure):
    _fields_ = [
        ("d_A", c_void_p), ("lda", c_int),
        ("d_W", c_void_p), ("ldw", c_int),
        ("M", c_int), ("N", c_int), ("K", c_int),
        ("dtype", c_int),
        ("d_scale", c_void_p), ("d_bias", c_void_p),
        ("relu", c_int),
        ("d_mul", c_void_p), ("ld_mul", c_int), ("mul_row_div", c_int),
        ("d_logit_w", c_void_p),
        ("d_out", c_void_p), ("ldo", c_int), ("out_dtype", c_int),
        ("d_add", c_void_p), ("ld_add", c_int), ("add_row_div", c_int),
        ("trans_a", c_int), ("trans_w", c_int),
        ("d_mask", c_void_p), ("ld_mask", c_int), ("mask_dtype", c_int),
        ("leaky_slope", c_float), ("add_after_act", c_int), ("sigmoid", c_int),
        ("d_argmax_label", c_void_p), ("d_argmax_ws", c_void_p),
        ("tile_begin", c_int), ("tile_end", c_int), ("cta_limit", c_int),
        ("d_progress", c_void_p),
        ("a_plane", c_size_t), ("w_plane", c_size_t), ("out_plane", c_size_t),
    ]


class GruArgs(C.Structure):
    _fields_ = [
        ("d_tokens", c_void_p),
        ("B", c_int), ("T", c_int), ("H", c_int), ("E_pad", c_int), ("ntoken_rows", c_int),
        ("dtype", c_int),
        ("d_emb", c_void_p),
        ("d_w_ih", c_void_p), ("d_b_ih", c_void_p),
        ("d_w_hh", c_void_p), ("d_b_hh", c_void_p),
        ("d_wx_packed", c_void_p), ("d_wh_packed", c_void_p), ("d_bias_packed", c_void_p),
        ("d_workspace", c_void_p), ("workspace_bytes", c_size_t),
        ("d_h_last", c_void_p), ("d_h_last_lp", c_void_p),
        ("d_x", c_void_p), ("d_out_all", c_void_p),
        ("d_gi_table", c_void_p),
    ]


class GraphAttentionArgs(C.Structure):
    _fields_ = [
        ("d_Y", c_void_p), ("ldy", c_int),
        ("d_att", c_void_p),
        ("d_labels", c_void_p),
        ("d_label_bias", c_void_p), ("num_labels", c_int),
        ("d_ba", c_void_p), ("d_bb", c_void_p),
        ("B", c_int), ("K", c_int), ("V", c_int), ("dtype", c_int),
        ("d_out", c_void_p), ("d_vsum", c_void_p), ("d_alpha", c_void_p),
        ("layout", c_int), ("d_x", c_void_p), ("ldx", c_int), ("d_wvec", c_void_p), ("c0", c_float),
        ("d_label_bias_lp", c_void_p),
        ("d_progress", c_void_p), ("progress_target", c_int), ("cta_limit", c_int),
    ]


class ForwardArgs(C.Structure):
    _fields_ = [
        ("B", c_int), ("K", c_int), ("V", c_int), ("H", c_int), ("A", c_int), ("T", c_int),
        ("E_pad", c_int), ("ntoken_rows", c_int), ("num_labels", c_int),
        ("dtype", c_int), ("relation", c_int),
        ("d_img", c_void_p), ("d_tokens", c_void_p), ("d_labels", c_void_p), ("d_bbox", c_void_p),
        ("img_w", c_float), ("img_h", c_float),
        ("d_emb", c_void_p), ("d_w_ih", c_void_p), ("d_b_ih", c_void_p),
        ("d_w_hh", c_void_p), ("d_b_hh", c_void_p),
        ("d_wx_packed", c_void_p), ("d_wh_packed", c_void_p), ("d_bias_packed", c_void_p),
        ("d_Wv", c_void_p), ("d_sv", c_void_p), ("d_bv", c_void_p),
        ("d_Wqq", c_void_p), ("d_sqq", c_void_p), ("d_bqq", c_void_p),
        ("d_wlin", c_void_p), ("b_lin", c_float),
        ("att_concat", c_int), ("d_W1q", c_void_p), ("d_b1", c_void_p),
        ("d_Wg", c_void_p), ("d_label_bias", c_void_p), ("d_ba", c_void_p), ("d_bb", c_void_p),
        ("d_Wg3", c_void_p), ("d_wvec", c_void_p), ("gat_c0", c_float), ("d_label_bias_lp", c_void_p),
        ("d_Wvn", c_void_p), ("d_svn", c_void_p), ("d_bvn", c_void_p),
        ("d_Wc0", c_void_p), ("d_sc0", c_void_p), ("d_bc0", c_void_p),
        ("d_Wc1", c_void_p), ("d_sc1", c_void_p), ("d_bc1", c_void_p),
        ("d_workspace", c_void_p), ("workspace_bytes", c_size_t),
        ("d_logits", c_void_p), ("d_label", c_void_p), ("d_att", c_void_p), ("d_q", c_void_p),
        ("d_v", c_void_p), ("d_alpha", c_void_p), ("d_labels_out", c_void_p),
        ("overlap", c_int), ("side_sms", c_int), ("side_tile_permille", c_int), ("gat_chase_sms", c_int),
        ("d_gi_table", c_void_p),
    ]


class ForwardHostArgs(C.Structure):
    _fields_ = [
        ("fwd", ForwardArgs),
        ("h_img", c_void_p), ("h_tokens", c_void_p), ("h_labels", c_void_p), ("h_bbox", c_void_p),
        ("h_label", c_void_p),
        ("chunk_rows", c_int), ("pack_on_host", c_int), ("raw_chunk_period", c_int),
        ("h2d_bytes", c_size_t), ("d2h_bytes", c_size_t),
        ("img_is_bf16", c_int),
    ]


TRAIN_LAYERS = 7
_fp = C.c_void_p * TRAIN_LAYERS


class TrainArgs(C.Structure):
    _fields_ = [
        ("B", c_int), ("K", c_int), ("V", c_int), ("H", c_int), ("A", c_int), ("T", c_int), ("E", c_int),
        ("ntoken_rows", c_int), ("dtype", c_int),
        ("dropout_att", c_float), ("dropout_cls", c_float), ("seed", C.c_ulonglong),
        ("d_img", c_void_p), ("d_tokens", c_void_p), ("d_target", c_void_p),
        ("p_emb", c_void_p), ("p_w_ih", c_void_p), ("p_w_hh", c_void_p), ("p_b_ih", c_void_p), ("p_b_hh", c_void_p),
        ("p_v", _fp), ("p_g", _fp), ("p_b", _fp),
        ("g_emb", c_void_p), ("g_w_ih", c_void_p), ("g_w_hh", c_void_p), ("g_b_ih", c_void_p), ("g_b_hh", c_void_p),
        ("g_v", _fp), ("g_g", _fp), ("g_b", _fp),
        ("d_loss", c_void_p), ("d_logits", c_void_p),
        ("d_workspace", c_void_p), ("workspace_bytes", c_size_t),
        ("ev_head_done", c_void_p),
    ]


class LstmArgs(C.Structure):
    _fields_ = [
        ("B", c_int), ("T", c_int), ("H", c_int), ("E_pad", c_int), ("dtype", c_int),
        ("d_x", c_void_p), ("d_w_ih", c_void_p), ("d_b_ih", c_void_p), ("d_w_hh", c_void_p), ("d_b_hh", c_void_p),
        ("d_workspace", c_void_p), ("workspace_bytes", c_size_t),
        ("d_h_last", c_void_p), ("d_out_all", c_void_p),
    ]


OPTIM_MAX_TENSORS = 64


class OptimTensor(C.Structure):
    _fields_ = [("d_p", c_void_p), ("d_g", c_void_p), ("d_m", c_void_p), ("d_u", c_void_p), ("n", c_size_t), ("lr", c_float)]


class CaptionDecodeArgs(C.Structure):
    _fields_ = [
        ("B", c_int), ("K", c_int), ("V", c_int), ("Hd", c_int), ("T", c_int), ("dtype", c_int),
        ("h_batches", C.POINTER(c_int)),
        ("d_x", c_void_p), ("d_proj", c_void_p),
        ("att_mode", c_int),
        ("d_wq", c_void_p), ("d_wq_scale", c_void_p), ("d_wq_bias", c_void_p),
        ("d_logit_w", c_void_p), ("logit_bias", c_float),
        ("d_gi_prev", c_void_p),
        ("d_w_att", c_void_p), ("d_w_hh", c_void_p), ("d_b_hh", c_void_p),
        ("d_h_all", c_void_p), ("d_h", c_void_p), ("d_h0_lp", c_void_p),
        ("d_workspace", c_void_p), ("workspace_bytes", c_size_t),
        ("cell", c_int), ("d_c", c_void_p),
    ]


# every symbol include/vqa_b200.h declares: name -> (restype, argtypes)
SYMBOLS = {
    "vqa_abi_version": (c_int, []),
    "vqa_last_error": (C.c_char_p, []),
    "vqa_device_info": (c_int, [C.POINTER(c_int)] * 3),
    "vqa_relation_labels": (c_int, [c_void_p, c_void_p, c_int, c_int, c_float, c_float, c_void_p, c_void_p]),
    "vqa_relation_labels_host": (c_int, [c_void_p, c_int, c_int, c_float, c_float, c_void_p]),
    "vqa_relation_near_threshold": (c_float, [c_float, c_float]),
    "vqa_cast_f32_to_bf16": (c_int, [c_void_p, c_void_p, c_size_t, c_void_p]),
    "vqa_split_f32": (c_int, [c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]),
    "vqa_cast_bf16_to_f32": (c_int, [c_void_p, c_void_p, c_size_t, c_void_p]),
    "vqa_linear": (c_int, [C.POINTER(LinearArgs), c_void_p]),
    "vqa_linear_part_width": (c_int, [c_int]),
    "vqa_linear_tile_count": (c_int, [C.POINTER(LinearArgs)]),
    "vqa_linear_tiles_n": (c_int, [C.POINTER(LinearArgs)]),
    "vqa_linear_argmax_workspace_bytes": (c_size_t, [c_int]),
    "vqa_gru_last_state": (c_int, [C.POINTER(GruArgs), c_void_p]),
    "vqa_gru_workspace_bytes": (c_size_t, [c_int, c_int, c_int, c_int, c_int]),
    "vqa_lstm_sequence": (c_int, [C.POINTER(LstmArgs), c_void_p]),
    "vqa_lstm_workspace_bytes": (c_size_t, [c_int, c_int, c_int, c_int]),
    "vqa_attention_pool": (c_int, [c_void_p, c_int, c_float, c_void_p, c_int, c_int, c_int, c_int,
                                   c_void_p, c_void_p, c_void_p, c_void_p]),
    "vqa_graph_attention": (c_int, [C.POINTER(GraphAttentionArgs), c_void_p]),
    "vqa_add_inplace": (c_int, [c_void_p, c_void_p, c_size_t, c_int, c_void_p]),
    "vqa_answer_scores": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p]),
    "vqa_caption_gate_scale": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_void_p,
                                       c_void_p]),
    "vqa_seq_max": (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_void_p]),
    "vqa_softmax_mul": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_void_p, c_void_p]),
    "vqa_attention_logits": (c_int, [c_void_p, c_int, c_void_p, c_int, c_void_p, c_int, c_int, c_int, c_int, c_int,
                                     c_void_p, c_void_p]),
    "vqa_lstm_cell": (c_int, [c_void_p, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p, c_int, c_void_p]),
    "vqa_gru_cell": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_void_p, c_void_p, c_int, c_void_p]),
    "vqa_caption_decode_workspace_bytes": (c_size_t, [c_int, c_int, c_int, c_int, c_int]),
    "vqa_caption_decode_steps": (c_int, [C.POINTER(CaptionDecodeArgs), c_void_p]),
    "vqa_grad_clip_workspace_bytes": (c_size_t, []),
    "vqa_grad_clip": (c_int, [C.POINTER(OptimTensor), c_int, c_float, c_int, c_void_p, c_void_p, c_void_p, c_void_p]),
    "vqa_adamax_step": (c_int, [C.POINTER(OptimTensor), c_int, c_float, c_float, c_float, c_float, c_int, c_void_p,
                                c_void_p]),
    "vqa_argmax_rows": (c_int, [c_void_p, c_int, c_int, c_int, c_void_p, c_void_p]),
    "vqa_forward_workspace_bytes": (c_size_t, [C.POINTER(ForwardArgs)]),
    "vqa_forward": (c_int, [C.POINTER(ForwardArgs), c_void_p]),
    "vqa_forward_last_launch_count": (c_int, []),
    "vqa_host_ctx_create": (c_int, [C.POINTER(c_void_p), c_int]),
    "vqa_host_ctx_destroy": (None, [c_void_p]),
    "vqa_host_ctx_threads": (c_int, [c_void_p]),
    "vqa_forward_host": (c_int, [c_void_p, C.POINTER(ForwardHostArgs), c_void_p]),
    "vqa_forward_host_submit": (c_int, [c_void_p, C.POINTER(ForwardHostArgs), c_void_p]),
    "vqa_forward_host_wait": (c_int, [c_void_p]),
    "vqa_train_workspace_bytes": (c_size_t, [C.POINTER(TrainArgs)]),
    "vqa_updown_train_step": (c_int, [C.POINTER(TrainArgs), c_void_p]),
}

_lib = None


def load():
    """dlopen the C-ABI library and bind every declared symbol (raises if missing)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"vqa_collection_b200: CUDA library {LIB_PATH} is not built and there is no CPU fallback; "
            "run `python -m vqa_collection_b200.build` (or __graft_entry__.build()) first")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SYMBOLS.items():
        fn = getattr(lib, name)          # AttributeError if the .so does not export it
        fn.restype = res
        fn.argtypes = args
    if lib.vqa_abi_version() != ABI_VERSION:
        raise RuntimeError(f"vqa_collection_b200: ABI version mismatch ({lib.vqa_abi_version()} != {ABI_VERSION})")
    _lib = lib
    return lib


class VqaError(RuntimeError):
    pass


def check(rc):
    if rc != 0:
        msg = load().vqa_last_error().decode("utf-8", "replace")
        raise VqaError(f"vqa_b200 error {rc}: {msg}")
