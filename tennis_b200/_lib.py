"""ctypes binding of libtennis_b200.so (the C ABI declared in include/tennis_b200.h).

There is no CPU or PyTorch fallback anywhere below this module: if the shared object is missing, or a
call fails, a TennisB200Error is raised.
"""
import ctypes
import os
from ctypes import POINTER, c_char_p, c_float, c_int, c_int32, c_size_t, c_void_p

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libtennis_b200.so")

TN_OK, TN_ERR_INVALID, TN_ERR_CUDA, TN_ERR_ARCH, TN_ERR_WORKSPACE, TN_ERR_NCCL = 0, -1, -2, -3, -4, -5
ARCH_DENSENET121, ARCH_RESNET18_V2 = 0, 1
FRAMES_F32_NCHW, FRAMES_U8_NHWC = 0, 1
CELL_GRU, CELL_LSTM = 0, 1
POOL_MAX, POOL_MEAN = 0, 1
MODE_CONV, MODE_POOL2, MODE_STEM = 0, 1, 2


class TennisB200Error(RuntimeError):
    pass


_FLOATP = POINTER(c_float)
_FLOATPP = POINTER(_FLOATP)

# name -> (restype, argtypes); kept in the order of include/tennis_b200.h
SIGNATURES = {
    "tn_version": (c_int, []),
    "tn_last_error": (c_char_p, []),
    "tn_device_check": (c_int, [c_int]),
    "tn_profile_enable": (c_int, [c_int]),
    "tn_profile_read": (c_int, [POINTER(ctypes.c_double), POINTER(ctypes.c_longlong), POINTER(ctypes.c_double),
                                POINTER(ctypes.c_longlong), c_int]),
    "tn_backbone_param_count": (c_size_t, [c_int]),
    "tn_backbone_feature_dim": (c_int, [c_int, c_int, c_int]),
    "tn_backbone_create": (c_int, [POINTER(c_void_p), c_int, c_int, c_void_p, c_size_t]),
    "tn_backbone_destroy": (None, [c_void_p]),
    "tn_backbone_set_precision": (c_int, [c_void_p, c_int]),
    "tn_backbone_workspace_bytes": (c_size_t, [c_void_p, c_int, c_int, c_int]),
    "tn_backbone_forward": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p,
                                    c_size_t, c_void_p]),
    "tn_conv_create": (c_int, [POINTER(c_void_p), c_int, c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p,
                               c_void_p, c_void_p, c_void_p]),
    "tn_conv_destroy": (None, [c_void_p]),
    "tn_conv_forward": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_void_p,
                                c_int, c_int, c_int, c_void_p, c_int, c_void_p]),
    "tn_frames_to_nhwc4": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_void_p]),
    "tn_dense_forward": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_void_p]),
    "tn_temporal_pool": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p]),
    "tn_birnn_create": (c_int, [POINTER(c_void_p), c_int, c_int, c_int, c_int, c_int, POINTER(c_void_p),
                                POINTER(c_void_p), POINTER(c_void_p), POINTER(c_void_p)]),
    "tn_birnn_destroy": (None, [c_void_p]),
    "tn_birnn_set_precise": (c_int, [c_void_p, c_int]),
    "tn_birnn_update_weights": (c_int, [c_void_p, POINTER(c_void_p), POINTER(c_void_p), POINTER(c_void_p), POINTER(c_void_p),
                                        c_void_p]),
    "tn_birnn_workspace_bytes": (c_size_t, [c_void_p, c_int, c_int]),
    "tn_birnn_forward": (c_int, [c_void_p, c_void_p, c_int, c_void_p, c_int, c_int, c_void_p, c_void_p, c_void_p,
                                 c_void_p, c_void_p, c_size_t, c_void_p]),
}

_lib = None


def lib():
    """Load (once) and return the CDLL; raises if the extension has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise TennisB200Error(
                "%s not found: build it with `python -m tennis_b200._build` (or __graft_entry__.build()); "
                "tennis_b200 has no CPU fallback" % LIB_PATH)
        handle = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(handle, name)  # AttributeError here == header/library mismatch
            fn.restype = res
            fn.argtypes = args
        _lib = handle
    return _lib


def check(rc):
    if rc != TN_OK:
        msg = lib().tn_last_error()
        raise TennisB200Error("tennis_b200 error %d: %s" % (rc, msg.decode() if msg else "?"))


def stream_ptr(stream=None):
    import torch
    s = stream if stream is not None else torch.cuda.current_stream()
    return c_void_p(s.cuda_stream)


def dptr(t):
    """Device (or host) address of a torch tensor / None."""
    return c_void_p(0) if t is None else c_void_p(t.data_ptr())


def profile_enable(on=True):
    check(lib().tn_profile_enable(int(on)))


def profile_read(reset=True):
    """-> dict(conv_ms, conv_launches, other_ms, other_launches) accumulated since the last reset."""
    a, b = ctypes.c_double(), ctypes.c_double()
    na, nb = ctypes.c_longlong(), ctypes.c_longlong()
    check(lib().tn_profile_read(ctypes.byref(a), ctypes.byref(na), ctypes.byref(b), ctypes.byref(nb), int(reset)))
    return {"conv_ms": a.value, "conv_launches": na.value, "other_ms": b.value, "other_launches": nb.value}


SIGNATURES.update({
    "tn_gnmt_create": (c_int, [POINTER(c_void_p), c_int, c_int, c_int, c_int, c_int, c_int, c_int, POINTER(c_void_p),
                               POINTER(c_void_p), POINTER(c_void_p), POINTER(c_void_p), c_void_p, c_void_p, c_void_p,
                               c_void_p]),
    "tn_gnmt_destroy": (None, [c_void_p]),
    "tn_gnmt_workspace_bytes": (c_size_t, [c_void_p, c_int, c_int]),
    "tn_gnmt_decode_step": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int,
                                    c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]),
    "tn_metrics_update": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p]),
    "tn_gnmt_decoder_step": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int,
                                     c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]),
    "tn_gnmt_decode_seq": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int,
                                   c_int, c_void_p, c_void_p, c_size_t, c_void_p]),
    "tn_gnmt_beam_search": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int,
                                    c_float, c_float, c_int, c_int, c_void_p, c_void_p, c_void_p, POINTER(c_int),
                                    c_void_p, c_size_t, c_void_p]),
})


SIGNATURES.update({
    "tn_birnn_forward_train": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p,
                                       c_void_p, c_size_t, c_void_p]),
    "tn_birnn_backward_workspace_bytes": (c_size_t, [c_void_p, c_int, c_int]),
    "tn_birnn_backward": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                  c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]),
    "tn_softmax_ce": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_void_p]),
    "tn_dense_backward": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_void_p]),
    "tn_sgd_mom_update": (c_int, [c_void_p, c_void_p, c_void_p, c_size_t, c_float, c_float, c_float, c_float, c_void_p]),
    "tn_adam_update": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_size_t, c_float, c_float, c_float, c_float, c_float,
                               c_float, c_int, c_void_p]),
})

SIGNATURES["tn_masked_softmax_ce"] = (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_void_p])

_LL = ctypes.c_longlong
SIGNATURES.update({
    "tn_sgemm": (c_int, [c_int, c_int, c_int, c_int, c_int, c_float, c_void_p, c_int, c_void_p, c_int, c_float, c_void_p, c_int,
                         c_void_p]),
    "tn_rnn_cell_forward": (c_int, [c_int, c_int, c_int, c_void_p, _LL, c_void_p, _LL, c_void_p, c_void_p, c_void_p, _LL,
                                    c_void_p, _LL, c_void_p, _LL, c_void_p, _LL, c_void_p, _LL, c_void_p, _LL, c_void_p]),
    "tn_rnn_cell_backward": (c_int, [c_int, c_int, c_int, c_int, c_void_p, c_void_p, _LL, c_void_p, _LL, c_void_p, _LL, c_void_p,
                                     _LL, c_void_p, _LL, c_void_p, _LL, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, _LL,
                                     c_void_p, _LL, c_void_p]),
    "tn_rnn_unroll_forward": (c_int, [c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                      c_void_p, c_void_p, c_void_p]),
    "tn_rnn_unroll_backward": (c_int, [c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                       c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "tn_attention_forward": (c_int, [c_void_p, _LL, c_void_p, c_void_p, c_int, c_int, c_int, c_void_p, c_void_p, _LL, c_void_p,
                                     _LL, c_void_p]),
    "tn_attention_backward": (c_int, [c_void_p, _LL, c_void_p, c_void_p, c_int, c_int, c_int, c_void_p, c_void_p, _LL, c_void_p,
                                      _LL, c_void_p, _LL, c_void_p, c_void_p]),
    "tn_embedding_backward": (c_int, [c_void_p, c_void_p, _LL, c_void_p, c_int, c_int, c_int, c_void_p]),
    "tn_masked_softmax_ce_grad": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int,
                                          c_int, c_void_p]),
    "tn_dropout_mask": (c_int, [c_void_p, c_size_t, c_float, ctypes.c_ulonglong, c_void_p]),
    "tn_mul_mask": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_void_p]),
    "tn_axpy": (c_int, [c_void_p, c_void_p, c_float, c_size_t, c_void_p]),
})

SIGNATURES.update({
    "tn_gemm_tc_workspace_bytes": (_LL, [c_int, c_int, c_int]),
    "tn_split_bf16": (c_int, [c_void_p, _LL, _LL, c_int, c_int, c_int, c_int, c_int, c_int, c_void_p, c_void_p, _LL, c_void_p]),
    "tn_gemm_tc": (c_int, [c_int, c_int, c_int, c_int, c_void_p, c_int, c_void_p, c_void_p, _LL, _LL, _LL, c_void_p, c_void_p, _LL,
                           _LL, _LL, c_float, c_float, c_void_p, _LL, _LL, c_int, _LL, c_int, c_int, c_void_p, _LL, c_void_p]),
    "tn_im2col_nhwc": (c_int, [c_void_p, _LL, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_void_p, c_void_p]),
    "tn_col2im_nhwc": (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_void_p, _LL, c_void_p]),
    "tn_bn_train_forward": (c_int, [c_void_p, _LL, _LL, c_int, c_void_p, c_void_p, c_float, c_float, c_void_p, c_void_p, c_int,
                                    c_void_p, c_void_p, c_void_p, _LL, c_void_p]),
    "tn_bn_train_backward": (c_int, [c_void_p, _LL, c_void_p, _LL, c_void_p, _LL, _LL, c_int, c_void_p, c_void_p, c_void_p, c_float,
                                     c_int, c_void_p, c_void_p, c_void_p, _LL, c_int, c_void_p]),
    "tn_maxpool_nhwc_forward": (c_int, [c_void_p, _LL, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_void_p, _LL, c_void_p,
                                        c_void_p]),
    "tn_maxpool_nhwc_backward": (c_int, [c_void_p, _LL, c_void_p, _LL, c_int, c_void_p, _LL, c_void_p]),
    "tn_avgpool_nhwc_forward": (c_int, [c_void_p, _LL, c_int, c_int, c_int, c_int, c_int, c_int, c_void_p, _LL, c_void_p]),
    "tn_avgpool_nhwc_backward": (c_int, [c_void_p, _LL, c_int, c_int, c_int, c_int, c_int, c_int, c_void_p, _LL, c_int, c_void_p]),
})
