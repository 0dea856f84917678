"""Host side of the tensor-core training GEMM (csrc/tn_gemm_tc.cu): operand planes and the contraction call.

The convolutions of the trainable CNN (reference train.py:415-421: `ag.record()` / `ag.backward` through gluoncv's DenseNet-121 /
ResNet-18 v2, fp32 MXNet Convolution) run their forward, data-gradient and weight-gradient contractions through these two calls.
Torch tensors are containers only; nothing here computes.
"""
import ctypes
import os

import torch

from ._lib import check, dptr, lib, stream_ptr


def mode():
    """$TN_TRAIN_GEMM: 'x3' (default: split-bf16, three tensor-core products, fp32-grade), 'bf16' (one product) or 'fp32' (the
    SIMT SGEMM of round 1, kept as the parity anchor)."""
    m = os.environ.get("TN_TRAIN_GEMM", "x3")
    if m not in ("x3", "bf16", "fp32"):
        raise ValueError("TN_TRAIN_GEMM must be x3, bf16 or fp32, got %r" % m)
    return m


def _r8(n):
    return (n + 7) // 8 * 8


class Planes(object):
    """K-major bf16 operand: `rows` rows of `kdim` contraction elements (row stride ld), as hi (+ lo) planes."""
    __slots__ = ("hi", "lo", "rows", "kdim", "ld")

    def __init__(self, hi, lo, rows, kdim, ld):
        self.hi, self.lo, self.rows, self.kdim, self.ld = hi, lo, rows, kdim, ld


def padded_rows(n_frames, h, w, pitch=None):
    return n_frames * (h + 2) * (pitch or (w + 2))


def pitch8(w):
    """Row pitch of the padded grid for planes whose CONTRACTION runs over the pixels (weight gradient): a multiple of 8 pixels,
    so that a dy tap shift is a 16-byte aligned TMA coordinate."""
    return _r8(w + 2)


# ---- tap lists of a 3x3 / stride 1 / pad 1 convolution in zero-padded row space (consumed by `gemm`; the index algebra is
# checked on the CPU against torch.conv2d and its autograd in tests/test_host_cpu.py::test_tap_lists_reproduce_conv3x3)
def conv3x3_offsets(w):
    """Row shift of tap t = 3*r + s in the padded grid of pitch w + 2."""
    return [(r - 1) * (w + 2) + (s - 1) for r in range(3) for s in range(3)]


def taps_conv3x3_forward(w, cin):
    """Y[m, n] = sum_t sum_c X[m + off_t, c] * Wk[n, t*cin + c]   (A = padded X planes, B = weights (Cout, 9*cin))."""
    return [(o, 0, 0, t * cin) for t, o in enumerate(conv3x3_offsets(w))]


def taps_conv3x3_dgrad(w, cout):
    """dX[m, c] = sum_t sum_n dY[m - off_t, n] * Wt[c, t*cout + n]   (A = padded dY planes, B = (Cin, 9*cout))."""
    return [(-o, 0, 0, t * cout) for t, o in enumerate(conv3x3_offsets(w))]


def taps_conv3x3_wgrad(w, cout):
    """Nine independent products (tile_taps): D_t[c, n] = sum_j X^T[c, j] * dYs_dx^T[n, j - dy*P8] with t = (dy+1)*3 + (dx+1),
    A = X^T over the padded pixels at pitch P8 = pitch8(w), B = the three dx-shifted copies of dY^T stacked by rows ((dx+1)*cout + n):
    the dx shift lives in the copy (`planes(..., shift=dx)`), the dy shift is the 16-byte aligned contraction offset -dy*P8."""
    p8 = pitch8(w)
    return [(0, 0, (t % 3) * cout, -(t // 3 - 1) * p8) for t in range(9)]


def planes(src, transpose=False, pad_hw=None, lo=True, pitch=None, shift=0, into=None, row0=0):
    """src: 2-D fp32 view (rows, cols) with unit column stride.  transpose=False: the contraction runs along the columns;
    True: along the rows.  pad_hw=(H, W): rows are (n, y, x) pixels, re-indexed into the zero-padded grid of (H+2) rows of `pitch`
    (default W+2) pixels; shift (-1/0/+1, transposed padded planes): every pixel lands `shift` positions later.
    into / row0: write into rows [row0, row0 + rows) of an existing Planes of the same contraction length."""
    assert src.dim() == 2 and src.dtype == torch.float32 and src.is_cuda and (src.shape[1] == 1 or src.stride(1) == 1), \
        (src.shape, src.stride(), src.dtype)
    R, C = src.shape
    if pad_hw is not None:
        h, w = pad_hw
        assert R % (h * w) == 0, (R, h, w)
        Rp = padded_rows(R // (h * w), h, w, pitch)
    else:
        h = w = 0
        Rp = R
    rows, kdim = (C, Rp) if transpose else (Rp, C)
    if into is not None:
        assert into.kdim == kdim and row0 + rows <= into.rows and (into.lo is not None) == bool(lo)
        hi, lo_t, ld = into.hi[row0:], (into.lo[row0:] if lo else None), into.ld
    else:
        ld = _r8(kdim)
        hi = torch.empty((rows, ld), dtype=torch.bfloat16, device=src.device)
        lo_t = torch.empty((rows, ld), dtype=torch.bfloat16, device=src.device) if lo else None
    check(lib().tn_split_bf16(dptr(src), src.stride(0), R, C, int(transpose), h, w, int(pitch or 0), int(shift), dptr(hi),
                              dptr(lo_t) if lo else None, ld, stream_ptr()))
    return into if into is not None else Planes(hi, lo_t, rows, kdim, ld)


def empty_planes(rows, kdim, device, lo=True):
    ld = _r8(kdim)
    return Planes(torch.empty((rows, ld), dtype=torch.bfloat16, device=device),
                  torch.empty((rows, ld), dtype=torch.bfloat16, device=device) if lo else None, rows, kdim, ld)


_WS = {}


def _workspace(device, nbytes):
    ws = _WS.get(device)
    if ws is None or ws.numel() < nbytes:
        ws = torch.empty(max(nbytes, 1 << 20), dtype=torch.uint8, device=device)
        _WS[device] = ws
    return ws


def gemm(A, B, M, N, K, C, c_row_stride, c_col_stride=1, taps=None, alpha=1.0, beta=0.0, unpad_hw=None, passes=3,
         tile_taps=False, c_tap_stride=0):
    """C[orow(m)*c_row_stride + n*c_col_stride] = alpha * sum_t sum_{k<K} A[m + ar_t, ak_t + k] * B[n + br_t, bk_t + k] + beta * C.
    C: the fp32 tensor (or view) whose data pointer is element (0, 0) of the output; taps: list of (ar, ak, br, bk).
    tile_taps: the taps are independent products of one launch, tap t written at C + t * c_tap_stride."""
    if passes == 3:
        assert A.lo is not None and B.lo is not None
    nt = 1 if taps is None else len(taps)
    tap_arr = None
    if taps is not None:
        flat = [int(v) for t in taps for v in t]
        tap_arr = (ctypes.c_int * len(flat))(*flat)
    nbytes = int(lib().tn_gemm_tc_workspace_bytes(M, N, nt if tile_taps else 0))
    ws = _workspace(C.device, nbytes)
    uh, uw = unpad_hw if unpad_hw is not None else (0, 0)
    check(lib().tn_gemm_tc(M, N, K, nt, tap_arr, passes, dptr(A.hi), dptr(A.lo) if A.lo is not None else None, A.rows, A.kdim, A.ld,
                           dptr(B.hi), dptr(B.lo) if B.lo is not None else None, B.rows, B.kdim, B.ld, alpha, beta, dptr(C),
                           c_row_stride, c_col_stride, int(tile_taps), c_tap_stride, uh, uw, dptr(ws), ws.numel(), stream_ptr()))
    return C


def matmul(A2, B2, C, ta=False, tb=False, alpha=1.0, beta=0.0, passes=3):
    """Drop-in for the fp32 `sgemm(A, B, C, ta, tb, alpha, beta)` of models/captioning/train_graph.py on the tensor cores:
    C (M,N) = alpha * op(A) op(B) + beta * C on 2-D row-major fp32 views."""
    lo = passes == 3
    M, K = (A2.shape[1], A2.shape[0]) if ta else (A2.shape[0], A2.shape[1])
    N = B2.shape[0] if tb else B2.shape[1]
    Ap = planes(A2, transpose=ta, lo=lo)          # (M, K) K-major
    Bp = planes(B2, transpose=not tb, lo=lo)      # (N, K) K-major
    return gemm(Ap, Bp, M, N, K, C, C.stride(0), 1, alpha=alpha, beta=beta, passes=passes)
