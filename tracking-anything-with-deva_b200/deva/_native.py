"""ctypes binding of libdeva_b200.so (C ABI declared in include/deva_b200.h).

This is the only door between the Python host code and the sm_100a kernels.  There is no
fallback: if the library is missing, was built for another ABI version, or a kernel reports an
error, a RuntimeError is raised.  PyTorch is used for device memory and streams only.
"""
import ctypes
import os
from ctypes import c_char_p, c_int, c_int32, c_int64, c_size_t, c_uint64, c_void_p, POINTER

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(os.path.dirname(_HERE), 'csrc', 'libdeva_b200.so')
ABI_VERSION = 1
LIST_PITCH = 32
MAX_GROUPS = 256

_lib = None

_SIGNATURES = {
    'deva_b200_abi_version': (c_int, []),
    'deva_b200_last_error': (c_char_p, []),
    'deva_b200_launch_count': (c_uint64, []),
    'deva_b200_device_check': (c_int, []),
    'deva_b200_pack_query': (c_int, [c_void_p, c_void_p, c_int64, c_int64, c_int, c_int, c_void_p, c_void_p,
                                     c_void_p, c_void_p]),
    'deva_b200_pack_keys': (c_int, [c_void_p, c_void_p, c_int64, c_int64, c_void_p, c_int, c_int, c_void_p, c_void_p,
                                    c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    'deva_b200_append_values': (c_int, [c_void_p, c_int64, c_void_p, c_int64, c_int, c_int, c_void_p]),
    'deva_b200_simtopk_workspace_bytes': (c_size_t, [c_int]),
    'deva_b200_sim_topk': (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_void_p, c_void_p, c_void_p, c_int,
                                   c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_void_p, c_void_p,
                                   c_int, c_int, c_int, c_void_p]),
    'deva_b200_sim_dense_softmax': (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_void_p, c_void_p,
                                            c_void_p, c_int, c_int, c_void_p, c_int64, c_void_p, c_int64, c_void_p,
                                            c_void_p]),
    'deva_b200_readout': (c_int, [c_void_p, c_int64, c_int64, POINTER(c_int32), POINTER(c_int32), c_int, c_int,
                                  c_void_p, c_int64, c_int, c_int, c_void_p, c_int64, c_void_p]),
    'deva_b200_gather_rows': (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_void_p]),
    'deva_b200_gather_f32': (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_void_p]),
    'deva_b200_gather_cols_f16': (c_int, [c_void_p, c_int64, c_void_p, c_int64, c_void_p, c_int, c_int, c_void_p]),
    'deva_b200_usage': (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_void_p]),
}
EXPORTS = tuple(_SIGNATURES.keys())


def lib():
    """Load (once) and return the shared library; raises if it is missing or mismatched."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(f'deva_b200: CUDA library not built: {LIB_PATH} (run __graft_entry__.build() or '
                               f'`make -C {os.path.dirname(LIB_PATH)}`); there is no CPU fallback')
        handle = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in _SIGNATURES.items():
            fn = getattr(handle, name)  # AttributeError if the symbol is not exported
            fn.restype = res
            fn.argtypes = args
        if handle.deva_b200_abi_version() != ABI_VERSION:
            raise RuntimeError('deva_b200: ABI version mismatch between deva/_native.py and libdeva_b200.so')
        _lib = handle
    return _lib


def _check(rc: int, what: str):
    if rc != 0:
        msg = lib().deva_b200_last_error()
        raise RuntimeError(f'deva_b200.{what} failed ({rc}): {msg.decode() if msg else "?"}')


def _ptr(t):
    return None if t is None else c_void_p(t.data_ptr())


def _stream():
    return c_void_p(torch.cuda.current_stream().cuda_stream)


def launch_count() -> int:
    return int(lib().deva_b200_launch_count())


def require_device():
    if not torch.cuda.is_available():
        raise RuntimeError('deva_b200: no CUDA device; the sm_100a kernels have no CPU fallback')
    _check(lib().deva_b200_device_check(), 'device_check')


def _f32(t):
    assert t.dtype == torch.float32 and t.is_cuda, (t.dtype, t.device)
    return t


# ------------------------------------------------------------------------------------------ wrappers
def pack_query(qk, qe, stride_c, stride_q, ck, q, q_hi, q_lo, bsq):
    _check(lib().deva_b200_pack_query(_ptr(_f32(qk)), _ptr(_f32(qe)), stride_c, stride_q, ck, q, _ptr(q_hi),
                                      _ptr(q_lo), _ptr(bsq), _stream()), 'pack_query')


def pack_keys(key, selection, stride_c, stride_t, shrinkage, ck, n, k_hi, k_lo, neg_s, raw_key, raw_sel, raw_shr):
    _check(lib().deva_b200_pack_keys(_ptr(_f32(key)), _ptr(selection), stride_c, stride_t, _ptr(_f32(shrinkage)), ck,
                                     n, _ptr(k_hi), _ptr(k_lo), _ptr(neg_s), _ptr(raw_key), _ptr(raw_sel),
                                     _ptr(raw_shr), _stream()), 'pack_keys')


def append_values(src, ld_src, dst, ld_dst, rows, n):
    _check(lib().deva_b200_append_values(_ptr(_f32(src)), ld_src, _ptr(dst), ld_dst, rows, n, _stream()),
           'append_values')


def simtopk_workspace_bytes(q):
    return int(lib().deva_b200_simtopk_workspace_bytes(q))


def sim_topk(k_hi, k_lo, neg_s, n_window, n_lead, q_hi, q_lo, bsq, q, ck, top_k, workspace, out_idx, out_w,
             affinity, ld_affinity, use_cnt, life_cnt, n_long, count_long, count_work):
    _check(lib().deva_b200_sim_topk(_ptr(k_hi), _ptr(k_lo), _ptr(neg_s), n_window, n_lead, _ptr(q_hi), _ptr(q_lo),
                                    _ptr(bsq), q, ck, top_k, _ptr(workspace), _ptr(out_idx), _ptr(out_w),
                                    _ptr(affinity), ld_affinity, _ptr(use_cnt), _ptr(life_cnt), n_long,
                                    int(count_long), int(count_work), _stream()), 'sim_topk')


def sim_dense_softmax(k_hi, k_lo, neg_s, shrinkage, n_window, n_lead, q_hi, q_lo, bsq, q, ck, sim_ws, ld_sim,
                      affinity, ld_affinity, shr_out):
    _check(lib().deva_b200_sim_dense_softmax(_ptr(k_hi), _ptr(k_lo), _ptr(neg_s), _ptr(shrinkage), n_window, n_lead,
                                             _ptr(q_hi), _ptr(q_lo), _ptr(bsq), q, ck, _ptr(sim_ws), ld_sim,
                                             _ptr(affinity), ld_affinity, _ptr(shr_out), _stream()),
           'sim_dense_softmax')


def readout(values, values_ld, values_rows, val_row, out_row, rows_per_group, affinity, ld_affinity, n_window, q,
            out, ld_out):
    n = len(val_row)
    assert n == len(out_row)
    arr_v = (c_int32 * n)(*val_row)
    arr_o = (c_int32 * n)(*out_row)
    _check(lib().deva_b200_readout(_ptr(values), values_ld, values_rows, arr_v, arr_o, n, rows_per_group,
                                   _ptr(affinity), ld_affinity, n_window, q, _ptr(out), ld_out, _stream()),
           'readout')


def gather_rows(dst, src, idx, n, row_bytes):
    _check(lib().deva_b200_gather_rows(_ptr(dst), _ptr(src), _ptr(idx), n, row_bytes, _stream()), 'gather_rows')


def gather_f32(dst, src, idx, n):
    _check(lib().deva_b200_gather_f32(_ptr(dst), _ptr(src), _ptr(idx), n, _stream()), 'gather_f32')


def gather_cols_f16(dst, ld_dst, src, ld_src, idx, rows, n):
    _check(lib().deva_b200_gather_cols_f16(_ptr(dst), ld_dst, _ptr(src), ld_src, _ptr(idx), rows, n, _stream()),
           'gather_cols_f16')


def usage(out, use_cnt, life_cnt, n):
    _check(lib().deva_b200_usage(_ptr(out), _ptr(use_cnt), _ptr(life_cnt), n, _stream()), 'usage')
