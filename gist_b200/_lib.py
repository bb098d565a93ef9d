"""ctypes binding of the C-ABI shared library (include/gist_b200.h).

The library is built in-tree (``gist_b200/csrc/libgist_b200.so``) by
``__graft_entry__.build()`` / ``make -C gist_b200/csrc``.  There is NO fallback:
if the library is missing, or a compute entry point is called on a non-CUDA
tensor, a RuntimeError is raised.
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# GIST_B200_LIB: an alternative build of the same library (A/B measurements of compile-time knobs)
LIB_PATH = os.environ.get('GIST_B200_LIB') or os.path.join(_HERE, 'csrc', 'libgist_b200.so')

_c_i32p = ctypes.c_void_p
_P = ctypes.c_void_p
_I32 = ctypes.c_int32
_I64 = ctypes.c_int64
_U32 = ctypes.c_uint32
_SZ = ctypes.c_size_t
_F32 = ctypes.c_float

# name -> (restype, argtypes); must list every symbol include/gist_b200.h declares
SIGNATURES = {
    'gist_abi_version': (ctypes.c_int, []),
    'gist_status_string': (ctypes.c_char_p, [ctypes.c_int]),
    'gist_set_device': (ctypes.c_int, [ctypes.c_int]),
    'gist_launch_count': (ctypes.c_uint64, []),
    'gist_spmm_csr_f32': (ctypes.c_int, [_P, _P, _I32, _I32, _P, _I64, _I32, _P, _I64,
                                         _P, _P, _P, _P, _I64, _P, _I64, _U32, _P]),
    'gist_spmm_csc_f32': (ctypes.c_int, [_P, _P, _I32, _I32, _P, _I64, _I32, _P, _I64,
                                         _P, _P, _P, _I64, _U32, _P]),
    'gist_degree_norm_f32': (ctypes.c_int, [_P, _I32, _I32, _P, _P]),
    'gist_scan_workspace_bytes': (_SZ, [_I32]),
    'gist_exclusive_scan_i32': (ctypes.c_int, [_P, _I32, _P, _P, _SZ, _P]),
    'gist_cluster_batch_build': (ctypes.c_int, [_P, _P, _I32, _P, _I32, _P, _P, _P, _I64, _P, _P,
                                                _P, _SZ, _P]),
    'gist_cluster_batch_build_v2_workspace_bytes': (_SZ, [_I32, _I64]),
    'gist_cluster_batch_build_v2': (ctypes.c_int, [_P, _P, _I32, _P, _I32, _P, _P, _P, _I64, _P, _P, _I64,
                                                   _P, _SZ, _P]),
    'gist_gather_rows': (ctypes.c_int, [_P, _I64, _P, _I64, _P, _I64, _I64, _P]),
    'gist_slice_gather_f32': (ctypes.c_int, [_P, _I64, _P, _I64, _P, _I64, _P, _I64, _P]),
    'gist_slice_scatter_f32': (ctypes.c_int, [_P, _I64, _P, _I64, _P, _I64, _P, _I64, _P]),
    'gist_slice_multi_f32': (ctypes.c_int, [_I32, _I32, _P, _P]),
    'gist_slice_scatter_rows_f32': (ctypes.c_int, [_I32, _P, _P]),
    'gist_index_invert_i32': (ctypes.c_int, [_P, _I64, _P, _I64, _P]),
    'gist_gemm_tf32_workspace_bytes': (_SZ, [_I32, _I32, _I32, _U32]),
    'gist_gemm_tf32': (ctypes.c_int, [_P, _I64, _I32, _P, _I64, _I32, _P, _I64, _I32, _I32, _I32, _P, _U32,
                                      _P, _SZ, _P]),
    'gist_gemm_3xtf32_workspace_bytes': (_SZ, [_I32, _I32, _I32, _U32]),
    'gist_gemm_3xtf32': (ctypes.c_int, [_P, _P, _I64, _I64, _I32, _P, _P, _I64, _I64, _I32, _P, _I64, _I32, _I32,
                                        _I32, _P, _U32, _P, _SZ, _P]),
    'gist_split_tf32_f32': (ctypes.c_int, [_P, _I64, _I32, _I32, _P, _I64, _P, _I64, _P]),
    'gist_split_tf32_multi_f32': (ctypes.c_int, [_I32, _P, _P, _P, _P, _P, _P, _P]),
    'gist_gemm_set_trace': (ctypes.c_int, [_P, _I32]),
    'gist_gemm_trace_slots': (ctypes.c_int, []),
    'gist_gemm_plan': (ctypes.c_int, [_I32, _I32, _I32, _U32, _I32, _P, _P, _P]),
    'gist_gemm_tn_tf32': (ctypes.c_int, [_P, _I64, _P, _I64, _P, _I64, _I32, _I32, _I32, _P, _U32, _P]),
    'gist_transpose_f32': (ctypes.c_int, [_P, _I64, _I32, _I32, _P, _I64, _P]),
    'gist_layernorm_act_fwd_f32': (ctypes.c_int, [_P, _I64, _I32, _I32, _F32, _U32, _P, _I64, _P, _P]),
    'gist_layernorm_act_bwd_f32': (ctypes.c_int, [_P, _I64, _P, _I64, _P, _I32, _I32, _U32, _P, _I64, _P, _I64, _P]),
    'gist_tensor_layernorm_workspace_bytes': (_SZ, [_I32, _I32]),
    'gist_tensor_layernorm_fwd_f32': (ctypes.c_int, [_P, _I64, _I32, _I32, _F32, _P, _I64, _P, _P, _SZ, _P]),
    'gist_tensor_layernorm_bwd_f32': (ctypes.c_int, [_P, _I64, _P, _I64, _P, _I32, _I32, _P, _I64, _P, _SZ, _P]),
    'gist_colsum_workspace_bytes': (_SZ, [_I32, _I32]),
    'gist_colsum_f32': (ctypes.c_int, [_P, _I64, _I32, _I32, _P, _P, _SZ, _P]),
    'gist_masked_ce_fwd_f32': (ctypes.c_int, [_P, _I64, _I32, _I32, _P, _P, _P, _P, _P, _P]),
    'gist_masked_ce_bwd_f32': (ctypes.c_int, [_P, _I64, _I32, _I32, _P, _P, _P, _P, _P, _P, _I64, _I32, _P, _P]),
    'gist_masked_ce_fused_workspace_bytes': (_SZ, [_I32]),
    'gist_masked_ce_fused_f32': (ctypes.c_int, [_P, _I64, _I32, _I32, _P, _P, _P, _I64, _I32, _P, _P, _P, _SZ, _P, _P]),
    'gist_dropout_f32': (ctypes.c_int, [_P, _I64, _I32, _I32, _I32, _P, _I64, _P, _P]),
    'gist_counter_add_i64': (ctypes.c_int, [_P, _I64, _P]),
    'gist_spmm_csr_ex_f32': (ctypes.c_int, [_P, _P, _I32, _I32, _P, _I64, _I32, _P, _I64,
                                            _P, _P, _P, _P, _I64, _P, _I64, _U32, _P, _P]),
    'gist_spmm_schedule_workspace_bytes': (_SZ, [_I32]),
    'gist_spmm_schedule_build': (ctypes.c_int, [_P, _I32, _I32, _P, _P, _I64, _P, _SZ, _P]),
    'gist_spmm_schedule_build_meta': (ctypes.c_int, [_P, _I32, _I32, _P, _P, _P, _I64, _P, _SZ, _P]),
    'gist_gemm_has_inkernel_splitk': (ctypes.c_int, []),
    'gist_gat_scores_heads_f32': (ctypes.c_int, [_P, _I64, _I32, _I32, _I32, _P, _P, _P]),
    'gist_gat_aggregate_heads_f32': (ctypes.c_int, [_P, _P, _I32, _P, _I64, _I32, _I32, _P, _F32, _P, _I64, _P, _P]),
    'gist_gat_backward_heads_workspace_bytes': (_SZ, [_I32, _I32, _I32]),
    'gist_gat_backward_heads_f32': (ctypes.c_int, [_P, _P, _P, _P, _I32, _P, _I64, _I32, _I32, _P, _P, _P, _F32, _P, _I64,
                                                   _P, _I64, _P, _I64, _P, _P, _SZ, _P]),
    'gist_gemm_ex_workspace_bytes': (_SZ, [_I32, _I32, _I32, _U32, _I32, _P]),
    'gist_gemm_ex_f32': (ctypes.c_int, [_P, _P, _I64, _I64, _I32, _P, _P, _I64, _I64, _I32, _P, _I64, _I32, _I32,
                                        _I32, _P, _U32, _P, _SZ, _P, _P]),
    'gist_gemm_dropmask_f32': (ctypes.c_int, [_P, _P, _I64, _I64, _I32, _P, _P, _I64, _I64, _I32, _P, _I64,
                                              _I32, _I32, _I32, _U32, _P, _P]),
    'gist_gat_scores_f32': (ctypes.c_int, [_P, _I64, _I32, _I32, _P, _P, _P]),
    'gist_gat_aggregate_f32': (ctypes.c_int, [_P, _P, _I32, _P, _I64, _I32, _P, _F32, _P, _I64, _P, _P]),
    'gist_gat_backward_workspace_bytes': (_SZ, [_I32, _I32]),
    'gist_gat_backward_f32': (ctypes.c_int, [_P, _P, _P, _P, _I32, _P, _I64, _I32, _P, _P, _P, _F32, _P, _I64,
                                             _P, _I64, _P, _I64, _P, _P, _SZ, _P]),
    'gist_adam_multi_f32': (ctypes.c_int, [_I32, _P, _P, _P, _P, _P, _F32, _F32, _F32, _F32, _F32, _P, _P, _P]),
    'gist_adam_multi_ex_f32': (ctypes.c_int, [_I32, _P, _P, _P, _P, _P, _F32, _F32, _F32, _F32, _F32, _P, _P, _P, _P, _P]),
    'gist_set_pdl': (ctypes.c_int, [_I32]),
    'gist_get_pdl': (ctypes.c_int, []),
}

SPMM_RELU, SPMM_NARROW, SPMM_WIDE = 1, 2, 4
SPMM_SLAB_OFF = 32
SPMM_BG_SHIFT = 8
SPMM_VEC_SHIFT = 16        # flags |= code << 16: 1, 2, 3 -> gathers of at most 32, 64, 128 bits
SPMM_LANES_SHIFT = 12      # flags |= code << 12: 1, 2, 3, 4 -> 4, 8, 16, 32 lanes per row
NORM_INV, NORM_RSQRT_CLAMP = 0, 1
GEMM_RELU, GEMM_NO_SPLITK, GEMM_TILE_N64, GEMM_TILE_N128, GEMM_TILE_N256 = 1, 2, 4, 8, 16
GEMM_BACKGROUND = 32
GEMM_K_MAJOR, GEMM_MN_MAJOR = 0, 1
ACT_RELU = 1
SPMM_SCHED_PREFETCH = 1



class DropoutDesc(ctypes.Structure):
    """gist_dropout_t"""
    _fields_ = [('p', ctypes.c_float), ('seed', ctypes.c_uint64), ('stream_id', ctypes.c_uint32),
                ('step', ctypes.c_void_p), ('step_saved', ctypes.c_void_p)]


class SpmmSchedule(ctypes.Structure):
    """gist_spmm_schedule_t"""
    _fields_ = [('seg_ptr', ctypes.c_void_p), ('seg_row', ctypes.c_void_p), ('seg_len', ctypes.c_int32),
                ('max_segments', ctypes.c_int64), ('counters', ctypes.c_void_p), ('workspace', ctypes.c_void_p),
                ('ld_workspace', ctypes.c_int64), ('seg_meta', ctypes.c_void_p), ('flags', ctypes.c_uint32)]


class SpmmEx(ctypes.Structure):
    """gist_spmm_ex_t"""
    _fields_ = [('y_lo', ctypes.c_void_p), ('ld_y_lo', ctypes.c_int64), ('self_lo', ctypes.c_void_p),
                ('ld_self_lo', ctypes.c_int64), ('drop', ctypes.POINTER(DropoutDesc)),
                ('drop_col0_y', ctypes.c_int32), ('drop_col0_self', ctypes.c_int32),
                ('schedule', ctypes.POINTER(SpmmSchedule))]


class SliceJob(ctypes.Structure):
    """gist_slice_job_t"""
    _fields_ = [('src', ctypes.c_void_p), ('ld_src', ctypes.c_int64), ('ridx', ctypes.c_void_p), ('n_rows', ctypes.c_int64),
                ('cidx', ctypes.c_void_p), ('n_cols', ctypes.c_int64), ('dst', ctypes.c_void_p), ('ld_dst', ctypes.c_int64)]


class SliceRowsJob(ctypes.Structure):
    """gist_slice_rows_job_t"""
    _fields_ = [('src', ctypes.c_void_p), ('ld_src', ctypes.c_int64), ('ridx', ctypes.c_void_p), ('n_rows', ctypes.c_int64),
                ('n_cols', ctypes.c_int64), ('inv_col', ctypes.c_void_p), ('dst', ctypes.c_void_p), ('ld_dst', ctypes.c_int64),
                ('dst_cols', ctypes.c_int64)]


class GemmEx(ctypes.Structure):
    """gist_gemm_ex_t"""
    _fields_ = [('tile_counters', ctypes.c_void_p), ('n_counters', ctypes.c_int64), ('rowsum', ctypes.c_void_p),
                ('ln_out', ctypes.c_void_p), ('ld_ln', ctypes.c_int64), ('ln_stats', ctypes.c_void_p),
                ('ln_eps', ctypes.c_float), ('ln_flags', ctypes.c_uint32), ('drop', ctypes.POINTER(DropoutDesc))]


_lib = None
_device_set = None


class GistLibraryError(RuntimeError):
    pass


def load():
    """Load (once) and return the ctypes handle; raise loudly if it is not built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise GistLibraryError(
            'gist_b200: %s not found. Build it with `python -c "import __graft_entry__ as g; '
            'g.build()"` or `make -C gist_b200/csrc`. There is no CPU / PyTorch fallback.' % LIB_PATH)
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the symbol is not exported
        fn.restype = res
        fn.argtypes = args
    if lib.gist_abi_version() != 1:
        raise GistLibraryError('gist_b200: ABI version mismatch')
    _lib = lib
    return lib


def check(status, what):
    if status != 0:
        msg = load().gist_status_string(status)
        raise GistLibraryError('gist_b200.%s failed: %s (status %d)'
                               % (what, msg.decode() if msg else '?', status))


def stream_ptr(device):
    """Raw cudaStream_t of torch's current stream on `device` and bind the runtime."""
    import torch
    global _device_set
    idx = device.index if device.index is not None else torch.cuda.current_device()
    if _device_set != idx:
        check(load().gist_set_device(idx), 'set_device')
        _device_set = idx
    return ctypes.c_void_p(torch.cuda.current_stream(idx).cuda_stream)


def ptr(t):
    """Device pointer of a tensor (None -> NULL)."""
    if t is None:
        return None
    return ctypes.c_void_p(t.data_ptr())


def require_cuda(*tensors):
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise GistLibraryError(
                'gist_b200 has no CPU path: got a %s tensor; move the graph / features to a '
                'CUDA device (B200, sm_100a)' % t.device)


def launch_count():
    return int(load().gist_launch_count())


def set_pdl(enabled):
    """Programmatic dependent launch for the training step's kernel chain (gist_set_pdl); process-wide.
    Set it BEFORE a step is captured: the attribute is recorded into the graph's edges."""
    check(load().gist_set_pdl(1 if enabled else 0), 'set_pdl')


def get_pdl():
    return bool(load().gist_get_pdl())
