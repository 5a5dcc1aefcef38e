"""ctypes binding of libmeshcnn_b200.so (the C ABI declared in include/meshcnn_b200.h).

There is no CPU fallback: if the library is missing, ``lib()`` raises; every layer calls it on first use.
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get('MESHCNN_B200_LIB') or os.path.join(_HERE, 'libmeshcnn_b200.so')   # override: diagnostics builds only

_c_int = ctypes.c_int
_c_vp = ctypes.c_void_p
_c_sz = ctypes.c_size_t


class PoolArgs(ctypes.Structure):
    """mcnn_pool_args (include/meshcnn_b200.h)"""
    _fields_ = [
        ('B', ctypes.c_int32), ('E_in', ctypes.c_int32), ('T', ctypes.c_int32), ('V', ctypes.c_int32),
        ('ve_cap', ctypes.c_int32), ('nnz_cap', ctypes.c_int32), ('log_cap', ctypes.c_int32), ('reserved', ctypes.c_int32),
        ('gemm_in', _c_vp), ('sides_in', _c_vp), ('ec_in', _c_vp), ('norms', _c_vp),
        ('edges', _c_vp), ('edges_stride', ctypes.c_int32), ('reserved2', ctypes.c_int32),
        ('ve_ptr', _c_vp), ('ve_idx', _c_vp), ('v_mask', _c_vp), ('vs', _c_vp),
        ('gemm_out', _c_vp), ('sides_out', _c_vp), ('ec_out', _c_vp),
        ('grp_rowptr', _c_vp), ('grp_cols', _c_vp), ('grp_colptr', _c_vp), ('grp_rows', _c_vp), ('occ', _c_vp),
        ('status', _c_vp),
        ('log_npops', _c_vp), ('log_pops', _c_vp), ('log_outcome', _c_vp), ('log_nunions', _c_vp), ('log_unions', _c_vp),
        ('edge_mask_out', _c_vp),
        ('workspace', _c_vp), ('workspace_stride', ctypes.c_longlong), ('force_workspace', ctypes.c_int32), ('reserved3', ctypes.c_int32),
        ('dbg_clocks', _c_vp),
    ]


# name -> (restype, argtypes); must list every symbol of include/meshcnn_b200.h (tests/test_cabi.py checks it)
SIGNATURES = {
    'mcnn_version': (_c_int, []),
    'mcnn_launch_count': (ctypes.c_ulonglong, []),
    'mcnn_last_cuda_error': (ctypes.c_char_p, []),
    'mcnn_meshconv_fwd': (_c_int, [_c_vp] * 7 + [_c_int] * 4 + [_c_vp]),
    'mcnn_meshconv_tc_supported': (_c_int, [_c_int, _c_int]),
    'mcnn_meshconv_tc_weight_images': (_c_int, [_c_vp, _c_vp, _c_vp, _c_int, _c_int, _c_vp]),
    'mcnn_meshconv_fwd_tc': (_c_int, [_c_vp] * 7 + [_c_int] * 4 + [_c_vp]),
    'mcnn_meshconv_fwd_tc_signs': (_c_int, [_c_vp] * 8 + [_c_int] * 4 + [_c_vp]),
    'mcnn_meshconv_fwd_tc_sk_ws': (_c_sz, []),
    'mcnn_meshconv_fwd_tc_sk': (_c_int, [_c_vp] * 9 + [_c_int] * 4 + [_c_vp]),
    'mcnn_meshconv_fwd_tc_sk_p': (_c_int, [_c_vp] * 9 + [_c_int] * 5 + [_c_vp]),
    'mcnn_meshconv_bwd_data_tc_signs_p': (_c_int, [_c_vp] * 8 + [_c_int] * 5 + [_c_vp]),
    'mcnn_meshconv_bwd_weight_tc_p': (_c_int, [_c_vp] * 7 + [_c_int] * 5 + [_c_vp]),
    'mcnn_meshconv_bwd_data_tc_signs': (_c_int, [_c_vp] * 8 + [_c_int] * 4 + [_c_vp]),
    'mcnn_meshconv_bwd_data': (_c_int, [_c_vp] * 6 + [_c_int] * 4 + [_c_vp]),
    'mcnn_meshconv_bwd_weight': (_c_int, [_c_vp] * 6 + [_c_int] * 4 + [_c_vp]),
    'mcnn_meshconv_bwd_data_tc': (_c_int, [_c_vp] * 7 + [_c_int] * 4 + [_c_vp]),
    'mcnn_meshconv_bwd_weight_tc_ws': (_c_sz, [_c_int] * 4),
    'mcnn_meshconv_bwd_weight_tc': (_c_int, [_c_vp] * 7 + [_c_int] * 4 + [_c_vp]),
    'mcnn_pool_norms': (_c_int, [_c_vp, _c_vp, _c_int, _c_int, _c_int, _c_vp]),
    'mcnn_pool_collapse': (_c_int, [ctypes.POINTER(PoolArgs), _c_vp]),
    'mcnn_pool_smem_bytes': (_c_sz, [_c_int, _c_int, _c_int, _c_int]),
    'mcnn_pool_smem_limit': (_c_sz, []),
    'mcnn_pool_workspace_bytes': (_c_sz, [_c_int, _c_int, _c_int, _c_int]),
    'mcnn_segmean_fwd': (_c_int, [_c_vp] * 5 + [_c_int] * 5 + [_c_vp]),
    'mcnn_segmean_bwd': (_c_int, [_c_vp] * 6 + [_c_int] * 5 + [_c_vp]),
    'mcnn_unpool_fwd': (_c_int, [_c_vp] * 6 + [_c_int] * 6 + [_c_vp]),
    'mcnn_norm_ws': (_c_sz, [_c_int] * 4),
    'mcnn_norm_slabs': (_c_int, [_c_int] * 4),
    'mcnn_norm_stats': (_c_int, [_c_vp] * 3 + [_c_int] * 5 + [_c_vp]),
    'mcnn_norm_bwd_stats': (_c_int, [_c_vp] * 7 + [_c_int] * 6 + [_c_vp]),
    'mcnn_norm_bwd_apply': (_c_int, [_c_vp] * 11 + [_c_int] * 6 + [_c_vp]),
    'mcnn_norm_fwd': (_c_int, [_c_vp] * 11 + [ctypes.c_float, ctypes.c_float] + [_c_int] * 7 + [_c_vp]),
    'mcnn_norm_bwd': (_c_int, [_c_vp] * 14 + [_c_int] * 6 + [_c_vp]),
    'mcnn_norm_bwd_acc': (_c_int, [_c_vp] * 15 + [_c_int] * 6 + [_c_vp]),
    'mcnn_adam_flat': (_c_int, [_c_vp] * 4 + [ctypes.c_longlong] + [ctypes.c_float] * 5 + [_c_int, _c_vp]),
    'mcnn_adam_flat_state': (_c_int, [_c_vp] * 4 + [ctypes.c_longlong, _c_vp, _c_vp]),
    'mcnn_unpool_bwd': (_c_int, [_c_vp] * 6 + [_c_int] * 6 + [_c_vp]),
    'mcnn_host_drop_non_manifold': (_c_int, [_c_vp, _c_vp, _c_int, _c_int, _c_vp]),
    'mcnn_host_edge_features': (_c_int, [_c_vp, _c_int, _c_vp, _c_vp, _c_int] + [_c_vp] * 4),
    'mcnn_host_build_gemm': (_c_int, [_c_vp, _c_vp, _c_int, _c_int] + [_c_vp] * 6),
    'mcnn_head_fwd': (_c_int, [_c_vp] * 8 + [_c_int] * 5 + [_c_vp]),
    'mcnn_head_bwd': (_c_int, [_c_vp] * 11 + [_c_int] * 5 + [_c_vp]),
}

# diagnostics switches / profiling read-outs (include/meshcnn_b200_debug.h): process-global, not part of the production ABI
DEBUG_SIGNATURES = {
    'mcnn_debug_set_tc_sk_ctas': (_c_int, [_c_int]),
    'mcnn_debug_set_tc_bd_tile': (_c_int, [_c_int]),
    'mcnn_debug_set_tc_bw_ot': (_c_int, [_c_int]),
    'mcnn_debug_set_tc_bw_tmem_a': (_c_int, [_c_int]),
    'mcnn_debug_set_pool_mode': (_c_int, [_c_int]),
    'mcnn_debug_set_pdl': (_c_int, [_c_int]),
    'mcnn_debug_sk_prof': (_c_int, [_c_vp, _c_int]),
    'mcnn_debug_sk_cta_times': (_c_int, [_c_vp]),
    'mcnn_debug_set_tc': (_c_int, [_c_int]),
    'mcnn_debug_set_tc_ntile': (_c_int, [_c_int]),
    'mcnn_debug_set_tc_pair': (_c_int, [_c_int]),
    'mcnn_debug_tc_prof': (_c_int, [_c_vp, _c_int]),
    'mcnn_debug_set_norm_fused': (_c_int, [_c_int]),
}

# CTA-pair (tcgen05 cta_group::2) forward kernel for Cout % 256 == 0; MESHCNN_B200_TC_PAIR=0/1 overrides
TC_PAIR_DEFAULT = int(os.environ.get('MESHCNN_B200_TC_PAIR', '0'))

PREC_FP32, PREC_TF32 = 0, 1                   # MCNN_PREC_* (include/meshcnn_b200.h)

_ERRORS = {1: 'MCNN_E_BADARG', 2: 'MCNN_E_TOO_LARGE', 3: 'MCNN_E_CUDA'}
_lib = None


def lib():
    """The loaded shared library; raises if it has not been built (no fallback path exists)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                'meshcnn_b200: %s is missing. Build it with `python -c "import __graft_entry__ as g; g.build()"` '
                'or `make -C meshcnn_b200/csrc`. There is no CPU fallback.' % LIB_PATH)
        l = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in list(SIGNATURES.items()) + list(DEBUG_SIGNATURES.items()):
            fn = getattr(l, name)
            fn.restype = res
            fn.argtypes = args
        l.mcnn_debug_set_tc_pair(TC_PAIR_DEFAULT)
        if os.environ.get('MESHCNN_B200_NORM_FUSED') == '0':     # A/B switch: statistics / finalize / apply kernels everywhere
            l.mcnn_debug_set_norm_fused(0)
        _lib = l
    return _lib


def check(rc, what):
    if rc != 0:
        msg = _ERRORS.get(rc, str(rc))
        if rc == 3:
            msg += ': ' + lib().mcnn_last_cuda_error().decode()
        raise RuntimeError('%s failed with %s' % (what, msg))


def ptr(t):
    """Device pointer of a tensor (None -> NULL)."""
    return None if t is None else t.data_ptr()


def current_stream():
    import torch
    return torch.cuda.current_stream().cuda_stream


def launch_count():
    return int(lib().mcnn_launch_count())
