"""ctypes binding of libntf_b200.so (the C ABI declared in include/ntf_b200.h).

There is no fallback: if the shared library is missing or a call fails, an exception is raised.  The library
is built in-tree by `python -m opentf_b200.csrc.build` (or `__graft_entry__.build()`).
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get('NTF_B200_LIB') or os.path.join(_HERE, 'csrc', 'libntf_b200.so')  # (NTF_B200_LIB: bug-hunt builds, scripts/flip_variants.sh)

NTF_FP32, NTF_TF32 = 0, 1
NSD = {None: 0, '': 0, 'none': 0, 'None': 0, 'uniform': 1, 'unigram': 2, 'unigram_b': 3}
PRECISION = {'fp32': NTF_FP32, 'tf32': NTF_TF32}

vp, i32, i64, u32, u64, f32, f64, sz = C.c_void_p, C.c_int, C.c_int64, C.c_uint32, C.c_uint64, C.c_float, C.c_double, C.c_size_t


class OutTrainArgs(C.Structure):
    """mirror of ntf_out_train_args"""
    _fields_ = [('A', vp), ('W', vp), ('b', vp), ('special', vp), ('pitch_words', i32), ('m_indptr', vp), ('m_indices', vp),
                ('B', i32), ('h', i32), ('E', i32), ('tpw', f32), ('tnw', f32), ('loss_scale', f32),
                ('dW', vp), ('db', vp), ('dA', vp), ('loss_out', vp),
                ('A_s', vp), ('W_delta', vp), ('b_delta', vp), ('sign_out', vp), ('dW_delta', vp), ('db_delta', vp), ('dA_s', vp), ('special_t', vp), ('member_t', vp), ('e_lo', i32), ('A16', vp), ('W16', vp), ('neg', vp), ('ns', i32), ('act_prev', vp), ('dz_prev', vp), ('db_prev', vp), ('prepared', i32), ('defer_finish', i32)]


class InferTopkArgs(C.Structure):
    """mirror of ntf_infer_topk_args"""
    _fields_ = [('A', vp), ('A16', vp), ('W16', vp), ('b', vp), ('B', i32), ('h', i32), ('E', i32), ('K', i32), ('e_lo', i32), ('vals', vp), ('idx', vp)]


NTF_MAX_LAYERS = 8
_PA = vp * NTF_MAX_LAYERS


NTF_MAX_PEERS = 8


class Peers(C.Structure):
    """mirror of ntf_peers"""
    _fields_ = [('rank', i32), ('world', i32), ('grads', vp * NTF_MAX_PEERS), ('params', vp * NTF_MAX_PEERS), ('flags', vp * NTF_MAX_PEERS)]


class FnnInferTopkArgs(C.Structure):
    """mirror of ntf_fnn_infer_topk_args"""
    _fields_ = [('n_layers', i32), ('S', i32), ('E', i32), ('e_lo', i32), ('hidden', i32 * NTF_MAX_LAYERS), ('W', _PA), ('b', _PA), ('act', _PA),
                ('B', i32), ('s_indptr', vp), ('s_indices', vp), ('x_dense', vp), ('W16', vp), ('K', i32), ('vals', vp), ('idx', vp)]


class FnnStepArgs(C.Structure):
    """mirror of ntf_fnn_step_args"""
    _fields_ = [('n_layers', i32), ('S', i32), ('E', i32), ('e_lo', i32), ('E_total', i32), ('phase', i32), ('hidden', i32 * NTF_MAX_LAYERS),
                ('W', _PA), ('b', _PA), ('gW', _PA), ('gb', _PA), ('act', _PA), ('dact', _PA), ('dz', _PA),
                ('B', i32), ('s_indptr', vp), ('s_indices', vp), ('s_ent_row', vp), ('row_base', i32), ('m_indptr', vp), ('m_indices', vp),
                ('gB', i32), ('g_m_indptr', vp), ('nsd', i32), ('seed', u64), ('step', u64), ('row0', i32), ('ns', i32), ('neg_given', i32),
                ('neg', vp), ('counts', vp), ('cdf', vp), ('precision', i32), ('tpw', f32), ('tnw', f32), ('loss_scale', f32), ('loss_out', vp),
                ('special', vp), ('pitch_words', i32), ('special_t', vp), ('member_t', vp), ('train', i32), ('run_adam', i32),
                ('params', vp), ('grads', vp), ('adam_m', vp), ('adam_v', vp), ('n_params', sz),
                ('lr', f64), ('beta1', f64), ('beta2', f64), ('eps', f64), ('adam_t', i64), ('prof_ev', vp * 2), ('dyn', vp), ('comm', vp), ('allreduce', vp), ('peers', vp), ('x_dense', vp), ('W16', vp)]


# name -> (restype, argtypes); must list every symbol of include/ntf_b200.h (tests/test_abi.py checks it)
SIGNATURES = {
    'ntf_version': (i32, []),
    'ntf_last_error': (i32, [C.c_char_p, sz]),
    'ntf_create': (i32, [i32, C.POINTER(vp)]),
    'ntf_destroy': (i32, [vp]),
    'ntf_sm_count': (i32, [vp]),
    'ntf_launch_count': (C.c_ulonglong, [i32]),
    'ntf_dyn_update': (i32, [vp, vp, vp, u64, f64, f64, f64, f64, i64]),
    'ntf_set_dyn': (i32, [vp, vp]),
    'ntf_graph_begin': (i32, [vp, vp]),
    'ntf_graph_end': (i32, [vp, vp, C.POINTER(vp)]),
    'ntf_graph_launch': (i32, [vp, vp]),
    'ntf_graph_kernels': (i32, [vp]),
    'ntf_graph_destroy': (i32, [vp]),
    'ntf_csr_gather_workspace_bytes': (sz, [i32]),
    'ntf_csr_gather': (i32, [vp, vp, vp, i32, vp, vp, vp, vp, vp, vp, sz]),
    'ntf_peer_alloc': (i32, [vp, sz, C.POINTER(vp)]),
    'ntf_peer_free': (i32, [vp, vp]),
    'ntf_peer_export': (i32, [vp, vp, vp]),
    'ntf_peer_import': (i32, [vp, vp, C.POINTER(vp)]),
    'ntf_peer_release': (i32, [vp, vp]),
    'ntf_peer_exchange_adam': (i32, [vp, vp, vp, vp, vp, sz, sz, f64, f64, f64, f64, i64, i32]),
    'ntf_peer_allreduce': (i32, [vp, vp, vp, sz, vp, i32]),
    'ntf_rows_gather': (i32, [vp, vp, vp, i32, i32, vp, vp]),
    'ntf_csr_bag_fwd': (i32, [vp, vp, i32, vp, vp, vp, vp, i32, i32, vp]),
    'ntf_csr_bag_bwd_workspace_bytes': (sz, [i32, i32]),
    'ntf_csr_bag_bwd': (i32, [vp, vp, i32, vp, vp, vp, i32, vp, i32, i32, vp, vp, sz]),
    'ntf_dense_fwd': (i32, [vp, vp, vp, vp, vp, i32, i32, i32, i32, vp]),
    'ntf_act_bwd_workspace_bytes': (sz, [i32, i32]),
    'ntf_act_bwd': (i32, [vp, vp, vp, vp, i32, i32, i32, vp, vp, vp, sz]),
    'ntf_dense_bwd_workspace_bytes': (sz, [i32, i32, i32]),
    'ntf_dense_bwd': (i32, [vp, vp, vp, vp, vp, i32, i32, i32, vp, vp, vp, sz]),
    'ntf_expert_cdf_workspace_bytes': (sz, [i32]),
    'ntf_expert_cdf': (i32, [vp, vp, i32, vp, vp, i32, vp, vp, vp, sz]),
    'ntf_neg_sample': (i32, [vp, vp, i32, u64, u64, i32, i32, vp, vp, i32, i32, vp, vp, i32, vp]),
    'ntf_special_bits': (i32, [vp, vp, i32, i32, vp, vp, vp, i32, i32, i32, vp, i32]),
    'ntf_special_tiles_bytes': (sz, [i32, i32]),
    'ntf_special_tiles': (i32, [vp, vp, i32, i32, vp, vp, vp, i32, i32, i32, vp, vp]),
    'ntf_tc_supported': (i32, [i32, i32, i32, i32]),
    'ntf_out_train_workspace_bytes': (sz, [vp, i32, i32, i32, i32, i32]),
    'ntf_out_train': (i32, [vp, vp, i32, C.POINTER(OutTrainArgs), vp, sz]),
    'ntf_out_train_finish': (i32, [vp, vp, C.POINTER(OutTrainArgs), vp, sz]),
    'ntf_out_train_prepare': (i32, [vp, vp, i32, i32, i32, vp, vp, vp]),
    'ntf_adam_step': (i32, [vp, vp, vp, vp, vp, vp, sz, f64, f64, f64, f64, i64]),
    'ntf_infer_scores_workspace_bytes': (sz, [i32, i32, i32]),
    'ntf_infer_scores': (i32, [vp, vp, i32, vp, vp, vp, i32, i32, i32, vp, vp, vp, vp, i32, i32, vp, vp, sz]),
    'ntf_topk_select': (i32, [vp, vp, vp, i32, i32, i32, f32, vp, vp]),
    'ntf_infer_topk_supported': (i32, [i32, i32, i32, i32]),
    'ntf_infer_topk_workspace_bytes': (sz, [i32, i32, i32, i32]),
    'ntf_infer_topk': (i32, [vp, vp, C.POINTER(InferTopkArgs), vp, sz]),
    'ntf_to_half': (i32, [vp, vp, vp, sz, vp]),
    'ntf_topk_merge': (i32, [vp, vp, vp, vp, i32, i32, i32, vp, vp]),
    'ntf_row_entropy': (i32, [vp, vp, vp, i32, i32, f32, i32, vp]),
    'ntf_eval_ranked': (i32, [vp, vp, i32, i32, vp, vp, vp, vp, C.POINTER(i32), i32, vp]),
    'ntf_host_batch_upload': (i32, [vp, vp, i32, vp, vp, sz]),
    'ntf_host_loss_download': (i32, [vp, vp, i32, vp, vp]),
    'ntf_host_loss_wait': (i32, [vp, i32]),
    'ntf_csr_from_lists_workspace_bytes': (sz, [i32, sz]),
    'ntf_csr_from_lists': (i32, [vp, vp, i32, sz, vp, vp, i32, vp, vp, C.POINTER(C.c_int64), vp, sz]),
    'ntf_cooccur_workspace_bytes': (sz, [i32, sz]),
    'ntf_cooccur_count': (i32, [vp, vp, i32, i32, i32, sz, vp, vp, vp, vp, vp, vp, C.POINTER(C.c_int64), vp, sz]),
    'ntf_cooccur_fill': (i32, [vp, vp, i32, i32, i32, sz, vp, vp, vp, vp, vp, vp, sz]),
    'ntf_skill_coverage': (i32, [vp, vp, i32, i32, vp, vp, vp, vp, vp, C.POINTER(i32), i32, vp]),
    'ntf_axpy': (i32, [vp, vp, sz, f32, vp, vp]),
    'ntf_flipout_prepare_workspace_bytes': (sz, [vp]),
    'ntf_flipout_prepare': (i32, [vp, vp, vp, vp, vp, sz, f32, vp, vp, vp, sz]),
    'ntf_flipout_grads': (i32, [vp, vp, vp, vp, vp, vp, sz, f32, vp, vp]),
    'ntf_fill_normal': (i32, [vp, vp, u64, u64, u32, sz, vp]),
    'ntf_fill_sign_bits': (i32, [vp, vp, u64, u64, u32, sz, vp]),
    'ntf_apply_sign': (i32, [vp, vp, vp, vp, i32, i32, i32, vp]),
    'ntf_csr_bag_flipout_fwd': (i32, [vp, vp, i32, vp, vp, vp, vp, vp, vp, vp, vp, i32, i32, i32, vp]),
    'ntf_csr_bag_bwd_signed': (i32, [vp, vp, i32, vp, vp, vp, i32, vp, vp, i32, i32, vp, vp, sz]),
    'ntf_dense_flipout_fwd_workspace_bytes': (sz, [i32, i32]),
    'ntf_dense_flipout_fwd': (i32, [vp, vp, vp, vp, vp, vp, vp, vp, vp, i32, i32, i32, i32, i32, vp, vp, sz]),
    'ntf_add_signed': (i32, [vp, vp, vp, vp, i32, i32, i32, vp]),
    'ntf_fnn_step_workspace_bytes': (sz, [vp, C.POINTER(FnnStepArgs)]),
    'ntf_fnn_step': (i32, [vp, vp, C.POINTER(FnnStepArgs), vp, sz]),
    'ntf_fnn_infer_topk_workspace_bytes': (sz, [vp, C.POINTER(FnnInferTopkArgs)]),
    'ntf_fnn_infer_topk': (i32, [vp, vp, C.POINTER(FnnInferTopkArgs), vp, sz]),
    'ntf_pack_host_batch': (i32, [vp, i32, vp, vp, vp, vp, i32, i32, i32, i32, vp, sz]),
    'ntf_sum_parts': (i32, [vp, vp, vp, i32, sz, sz, vp]),
}

_lib = None


class NtfError(RuntimeError):
    pass


def lib():
    """load (once) and type the shared library; raises if it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise NtfError(f'{LIB_PATH} is missing: build it with `python -m opentf_b200.csrc.build` (there is no CPU fallback)')
        l = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(l, name)
            fn.restype, fn.argtypes = res, args
        _lib = l
    return _lib


def last_error():
    buf = C.create_string_buffer(512)
    lib().ntf_last_error(buf, 512)
    return buf.value.decode(errors='replace')


def check(rc, what=''):
    if rc != 0:
        raise NtfError(f'{what} failed with status {rc}: {last_error()}')


def new_ctx(device_index):
    """a handle of its own (side streams + events of ntf_fnn_step's fork/join): one per Engine, so that engines driven from
    different host threads never share them"""
    h = vp()
    check(lib().ntf_create(int(device_index), C.byref(h)), 'ntf_create')
    return h


_ctx = {}


def ctx(device_index):
    """one ntf_ctx per CUDA device (fails loudly on anything that is not sm_100)."""
    if device_index not in _ctx:
        h = vp()
        check(lib().ntf_create(int(device_index), C.byref(h)), 'ntf_create')
        _ctx[device_index] = h
    return _ctx[device_index]
