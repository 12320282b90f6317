"""Typed Python entry points over the C ABI: torch tensors in, kernels launched on torch's current stream.

torch is used for device memory and streams only (every tensor handed in is a caller-owned device buffer);
all arithmetic happens in libntf_b200.so.  Nothing here falls back to torch ops.
"""
import ctypes as C

import torch

from . import _lib
from ._lib import OutTrainArgs, check, lib

F32, I32 = torch.float32, torch.int32


def _dev(t):
    assert t.is_cuda, 'libntf_b200 works on CUDA buffers only (no CPU fallback)'
    return t.device.index if t.device.index is not None else torch.cuda.current_device()


def _stream(dev):
    return torch.cuda.current_stream(dev).cuda_stream


def _p(t, dtype=None):
    if t is None: return None
    assert t.is_cuda and t.is_contiguous(), 'device-resident contiguous tensor required'
    if dtype is not None: assert t.dtype == dtype, f'expected {dtype}, got {t.dtype}'
    return t.data_ptr()


class Workspace:
    """a grow-only scratch buffer owned by the caller side (the library never allocates)."""

    def __init__(self, device):
        self.device, self.buf, self.gen = device, None, 0  # gen counts re-allocations (captured graphs hold the old address)

    def get(self, nbytes):
        if nbytes <= 0: return None, 0
        if self.buf is None or self.buf.numel() < nbytes:
            self.buf = torch.empty(int(nbytes), dtype=torch.uint8, device=self.device)
            self.gen += 1
        return self.buf.data_ptr(), self.buf.numel()


def csr_gather(rows, n, src_indptr, src_indices, dst_indptr, dst_indices, dst_ent_row, ws):
    d = _dev(src_indptr)
    p, nb = ws.get(lib().ntf_csr_gather_workspace_bytes(n))
    check(lib().ntf_csr_gather(_lib.ctx(d), _stream(d), _p(rows, I32), n, _p(src_indptr, I32), _p(src_indices, I32),
                               _p(dst_indptr, I32), _p(dst_indices, I32), _p(dst_ent_row, I32), p, nb), 'ntf_csr_gather')


# ---- peer memory (data-parallel ranks): include/ntf_b200.h "gradient exchange + optimiser in one pass over peer memory"
class _RawCuda:
    """a cudaMalloc block of the library seen as a torch tensor (zero copy, __cuda_array_interface__)"""

    def __init__(self, ptr, shape, typestr):
        self.ptr = ptr
        self.__cuda_array_interface__ = {'shape': tuple(shape), 'typestr': typestr, 'data': (ptr, False), 'version': 2}


def peer_alloc(dev_index, n, dtype):
    """-> (tensor of n zeros living in an IPC-exportable cudaMalloc block, raw pointer)"""
    ptr = _lib.vp()
    item = 4
    check(lib().ntf_peer_alloc(_lib.ctx(dev_index), n * item, _lib.C.byref(ptr)), 'ntf_peer_alloc')
    t = torch.as_tensor(_RawCuda(ptr.value, (n,), '<f4' if dtype == torch.float32 else '<i4'), device=torch.device('cuda', dev_index))
    assert t.data_ptr() == ptr.value
    return t, ptr.value


def peer_export(dev_index, ptr):
    h = _lib.C.create_string_buffer(64)
    check(lib().ntf_peer_export(_lib.ctx(dev_index), ptr, h), 'ntf_peer_export')
    return h.raw


def peer_import(dev_index, handle):
    out = _lib.vp()
    check(lib().ntf_peer_import(_lib.ctx(dev_index), _lib.C.create_string_buffer(handle, 64), _lib.C.byref(out)), 'ntf_peer_import')
    return out.value


def peer_exchange_adam(dev_index, peers, m, v, offset, n, lr, b1, b2, eps, step, channel):
    check(lib().ntf_peer_exchange_adam(_lib.ctx(dev_index), _stream(dev_index), _lib.C.addressof(peers), _p(m, F32), _p(v, F32), offset, n, lr, b1, b2,
                                       eps, step, channel), 'ntf_peer_exchange_adam')


def rows_gather(rows, n, d, src, dst):
    dv = _dev(src)
    check(lib().ntf_rows_gather(_lib.ctx(dv), _stream(dv), _p(rows, I32), n, d, _p(src, F32), _p(dst, F32)), 'ntf_rows_gather')


def csr_bag_fwd(B, indptr_ptr, indices, W0T, b0, S, h, A):
    d = _dev(W0T)
    check(lib().ntf_csr_bag_fwd(_lib.ctx(d), _stream(d), B, indptr_ptr, _p(indices, I32), _p(W0T, F32), _p(b0, F32), S, h, _p(A, F32)),
          'ntf_csr_bag_fwd')


def csr_bag_bwd(B, indptr_ptr, indices, ent_row, row_base, dZ, S, h, dW0T, ws):
    d = _dev(dZ)
    p, nb = ws.get(lib().ntf_csr_bag_bwd_workspace_bytes(S, h))
    check(lib().ntf_csr_bag_bwd(_lib.ctx(d), _stream(d), B, indptr_ptr, _p(indices, I32), _p(ent_row, I32), row_base, _p(dZ, F32), S, h,
                                _p(dW0T, F32), p, nb), 'ntf_csr_bag_bwd')


def dense_fwd(A, W, b, B, inn, out, act, Y):
    d = _dev(A)
    check(lib().ntf_dense_fwd(_lib.ctx(d), _stream(d), _p(A, F32), _p(W, F32), _p(b, F32), B, inn, out, act, _p(Y, F32)), 'ntf_dense_fwd')


def act_bwd(dY, Y, B, h, act, dZ, db, ws):
    d = _dev(dY)
    p, nb = ws.get(lib().ntf_act_bwd_workspace_bytes(B, h))
    check(lib().ntf_act_bwd(_lib.ctx(d), _stream(d), _p(dY, F32), _p(Y, F32), B, h, act, _p(dZ, F32), _p(db, F32), p, nb), 'ntf_act_bwd')


def dense_bwd(A, W, dZ, B, inn, out, dW, dA, ws):
    d = _dev(A)
    p, nb = ws.get(lib().ntf_dense_bwd_workspace_bytes(B, inn, out))
    check(lib().ntf_dense_bwd(_lib.ctx(d), _stream(d), _p(A, F32), _p(W, F32), _p(dZ, F32), B, inn, out, _p(dW, F32), _p(dA, F32), p, nb),
          'ntf_dense_bwd')


def expert_cdf(B, m_indptr_ptr, m_indices, E, counts, cdf, ws):
    d = _dev(m_indices)
    p, nb = ws.get(lib().ntf_expert_cdf_workspace_bytes(E))
    check(lib().ntf_expert_cdf(_lib.ctx(d), _stream(d), B, m_indptr_ptr, _p(m_indices, I32), E, _p(counts), _p(cdf), p, nb), 'ntf_expert_cdf')


def neg_sample(nsd, seed, step, row0, B, m_indptr_ptr, m_indices, E, ns, cdf, neg, pool_indptr_ptr=None, pool_rows=0):
    d = _dev(m_indices)
    check(lib().ntf_neg_sample(_lib.ctx(d), _stream(d), nsd, seed, step, row0, B, m_indptr_ptr, _p(m_indices, I32), E, ns, _p(cdf),
                               pool_indptr_ptr, pool_rows, _p(neg, I32)), 'ntf_neg_sample')


def special_bits(op, B, m_indptr_ptr, m_indices, neg, ns, E, special, pitch, e_lo=0):
    d = _dev(m_indices)
    check(lib().ntf_special_bits(_lib.ctx(d), _stream(d), op, B, m_indptr_ptr, _p(m_indices, I32), _p(neg, I32), ns, E, e_lo, _p(special), pitch),
          'ntf_special_bits')


def special_tiles_bytes(B, E):
    return lib().ntf_special_tiles_bytes(B, E)


def special_tiles(op, B, m_indptr_ptr, m_indices, neg, ns, E, special_t, member_t, e_lo=0):
    d = _dev(m_indices)
    check(lib().ntf_special_tiles(_lib.ctx(d), _stream(d), op, B, m_indptr_ptr, _p(m_indices, I32), _p(neg, I32), ns, E, e_lo, _p(special_t), _p(member_t)),
          'ntf_special_tiles')


def out_train_workspace_bytes(dev, precision, B, h, E, flipout):
    return lib().ntf_out_train_workspace_bytes(_lib.ctx(dev), precision, B, h, E, int(flipout))


def out_train(dev, precision, args, ws):
    p, nb = ws.get(out_train_workspace_bytes(dev, precision, args.B, args.h, args.E, bool(args.W_delta)))
    check(lib().ntf_out_train(_lib.ctx(dev), _stream(dev), precision, C.byref(args), p, nb), 'ntf_out_train')


def adam_step(p, g, m, v, n, lr, b1, b2, eps, step):
    d = _dev(p)
    check(lib().ntf_adam_step(_lib.ctx(d), _stream(d), _p(p, F32), _p(g, F32), _p(m, F32), _p(v, F32), n, lr, b1, b2, eps, step), 'ntf_adam_step')


def infer_scores(precision, A, W, b, B, h, E, P, ws, accumulate=0, A_s=None, W_delta=None, b_delta=None, sign_out=None, pitch=0):
    d = _dev(A)
    p, nb = ws.get(lib().ntf_infer_scores_workspace_bytes(B, E, int(W_delta is not None)))
    check(lib().ntf_infer_scores(_lib.ctx(d), _stream(d), precision, _p(A, F32), _p(W, F32), _p(b, F32), B, h, E, _p(A_s, F32), _p(W_delta, F32),
                                 _p(b_delta, F32), _p(sign_out), pitch, accumulate, _p(P, F32), p, nb), 'ntf_infer_scores')


def topk_select(P, B, E, K, scale, vals, idx):
    d = _dev(P)
    check(lib().ntf_topk_select(_lib.ctx(d), _stream(d), _p(P, F32), B, E, K, scale, _p(vals, F32), _p(idx, I32)), 'ntf_topk_select')


def infer_topk_supported(B, h, E, K):
    return bool(lib().ntf_infer_topk_supported(int(B), int(h), int(E), int(K)))


def to_half(x, n, y):
    d = _dev(x)
    assert y.dtype == torch.float16 and y.numel() >= n
    check(lib().ntf_to_half(_lib.ctx(d), _stream(d), _p(x, F32), n, _p(y)), 'ntf_to_half')


def infer_topk(A, W16, b, B, h, E, K, vals, idx, ws, e_lo=0, A16=None):
    """fused output layer + sigmoid + top-K (csrc/infer_topk.cu): only [B,K] values / ids are written"""
    d = _dev(W16)
    a = _lib.InferTopkArgs()
    a.A, a.A16, a.W16, a.b = _p(A, F32), _p(A16), _p(W16), _p(b, F32)
    a.B, a.h, a.E, a.K, a.e_lo = B, h, E, K, e_lo
    a.vals, a.idx = _p(vals, F32), _p(idx, I32)
    p, nb = ws.get(lib().ntf_infer_topk_workspace_bytes(B, h, E, K))
    check(lib().ntf_infer_topk(_lib.ctx(d), _stream(d), C.byref(a), p, nb), 'ntf_infer_topk')


def fnn_infer_topk(dev, args, ws, nbytes=None):
    """one call per test batch: hidden layers + fused output layer / sigmoid / top-K (ntf_fnn_infer_topk).  `nbytes`: the workspace size if the
    caller cached it for this (B, K)"""
    if nbytes is None: nbytes = lib().ntf_fnn_infer_topk_workspace_bytes(_lib.ctx(dev), C.byref(args))
    p, nb = ws.get(nbytes)
    check(lib().ntf_fnn_infer_topk(_lib.ctx(dev), _stream(dev), C.byref(args), p, nb), 'ntf_fnn_infer_topk')
    return nbytes


def topk_merge(vals_in, idx_in, G, B, K, vals, idx):
    d = _dev(vals_in)
    check(lib().ntf_topk_merge(_lib.ctx(d), _stream(d), _p(vals_in, F32), _p(idx_in, I32), G, B, K, _p(vals, F32), _p(idx, I32)), 'ntf_topk_merge')


def row_entropy(P, B, E, scale, accumulate, out):
    d = _dev(P)
    check(lib().ntf_row_entropy(_lib.ctx(d), _stream(d), _p(P, F32), B, E, scale, accumulate, _p(out, F32)), 'ntf_row_entropy')


def eval_ranked(idx, vals, m_indptr, m_indices, ks, out):
    """ranking metrics per team on the device (ntf_eval_ranked): idx/vals [n,K] candidates, member CSR rows, ks ascending cut-offs,
    out [n,5,len(ks)] fp64 in the order P, recall, ndcg_cut, map_cut, success"""
    d = _dev(vals)
    n, K = vals.shape
    assert out.dtype == torch.float64 and tuple(out.shape) == (n, 5, len(ks))
    karr = (C.c_int * len(ks))(*[int(k) for k in ks])
    check(lib().ntf_eval_ranked(_lib.ctx(d), _stream(d), n, K, _p(idx, I32), _p(vals, F32), _p(m_indptr, I32), _p(m_indices, I32), karr, len(ks),
                                _p(out, torch.float64)), 'ntf_eval_ranked')


def csr_from_lists(indptr, ids, n_cols, ws):
    """ragged id lists (device int32 indptr [n+1], ids) -> (dst_indptr [n+1], dst_indices [nnz]) sorted duplicate-free rows (ntf_csr_from_lists)"""
    d = _dev(indptr)
    n, n_ids = indptr.numel() - 1, ids.numel()
    dst_ptr = torch.empty(n + 1, dtype=I32, device=indptr.device)
    dst_idx = torch.empty(max(n_ids, 1), dtype=I32, device=indptr.device)
    nnz = C.c_int64(0)
    p, nb = ws.get(lib().ntf_csr_from_lists_workspace_bytes(n, n_ids))
    check(lib().ntf_csr_from_lists(_lib.ctx(d), _stream(d), n, n_ids, _p(indptr, I32), _p(ids, I32) if n_ids else None, n_cols, _p(dst_ptr, I32), _p(dst_idx, I32),
                                   C.byref(nnz), p, nb), 'ntf_csr_from_lists')
    return dst_ptr, dst_idx[:nnz.value]


def cooccur(m_indptr, m_indices, s_indptr, s_indices, E, S, skip, ws):
    """member^T . skill on the device (ntf_cooccur_count + ntf_cooccur_fill): (co_indptr [E+1], co_indices, co_values) int32, columns ascending.
    skip: uint8 [T] device tensor or None"""
    d = _dev(m_indptr)
    T, nnz_m = m_indptr.numel() - 1, m_indices.numel()
    co_ptr = torch.empty(E + 1, dtype=I32, device=m_indptr.device)
    nnz = C.c_int64(0)
    p, nb = ws.get(lib().ntf_cooccur_workspace_bytes(E, nnz_m))
    check(lib().ntf_cooccur_count(_lib.ctx(d), _stream(d), T, E, S, nnz_m, _p(m_indptr, I32), _p(m_indices, I32) if nnz_m else None, _p(s_indptr, I32), _p(s_indices, I32),
                                  _p(skip, torch.uint8), _p(co_ptr, I32), C.byref(nnz), p, nb), 'ntf_cooccur_count')
    co_idx = torch.empty(max(nnz.value, 1), dtype=I32, device=m_indptr.device)
    co_val = torch.empty(max(nnz.value, 1), dtype=I32, device=m_indptr.device)
    check(lib().ntf_cooccur_fill(_lib.ctx(d), _stream(d), T, E, S, nnz_m, _p(s_indptr, I32), _p(s_indices, I32), _p(co_ptr, I32), _p(co_idx, I32), _p(co_val, I32), p, nb),
          'ntf_cooccur_fill')
    return co_ptr, co_idx[:nnz.value], co_val[:nnz.value]


def skill_coverage(idx, x_indptr, x_indices, co_indptr, co_indices, ks, out):
    """out [n, len(ks)] fp64: covered fraction of every team's skills by its first k ranked experts (ntf_skill_coverage)"""
    d = _dev(idx)
    n, K = idx.shape
    assert out.dtype == torch.float64 and tuple(out.shape) == (n, len(ks))
    karr = (C.c_int * len(ks))(*[int(k) for k in ks])
    check(lib().ntf_skill_coverage(_lib.ctx(d), _stream(d), n, K, _p(idx, I32), _p(x_indptr, I32), _p(x_indices, I32), _p(co_indptr, I32), _p(co_indices, I32), karr,
                                   len(ks), _p(out, torch.float64)), 'ntf_skill_coverage')


def axpy(n, a, x, y):
    d = _dev(x)
    check(lib().ntf_axpy(_lib.ctx(d), _stream(d), n, a, _p(x, F32), _p(y, F32)), 'ntf_axpy')


def flipout_prepare(mu, rho, eps, n, kl_scale, delta, kl_out, ws):
    d = _dev(mu)
    p, nb = ws.get(lib().ntf_flipout_prepare_workspace_bytes(_lib.ctx(d)))
    check(lib().ntf_flipout_prepare(_lib.ctx(d), _stream(d), _p(mu, F32), _p(rho, F32), _p(eps, F32), n, kl_scale, _p(delta, F32), _p(kl_out, F32), p, nb),
          'ntf_flipout_prepare')


def flipout_grads(mu, rho, eps, g_delta, n, kl_gscale, g_mu, g_rho):
    d = _dev(mu)
    check(lib().ntf_flipout_grads(_lib.ctx(d), _stream(d), _p(mu, F32), _p(rho, F32), _p(eps, F32), _p(g_delta, F32), n, kl_gscale, _p(g_mu, F32),
                                  _p(g_rho, F32)), 'ntf_flipout_grads')


def fill_normal(seed, step, stream_id, n, out):
    d = _dev(out)
    check(lib().ntf_fill_normal(_lib.ctx(d), _stream(d), seed, step, stream_id, n, _p(out, F32)), 'ntf_fill_normal')


def fill_sign_bits(seed, step, stream_id, n_words, bits):
    d = _dev(bits)
    check(lib().ntf_fill_sign_bits(_lib.ctx(d), _stream(d), seed, step, stream_id, n_words, _p(bits)), 'ntf_fill_sign_bits')


def apply_sign(A, bits, pitch, B, h, As):
    d = _dev(A)
    check(lib().ntf_apply_sign(_lib.ctx(d), _stream(d), _p(A, F32), _p(bits), pitch, B, h, _p(As, F32)), 'ntf_apply_sign')


def sum_parts(parts, nparts, n, stride, out):
    d = _dev(parts)
    check(lib().ntf_sum_parts(_lib.ctx(d), _stream(d), _p(parts, F32), nparts, n, stride, _p(out, F32)), 'ntf_sum_parts')


def csr_bag_flipout_fwd(B, indptr_ptr, indices, ent_sign, Wmu, bmu, Wd, bd, sign_out, pitch, S, h, A):
    d = _dev(Wmu)
    check(lib().ntf_csr_bag_flipout_fwd(_lib.ctx(d), _stream(d), B, indptr_ptr, _p(indices, I32), _p(ent_sign), _p(Wmu, F32), _p(bmu, F32), _p(Wd, F32),
                                        _p(bd, F32), _p(sign_out), pitch, S, h, _p(A, F32)), 'ntf_csr_bag_flipout_fwd')


def csr_bag_bwd_signed(B, indptr_ptr, indices, ent_row, row_base, ent_sign, dZs, S, h, dWd, ws):
    d = _dev(dZs)
    p, nb = ws.get(lib().ntf_csr_bag_bwd_workspace_bytes(S, h))
    check(lib().ntf_csr_bag_bwd_signed(_lib.ctx(d), _stream(d), B, indptr_ptr, _p(indices, I32), _p(ent_row, I32), row_base, _p(ent_sign), _p(dZs, F32), S, h,
                                       _p(dWd, F32), p, nb), 'ntf_csr_bag_bwd_signed')


def dense_flipout_fwd(A, W, b, A_s, Wd, bd, sign_out, pitch, B, inn, out, act, Y, ws):
    d = _dev(A)
    p, nb = ws.get(lib().ntf_dense_flipout_fwd_workspace_bytes(B, out))
    check(lib().ntf_dense_flipout_fwd(_lib.ctx(d), _stream(d), _p(A, F32), _p(W, F32), _p(b, F32), _p(A_s, F32), _p(Wd, F32), _p(bd, F32), _p(sign_out), pitch,
                                      B, inn, out, act, _p(Y, F32), p, nb), 'ntf_dense_flipout_fwd')


def add_signed(X, bits, pitch, B, h, Y):
    d = _dev(X)
    check(lib().ntf_add_signed(_lib.ctx(d), _stream(d), _p(X, F32), _p(bits), pitch, B, h, _p(Y, F32)), 'ntf_add_signed')


def dyn_update(dev, dyn, step, lr, b1, b2, eps, adam_t):
    check(lib().ntf_dyn_update(_lib.ctx(dev), _stream(dev), _p(dyn), int(step), lr, b1, b2, eps, int(adam_t)), 'ntf_dyn_update')


def set_dyn(dev, dyn):
    """while set (a device ntf_dyn block, or None to clear): step / lr / adam_t arguments of the eager entry points come from the block"""
    check(lib().ntf_set_dyn(_lib.ctx(dev), None if dyn is None else _p(dyn)), 'ntf_set_dyn')


class Graph:
    """`with Graph(dev) as g: <library calls>` captures them on torch's current stream (nothing runs); g.launch() replays."""

    def __init__(self, dev, handle, capture_stream):
        """capture_stream: a torch stream of the caller's own (the legacy default stream, torch's usual current stream, cannot be
        captured); handle: the ntf_ctx the captured calls use"""
        self.dev, self.h, self.ctx_h, self.cap = dev, None, handle, capture_stream

    def __enter__(self):
        self._ctx = torch.cuda.stream(self.cap)  # the calls inside the block pick this stream up as "current"
        self._ctx.__enter__()
        check(lib().ntf_graph_begin(self.ctx_h, _stream(self.dev)), 'ntf_graph_begin')
        return self

    def __exit__(self, et, ev, tb):
        h = _lib.vp()
        rc = lib().ntf_graph_end(self.ctx_h, _stream(self.dev), C.byref(h))
        self._ctx.__exit__(et, ev, tb)
        if et is None: check(rc, 'ntf_graph_end')
        self.h = h if rc == 0 else None
        return False

    def launch(self):
        check(lib().ntf_graph_launch(self.h, _stream(self.dev)), 'ntf_graph_launch')

    def __del__(self):
        try:
            if self.h: lib().ntf_graph_destroy(self.h)
        except Exception: pass
        self.h = None


def fnn_step_workspace(dev, args, ws):
    return ws.get(lib().ntf_fnn_step_workspace_bytes(_lib.ctx(dev), C.byref(args)))


def fnn_step(dev, args, ws, handle=None):
    """one whole Fnn batch (ntf_fnn_step): args is a filled _lib.FnnStepArgs; handle: the caller's own ntf_ctx (its side streams)"""
    h = handle if handle is not None else _lib.ctx(dev)
    p, nb = ws.get(lib().ntf_fnn_step_workspace_bytes(h, C.byref(args)))
    check(lib().ntf_fnn_step(h, _stream(dev), C.byref(args), p, nb), 'ntf_fnn_step')
