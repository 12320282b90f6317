"""World-size-2 `gloo` tests (CPU) of the host-side multi-GPU logic: how a global batch is sliced across ranks, that the
per-rank loss scaling + SUM all-reduce of the flat gradient arena reproduces the single-process gradient (the data-parallel
contract `Engine.step(loss_scale=1/B_global)` + `Engine.optimizer_step` implement on the GPU), that every rank draws the same
epoch order, and the expert-sharded top-k merge rule.  The kernels themselves are covered by the `-m gpu` tests."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import fnn_oracle as O


def _free_port():
    s = socket.socket(); s.bind(('127.0.0.1', 0)); p = s.getsockname()[1]; s.close()
    return p


def _worker(rank, world, port, fn, ret):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        ret[rank] = fn(rank, world)
    finally:
        dist.destroy_process_group()


def run2(fn):
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(2, _free_port(), fn, ret), nprocs=2, join=True)
    return [ret[0], ret[1]]


def _slice(B, G, r):  # Fnn._rank_slice
    per = -(-B // G)
    return min(B, r * per), min(B, (r + 1) * per)


def _dp_gradients(rank, world):
    """each rank: gradient of its slice of the global batch with the loss scaled by 1/B_global, then SUM all-reduce"""
    torch.manual_seed(0)
    rng = np.random.default_rng(0)
    B, S, E = 13, 9, 17  # 13 teams over 2 ranks: slices of 7 and 6
    X = torch.from_numpy((rng.random((B, S)) < 0.3).astype(np.float32))
    y = torch.from_numpy((rng.random((B, E)) < 0.2).astype(np.float32))
    neg = torch.from_numpy(rng.integers(0, E, (B, 3)))
    layers = O.init_params(S, [8], E)
    lo, hi = _slice(B, world, rank)
    logits, acts, pre = O.forward(layers, X[lo:hi])
    w = O.loss_weights(y[lo:hi], neg[lo:hi], 10, 1)
    # oracle.backward divides by ITS batch size: rescale to 1/B_global like Engine.step(loss_scale=1/B_global)
    g = [t * ((hi - lo) / B) for Wb in O.backward(layers, acts, pre, y[lo:hi], w) for t in Wb]
    flat = torch.cat([t.reshape(-1) for t in g])
    loss = O.bce_with_logits(logits, y[lo:hi], w).sum() / B
    dist.all_reduce(flat); dist.all_reduce(loss)
    full_logits, a2, p2 = O.forward(layers, X)
    wf = O.loss_weights(y, neg, 10, 1)
    ref = torch.cat([t.reshape(-1) for Wb in O.backward(layers, a2, p2, y, wf) for t in Wb])
    ref_loss = O.bce_with_logits(full_logits, y, wf).sum(1).mean()
    return float((flat - ref).abs().max() / ref.abs().max()), float(abs(loss - ref_loss) / ref_loss), (lo, hi)


def test_rank_slices_tile_the_global_batch():
    for B in (1, 2, 7, 1000, 1001):
        for G in (1, 2, 4, 8):
            s = [_slice(B, G, r) for r in range(G)]
            assert s[0][0] == 0 and s[-1][1] == B and all(a[1] == b[0] for a, b in zip(s, s[1:]))


def test_data_parallel_gradient_allreduce_equals_single_process():
    out = run2(_dp_gradients)
    assert [o[2] for o in out] == [(0, 7), (7, 13)]
    for g_err, l_err, _ in out: assert g_err < 1e-6 and l_err < 1e-6


def _epoch_order(rank, world):
    from opentf_b200.fnn import loader_order
    torch.manual_seed(0)  # Ntf.__init__ -> set_seed(seed): every rank seeds alike, so every rank shuffles alike
    order = torch.from_numpy(loader_order(torch, 101, True)).to(torch.int64)
    got = [torch.empty_like(order) for _ in range(world)]
    dist.all_gather(got, order)
    return bool(all(torch.equal(g, order) for g in got))


def test_every_rank_draws_the_same_epoch_permutation():
    assert run2(_epoch_order) == [True, True]


def _sharded_topk(rank, world):
    """expert-sharded inference: every rank ranks its expert range, the lists are all-gathered and merged (the rule ntf_topk_merge
    implements on the device: value descending, ties -> lower GLOBAL expert id)"""
    rng = np.random.default_rng(1)
    B, E, K = 5, 40, 4
    P = np.round(rng.random((B, E)), 1).astype(np.float32)  # coarse values: many ties
    per = E // world
    e0 = rank * per
    v, i = O.topk_rows(P[:, e0:e0 + per], K)
    mine = torch.from_numpy(np.concatenate([v, (i + e0).astype(np.float32)], axis=1).astype(np.float32))
    got = [torch.empty_like(mine) for _ in range(world)]
    dist.all_gather(got, mine)
    vals = np.concatenate([g[:, :K].numpy() for g in got], axis=1)
    idx = np.concatenate([g[:, K:].numpy().astype(np.int64) for g in got], axis=1)
    order = np.lexsort((idx, -vals), axis=1)[:, :K]
    mv, mi = np.take_along_axis(vals, order, 1), np.take_along_axis(idx, order, 1)
    rv, ri = O.topk_rows(P, K)
    return bool(np.array_equal(mv, rv) and np.array_equal(mi, ri))


def test_expert_sharded_topk_merge_rule():
    assert run2(_sharded_topk) == [True, True]
