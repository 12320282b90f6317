#!/usr/bin/env python
"""bench.py -- train teams/s of the Fnn hot path on a synthetic DBLP-v12-shaped workload (BASELINE.json configs[1]).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--workload dblp] [--batch 1000] [--precision tf32]

A step = one training batch: input layer, negative sampling, output layer forward + weighted BCE + backward, input
layer backward, Adam.  `value` times K steps with every input resident in HBM (CUDA events, max over ranks);
`e2e` times the same steps through the streaming entry point with HOST batches (pinned H2D of the batch CSR and a
D2H read of the loss inside the timed region).  Under torchrun each rank works on its own slice of a global
batch of N x --batch teams (weak scaling), gradients are all-reduced over NCCL.
`--impl reference` times the reference's own CPU implementation of the same step (the oracle port of
src/mdl/fnn.py, torch CPU, all host threads) on a bounded sample of the same workload.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path: sys.path.insert(0, ROOT)

METRIC, UNIT = 'train teams/s (Fnn, unigram_b, fwd+bwd+Adam)', 'teams/s'
SEEDS = {'dblp': 12, 'imdb': 3, 'uspt': 4, 'toy': 1}


def parse():
    p = argparse.ArgumentParser()
    p.add_argument('--gpus', type=int, default=1)
    p.add_argument('--steps', type=int, default=1000)
    p.add_argument('--warmup', type=int, default=10)
    p.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    p.add_argument('--workload', default='dblp', choices=list(SEEDS))
    p.add_argument('--batch', type=int, default=1000, help='teams per GPU per step (reference default b=1000)')
    p.add_argument('--precision', default='tf32', choices=['tf32', 'fp32'])
    p.add_argument('--nsd', default='unigram_b')
    p.add_argument('--parallel', default='dp', choices=['dp', 'shard'],
                   help="multi-GPU mode: 'dp' data-parallel teams (weak scaling, --batch teams per GPU); 'shard' output layer column-sharded by expert "
                        "(BASELINE configs[3]: every rank runs the same --batch teams on its expert range, strong scaling)")
    p.add_argument('--cpu-baseline-seconds', type=float, default=15.0)
    p.add_argument('--no-cpu-baseline', action='store_true')
    p.add_argument('--infer-k', type=int, default=10)
    p.add_argument('--leg', default='all', choices=['all', 'bnn'], help="'bnn': only the Bnn train leg (its dict is the output line; for profiling)")
    p.add_argument('--no-extras', action='store_true', help='skip the secondary legs (Bnn train on the imdb shape, top-K sweep)')
    p.add_argument('--extras', action='store_true', help='run the secondary legs under torchrun too (default: single-GPU runs only)')
    return p.parse_args()


def workload(name):
    """the synthetic teamsvecs + splits (cached as npz under /tmp: every rank and both arms see the same data)."""
    import scipy.sparse as sp
    from opentf_b200 import synth
    path = f'/tmp/ntf_b200_synth_{name}_{SEEDS[name]}.npz'
    if not os.path.exists(path) and int(os.environ.get('LOCAL_RANK', 0)) != 0:  # one rank of a node generates, the others wait for the file
        for _ in range(1800):
            if os.path.exists(path): break
            time.sleep(1.0)
    if not os.path.exists(path):
        tv = synth.make_teamsvecs(name, seed=SEEDS[name])
        tmp = f'{path}.{os.getpid()}.npz'
        np.savez(tmp, s_ptr=tv['skill'].indptr, s_idx=tv['skill'].indices, s_shape=tv['skill'].shape, m_ptr=tv['member'].indptr,
                 m_idx=tv['member'].indices, m_shape=tv['member'].shape)
        os.replace(tmp, path)
    z = np.load(path)
    mk = lambda p: sp.csr_matrix((np.ones(len(z[f'{p}_idx']), dtype=np.uint8), z[f'{p}_idx'], z[f'{p}_ptr']), shape=tuple(z[f'{p}_shape']))
    tv = {'skill': mk('s'), 'member': mk('m')}
    return tv, synth.make_splits(tv['skill'].shape[0], seed=SEEDS[name])


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed regions (B200_PROFILING.md recipe).  Samples carry
    nvidia-smi's own timestamp; only those inside a window marked with `window()` count as "under load"."""
    Q = 'timestamp,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap'

    def __init__(self, index):
        self.index, self.rows, self.proc, self.windows = index, [], None, []

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(['nvidia-smi', f'--query-gpu={self.Q}', '--format=csv,noheader,nounits', '-lms', '20', '-i', str(self.index)],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=lambda: [self.rows.append(l) for l in self.proc.stdout], daemon=True)
            self.t.start()
        except Exception:
            self.proc = None
        return self

    def __exit__(self, *a):
        if self.proc:
            time.sleep(0.1)
            self.proc.terminate()
            try: self.proc.wait(2)
            except Exception: self.proc.kill()
            self.t.join(1)

    def window(self, t0, t1):
        self.windows.append((t0, t1))

    def summary(self):
        import datetime
        sm, mx, reasons, allsm = [], [], set(), []
        for l in self.rows:
            f = [x.strip() for x in l.split(',')]
            if len(f) < 8: continue
            try:
                ts = datetime.datetime.strptime(f[0], '%Y/%m/%d %H:%M:%S.%f').timestamp()
                c, m = float(f[1]), float(f[2])
            except ValueError: continue
            allsm.append(c)
            if self.windows and not any(a <= ts <= b for a, b in self.windows): continue
            sm.append(c); mx.append(m)
            for name, v in zip(('hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap'), f[4:8]):
                if v.lower().startswith('active'): reasons.add(name)
        if not sm: return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': [], 'samples': 0, 'samples_total': len(allsm)}
        return {'sm_mhz': float(np.median(sm)), 'sm_max_mhz': float(max(mx)), 'reasons': sorted(reasons), 'samples': len(sm), 'samples_total': len(allsm)}


def peaks():
    p = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(p):
        d = json.load(open(p)); d['_source'] = 'measured'
        return d
    return {'hbm_gbs': 6650.0, 'bf16_tflops': 1590.0, 'bf16_tflops_sustained': 1400.0, '_source': 'fallback'}


# ------------------------------------------------------------------------------------------------ reference arm / cpu baseline
def cpu_reference_steps(tv, splits, batch, nsd, max_seconds, min_steps=1, warmup=1, max_steps=10 ** 9):
    """the reference's CPU implementation of one step (oracle port of fnn.py:118-140, incl. its per-row densification),
    on fold-0 train batches of the same workload.  Returns (teams/s, steps timed, seconds, threads)."""
    import torch
    from oracle import fnn_oracle as O
    torch.set_num_threads(os.cpu_count())
    S, E = tv['skill'].shape[1], tv['member'].shape[1]
    rows_all = np.asarray(splits['folds'][0]['train'])
    torch.manual_seed(0)
    layers = O.init_params(S, [128], E)
    flat = [t for Wb in layers for t in Wb]
    opt = O.Adam(flat, 1e-3)
    skill_lil, member_lil = tv['skill'].tolil(), tv['member'].tolil()  # what the reference is handed (team.py:154)

    def one(bi):
        rows = rows_all[(bi * batch) % (len(rows_all) - batch):][:batch]
        # ntf.py:23: one lil row at a time -> csr -> dense -> float, then the default collate stacks them
        X = torch.stack([torch.as_tensor(skill_lil[int(r)].tocsr().toarray()).float() for r in rows]).squeeze(1)
        y = torch.stack([torch.as_tensor(member_lil[int(r)].tocsr().toarray()).float() for r in rows]).squeeze(1)
        logits, acts, pre = O.forward(layers, X)
        neg = O.sample_negatives(y, nsd, 5, None)
        w = O.loss_weights(y, neg, 10, 1)
        loss = O.bce_with_logits(logits, y, w).sum(dim=1).mean()
        opt.step([g for Wb in O.backward(layers, acts, pre, y, w) for g in Wb])
        return loss.item()

    for i in range(warmup): one(i)
    t0, n = time.perf_counter(), 0
    while n < max_steps and (n < min_steps or time.perf_counter() - t0 < max_seconds):
        one(warmup + n); n += 1
    dt = time.perf_counter() - t0
    return n * batch / dt, n, dt, torch.get_num_threads()


def verbatim_reference_steps(tv, splits, batch, nsd, steps, warmup):
    """the UNMODIFIED reference (src/mdl/fnn.py `Fnn.learn`, placed under oracle/_ref by oracle/make_ref.py, imported through the
    environment shims of oracle/ref_shim.py) on the CPU with all host threads: one epoch over (warmup + steps) fold-0 train batches of
    the same workload, through its own DataLoader / NtfDataset (per-row lil -> dense, ntf.py:17-25), bxe, autograd and torch.optim.Adam.
    Step boundaries come from a global optimizer-step post hook (the reference calls optimizer.step() once per batch, fnn.py:139), so
    nothing of the reference is edited.  Returns (teams/s over the timed steps, steps, seconds, threads)."""
    import tempfile
    import torch
    from oracle import ref_shim
    torch.set_num_threads(os.cpu_count())
    Fnn = ref_shim.load_reference_fnn()
    rows = np.asarray(splits['folds'][0]['train'])[:(warmup + steps) * batch]
    assert len(rows) == (warmup + steps) * batch, 'workload smaller than the requested sample'
    sub = {'test': np.asarray(splits['test'])[:1], 'folds': {0: {'train': rows, 'valid': np.asarray(splits['folds'][0]['valid'])[:min(batch, 64)]}}}
    teamsvecs = {'skill': tv['skill'].tolil(), 'member': tv['member'].tolil()}  # what the reference is handed (team.py:154)
    stamps = []
    from torch.optim.optimizer import register_optimizer_step_post_hook
    handle = register_optimizer_step_post_hook(lambda *a, **k: stamps.append(time.perf_counter()))
    try:
        with tempfile.TemporaryDirectory() as out:
            f = Fnn(out, 'cpu', 0, ref_shim.default_cfg(b=batch, e=1, nsd=nsd, spe=0))
            f.learn(teamsvecs, sub, None)
    finally:
        handle.remove()
    assert len(stamps) == warmup + steps, (len(stamps), warmup, steps)
    dt = stamps[-1] - stamps[warmup - 1] if warmup > 0 else None
    assert dt is not None, 'the reference arm needs at least one warm-up step (its end is the start of the timed region)'
    return steps * batch / dt, steps, dt, torch.get_num_threads()


def reference_arm(tv, splits, batch, nsd, steps, warmup):
    """-> (teams/s, steps, seconds, threads, kind, what): the verbatim reference when oracle/_ref (or /root/reference) is there, else the oracle port"""
    from oracle import ref_shim
    if ref_shim.available():
        v, n, dt, th = verbatim_reference_steps(tv, splits, batch, nsd, steps, max(1, warmup))
        return v, n, dt, th, 'reference', ('UNMODIFIED reference src/mdl/fnn.py Fnn.learn (oracle/_ref via oracle/ref_shim.py): DataLoader + per-row lil->dense, bxe, '
                                           'autograd, Adam, loss.item(); step boundaries from an optimizer-step hook')
    v, n, dt, th = cpu_reference_steps(tv, splits, batch, nsd, 1e9, min_steps=steps, warmup=warmup, max_steps=steps)
    return v, n, dt, th, 'port', 'oracle port of fnn.py:118-140 incl. per-row densification (oracle/_ref absent)'


def run_reference(args):
    rank = int(os.environ.get('RANK', 0))
    if rank != 0: return  # rank 0 alone runs the CPU arm
    tv, splits = workload(args.workload)
    steps = max(1, min(args.steps, 40))  # bounded sample: ~1.3 s per step of b=1000 on 16 cores
    v, n, dt, threads, kind, what = reference_arm(tv, splits, args.batch, args.nsd, steps, args.warmup)
    sample = f'{n} steps of b={args.batch} on fold-0 train rows of the {args.workload}-shaped workload after {max(1, args.warmup)} warm-up steps; {what}'
    print(json.dumps({
        'impl': 'reference', 'metric': METRIC, 'value': v, 'unit': UNIT, 'n_gpus': args.gpus, 'steps': n, 'warmup': args.warmup,
        'ms_per_step': 1e3 * dt / n, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
        'config': config_of(args, tv), 'cpu_baseline': {'value': v, 'unit': UNIT, 'cores': threads, 'kind': kind, 'sample': sample},
        'e2e': {'value': v, 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}, 'gpu_launches': 0}))


def config_of(args, tv):
    N, S = tv['skill'].shape; E = tv['member'].shape[1]
    cfgno = {'dblp': 1, 'imdb': 2, 'uspt': 3}.get(args.workload, 1)
    shard = getattr(args, 'parallel', 'dp') == 'shard'
    return {'workload': f'{args.workload}-shaped synthetic teamsvecs (BASELINE configs[{cfgno}]): N={N} S={S} E={E}, Fnn h=[128], nsd={args.nsd}, ns=5, tpw=10, tnw=1',
            'batch_per_gpu': args.batch, 'global_batch': args.batch * (1 if shard else args.gpus),
            'parallelism': (f'shard{args.gpus} (output layer column-sharded by expert, {-(-E // args.gpus)} experts per GPU; layer 0 replicated)' if shard else f'dp{args.gpus}'),
            'precision': args.precision,
            'l2_policy': 'no flush: one step touches params+grads+Adam state (>140 MB) and the [B,E] work set, larger than the 126 MB L2'}


# ------------------------------------------------------------------------------------------------ our arm
def run_ours(args):
    import torch
    import torch.distributed as dist
    from opentf_b200 import _lib, ops
    from opentf_b200.engine import Engine, to_csr

    world, rank, local = int(os.environ.get('WORLD_SIZE', 1)), int(os.environ.get('RANK', 0)), int(os.environ.get('LOCAL_RANK', 0))
    torch.cuda.set_device(local)
    dev = torch.device(f'cuda:{local}')
    if world > 1: dist.init_process_group('nccl', device_id=dev)
    if args.leg == 'bnn':
        def sync0():
            torch.cuda.synchronize()
            if world > 1: dist.barrier()
            torch.cuda.synchronize()
        r = bnn_leg(args, dev, world, rank, sync0, dist)
        if rank == 0: print(json.dumps(r))
        if world > 1: dist.destroy_process_group()
        return
    tv, splits = workload(args.workload)
    N, S = tv['skill'].shape; E = tv['member'].shape[1]
    b, G = args.batch, world
    shard = args.parallel == 'shard'
    eng = Engine(S, [128], E, dev, precision=args.precision, tpw=10, tnw=1, nsd=args.nsd, ns=5, seed=0, max_batch=b, shard=(rank, world) if shard else None)
    eng.world, eng.rank = world, rank
    xmode = os.environ.get('NTF_DP_EXCHANGE', 'peer')  # how data-parallel ranks exchange gradients (opentf_b200/fnn.py: Fnn.init)
    if shard:
        if world > 1 and xmode == 'peer': eng.attach_shard_peers()  # the dA exchange inside the step, over peer memory
    elif world > 1 and xmode == 'peer': eng.attach_peers()
    elif world > 1 and xmode == 'nccl': eng.attach_comm()
    eng.stage(tv['skill'], tv['member'])
    torch.manual_seed(0)
    lin = [torch.nn.Linear(S, 128), torch.nn.Linear(128, E)]
    for m in lin: torch.nn.init.xavier_uniform_(m.weight)
    eng.load_state_dict({f'layers.{i}.{n}': getattr(m, n).detach() for i, m in enumerate(lin) for n in ('weight', 'bias')})
    train_rows = np.asarray(splits['folds'][0]['train'])
    sp = eng.split(train_rows[np.random.default_rng(0).permutation(len(train_rows))])
    gB = b if shard else b * G  # sharded layer: every rank runs the SAME batch on its expert range
    nb = min(sp.n // gB, 64)
    assert nb >= 1, 'workload smaller than one global batch'
    precision_used = 'tf32' if eng.precision == _lib.NTF_TF32 else 'fp32'

    def device_step(i):
        g0 = (i % nb) * gB
        eng.step(sp, g0 + (0 if shard else rank * b), b, True, lr=1e-3, loss_slot=i % nb, loss_scale=1.0 / gB, gbatch=(g0, gB))

    def sync():
        torch.cuda.synchronize()
        if world > 1: dist.barrier()
        torch.cuda.synchronize()

    # ---- kernel-resident timing: `value`, then the same steps again with an event pair around the output-layer call: `roofline` ----
    # Graph mode (default): every batch of the split is captured once per pass.  The event pairs are external event-record nodes INSIDE the
    # step graphs; they cost the step ~13 us (measured: 151 vs 138 us, scripts/e2e_host_profile.py) because nothing overlaps across them, so
    # `value` is timed on graphs without them and the dominant kernel group is timed in a second pass of the same K steps, live, with the
    # pairs re-recorded by every replay (after the loop each pair holds the duration of its last replay).  NTF_GRAPHS=0: one pair per step,
    # recorded by ntf_fnn_step as it enqueues.
    graphs = eng.use_graphs
    pairs = []

    def new_pair():
        p = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
        for e in p: e.record()  # (creates the underlying cudaEvent_t handles)
        pairs.append(p)
        return p

    def timed_pass(with_pairs):
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        for i in range(args.warmup): device_step(i)
        sync()
        _lib.lib().ntf_launch_count(1)
        w0 = time.time()
        ev0.record()
        h0 = time.perf_counter()
        for i in range(args.steps):
            if with_pairs and not graphs: eng.prof_events = pairs[i]  # ntf_fnn_step records them around its output-layer call, on the launching stream
            device_step(args.warmup + i)
        eng.prof_events = None
        host = (time.perf_counter() - h0) * 1e3 / args.steps  # CPU time to ENQUEUE one step (if ~ ms_per_step the loop is launch-bound)
        ev1.record()
        sync()
        clk.window(w0, time.time())
        n_launch = int(_lib.lib().ntf_launch_count(0))
        t = torch.tensor([ev0.elapsed_time(ev1)], device=dev)
        if world > 1: dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item()), host, n_launch

    if graphs:
        for i in range(nb): device_step(i)  # setup pass: capture (no event pairs)
    torch.cuda.synchronize()
    clk = ClockSampler(local).__enter__()  # samples through the timed legs (device-resident, roofline pass, end-to-end)
    time.sleep(0.3)                        # (the sampler's first reading) ... then the warm-up steps, so that the timed ones do not start on an idle GPU
    ms, host_ms, launches = timed_pass(False)
    value = args.steps * gB / (ms * 1e-3)
    if graphs:
        eng.clear_graphs()
        eng.graph_event_factory = lambda: tuple(e.cuda_event for e in new_pair())
        for i in range(nb): device_step(i)  # capture again, with a pair around the output-layer call
        eng.graph_event_factory = None
    else:
        for _ in range(args.steps): new_pair()
    ms_k, _, _ = timed_pass(True)
    used = pairs if (not graphs or args.steps >= nb) else [pairs[(args.warmup + i) % nb] for i in range(args.steps)]
    k_ms = float(np.mean([a.elapsed_time(c) for a, c in used]))

    # ---- end to end through the streaming entry point: host batches in, loss out ----
    # Every step: the global batch's rows are PACKED from the host teamsvecs CSR into a pinned block (ntf_pack_host_batch, the loader's work),
    # copied to the device (one H2D transfer), stepped, and the loss of every step is copied back and read by the host.  Packing and enqueueing of
    # batch i+1 run on the host while the GPU works on batch i, as a loader thread would; all of it is inside the timed region.
    host = HostBatches(tv, train_rows, b, 0 if shard else rank, 1 if shard else G)
    for i in range(min(3, args.warmup)): host.step(eng, i)
    host.run(eng, 3, max(4, args.warmup))  # (both loss slots of the two-in-flight loop: their step graphs are captured here, not in the timed region)
    sync()
    w0 = time.time()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    host.run(eng, 100, args.steps)
    e1.record()
    sync()
    e2e_ms = max(e0.elapsed_time(e1), 0.0)
    clk.window(w0, time.time())
    clk.__exit__()
    t = torch.tensor([e2e_ms], device=dev)
    if world > 1: dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_value = args.steps * gB / (float(t.item()) * 1e-3)

    # ---- top-K inference teams/s (second half of BASELINE's metric), device resident ----
    test_sp = eng.split(np.asarray(splits['test']))
    ib = min(b, test_sp.n)
    scores = torch.empty(ib, E, device=dev)
    vals, idx = torch.empty(ib, args.infer_k, device=dev), torch.empty(ib, args.infer_k, dtype=torch.int32, device=dev)
    for _ in range(3): eng.topk(test_sp, 0, ib, args.infer_k, scores, vals, idx)
    sync()
    i0, i1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 50
    i0.record()
    for r in range(reps): eng.topk(test_sp, (r * ib) % max(1, test_sp.n - ib + 1), ib, args.infer_k, scores, vals, idx)
    i1.record()
    sync()
    infer_value = reps * ib * (1 if shard else G) / (i0.elapsed_time(i1) * 1e-3)

    E_local = eng.E  # (the engine is released before the secondary legs)
    extras = {}
    # secondary legs: on one GPU by default; under torchrun only when asked for (--extras) -- the headline line of a scaling run should not
    # depend on them (a leg that fails on one rank would leave the others waiting in a collective).  `--leg bnn` runs the Bnn leg alone at any N.
    if not args.no_extras and (world == 1 or args.extras):
        def leg(name, fn):
            try: extras[name] = fn()
            except Exception as ex:  # (single process: report and go on; the headline measurement above is already taken)
                if world > 1: raise
                extras[name] = {'error': f'{type(ex).__name__}: {ex}'}
        lin_sd = {f'layers.{i}.{n}': getattr(m, n).detach() for i, m in enumerate(lin) for n in ('weight', 'bias')}
        if not shard: leg('kernel_rooflines', lambda: kernel_rooflines(eng, sp, b, dev, sync, peaks()))
        del eng, sp, test_sp, scores
        torch.cuda.empty_cache()
        if not shard: leg('infer_topk_sweep', lambda: topk_sweep(args, tv, splits, dev, sync, G, lin_sd))
        leg('bnn_train', lambda: bnn_leg(args, dev, world, rank, sync, dist))
        leg('batch_sweep', lambda: batch_sweep(args, tv, splits, dev, world, rank, sync, dist))
        if rank == 0: leg('staging', lambda: staging_leg(args, tv, splits, dev, sync, peaks()))
    if rank != 0:
        if world > 1: dist.destroy_process_group()
        return
    pk = peaks()
    flops = 6.0 * 128 * E_local * b  # K3 flops/team (Fnn train) = 6*h_L*E, per launch of b teams (sharded: this rank's experts)
    traffic = None
    tpath = os.path.join(ROOT, 'profiles', 'out_tc2_traffic.json')
    if precision_used == 'tf32' and args.workload == 'dblp' and b == 1000 and os.path.exists(tpath):
        tj = json.load(open(tpath)); traffic = tj['dram_bytes_read'] + tj['dram_bytes_write']  # one ncu --set full capture, per launch
    # the tensor-core kernel issues kind::f16 MMAs (fp16 operands with TF32's 10-bit mantissa, fp32 accumulate): the denominator is the
    # measured dense bf16/fp16 GEMM rate, sustained figure (the kernel is timed inside a long step)
    roof = {'bound': 'tensor', 'kernel': f'ntf_out_train[{precision_used}] (output layer fwd + weighted BCE + bwd)',
            'achieved': flops / (k_ms * 1e-3) / 1e12, 'peak': pk['bf16_tflops'], 'unit': 'TFLOP/s', 'traffic': traffic,
            'peak_source': (f"{pk['_source']} cuBLAS bf16 BURST rate (MEASURED_PEAKS.json: bf16_tflops; the kernel group is timed by itself with events around it)"
                            if pk['_source'] == 'measured' else 'fallback 1.59 PFLOP/s (B200_PROFILING.md)'),
            'frac_of_sustained_peak': flops / (k_ms * 1e-3) / 1e12 / pk['bf16_tflops_sustained'],
            'avg_launch_ms': k_ms, 'share_of_step': k_ms * args.steps / ms_k, 'ms_per_step_with_event_pairs': ms_k / args.steps,
            'timed': 'a second pass of the same W + K steps with an event pair around the output-layer call inside every step graph (external event-record nodes); `value` is timed on graphs without them: nothing overlaps across such a node and the step is ~10 us slower with them',
            'note': 'persistent dense pass (out_tc2_kernel) + sparse correction pass (out_fix_kernel); h=128: per logit 768 tensor flops vs ~10 issue slots + 1.25 MUFU ops of epilogue: the epilogue binds before the tensor pipe (DESIGN.md 4.1)'}
    roof['frac'] = roof['achieved'] / roof['peak']
    out = {'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': G, 'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': ms / args.steps,
           'higher_is_better': True, 'scaling': 'strong' if shard else 'weak', 'vs_baseline': None, 'dtype': 'f32' if precision_used == 'fp32' else 'f16 operands (10-bit mantissa, TF32-class), f32 accumulate',
           'data': 'synthetic', 'config': config_of(args, tv), 'clocks': clk.summary(),
           'e2e': {'value': e2e_value, 'unit': UNIT, 'h2d_bytes_per_step': host.h2d_bytes, 'd2h_bytes_per_step': 4,
                   'api': 'HostPacker.pack (ntf_pack_host_batch: rows of the host teamsvecs CSR -> pinned block; batch i+1 packed while the GPU works on batch i) '
                          '-> Engine.step_host: one H2D copy -> ntf_fnn_step (replayed as a CUDA graph) -> the loss of every step copied back and read by the host (two steps in flight: the host packs and enqueues batch i+1 while the GPU works on batch i, then waits for loss i; the copies run on the copy stream of the library); all inside the timed region'},
           'gpu_launches': launches, 'cuda_graphs': bool(graphs),
           'dp_exchange': None if G == 1 else {'peer': 'reduce-scatter + Adam + all-gather fused in one pass over peer memory inside ntf_fnn_step (csrc/peer.cu), 2 overlapped arena segments',
                                               'nccl': 'ncclAllReduce inside ntf_fnn_step: 2 overlapped arena segments, captured in the step graph'}.get(
                                                   xmode, 'torch.distributed.all_reduce between the halves of a step'), 'host_enqueue_ms_per_step': host_ms, 'roofline': roof, 'infer_topk': {'k': args.infer_k, 'value': infer_value, 'unit': 'teams/s', 'batch': ib}}
    out.update(extras)
    if not args.no_cpu_baseline:
        v, n, dt, threads, kind, what = reference_arm(tv, splits, b, args.nsd, max(1, int(args.cpu_baseline_seconds / 1.3)), 1)
        out['cpu_baseline'] = {'value': v, 'unit': UNIT, 'cores': threads, 'kind': kind, 'sample': f'{n} steps of b={b} ({dt:.1f} s) of the same workload; {what}'}
    print(json.dumps(out))
    if world > 1: dist.destroy_process_group()


def bnn_leg(args, dev, world, rank, sync, dist):
    """BASELINE configs[2]: Bnn (Flipout, one weight sample per step, KL/B in the loss) on the IMDB-shaped workload, data-parallel teams
    (weak scaling, b teams per GPU): train teams/s with everything resident, CUDA events, max over ranks."""
    import torch
    from opentf_b200 import _lib
    from opentf_b200.engine import Engine
    tv, splits = workload('imdb')
    N, S = tv['skill'].shape; E = tv['member'].shape[1]
    b, gB = args.batch, args.batch * world
    eng = Engine(S, [128], E, dev, bayesian=True, precision=args.precision, tpw=10, tnw=1, nsd=args.nsd, ns=5, seed=0, max_batch=b)
    eng.world, eng.rank = world, rank
    if world > 1 and os.environ.get('NTF_DP_EXCHANGE', 'peer') == 'peer': eng.attach_peers()
    eng.stage(tv['skill'], tv['member'])
    g = torch.Generator().manual_seed(0)
    sd = {}
    for i, (fin, fout) in enumerate(((S, 128), (128, E))):  # bnn.py:19-25: mu ~ N(0,.1), rho ~ N(-3,.1)
        sd[f'layers.{i}.mu_weight'] = torch.empty(fout, fin).normal_(0.0, 0.1, generator=g); sd[f'layers.{i}.rho_weight'] = torch.empty(fout, fin).normal_(-3.0, 0.1, generator=g)
        sd[f'layers.{i}.mu_bias'] = torch.empty(fout).normal_(0.0, 0.1, generator=g); sd[f'layers.{i}.rho_bias'] = torch.empty(fout).normal_(-3.0, 0.1, generator=g)
    eng.load_state_dict(sd)
    rows = np.asarray(splits['folds'][0]['train'])
    sp = eng.split(rows[np.random.default_rng(0).permutation(len(rows))])
    nb = sp.n // gB
    steps = max(1, min(args.steps, 100))

    def one(i):
        g0 = (i % nb) * gB
        eng.step(sp, g0 + rank * b, b, True, lr=1e-3, loss_slot=i % nb, loss_scale=1.0 / gB, gbatch=(g0, gB))

    for i in range(max(3, min(args.warmup, 5))): one(i)
    if eng.use_graphs:  # setup pass (untimed, like an epoch 0): every batch the timed loop visits is captured once; the timed loop replays
        for i in range(min(nb, steps)): one(5 + i)
    sync()
    _lib.lib().ntf_launch_count(1)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(steps): one(5 + i)
    e1.record()
    sync()
    launches = int(_lib.lib().ntf_launch_count(0))
    t = torch.tensor([e0.elapsed_time(e1)], device=dev)
    if world > 1: dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    loss = float(eng.loss_buf[(5 + steps - 1) % nb].item())
    prec = 'tf32' if eng.precision == _lib.NTF_TF32 else 'fp32'
    flops = 12.0 * 128 * E * b  # K3 flops/team (Bnn-Flipout train) = 12*h_L*E
    return {'metric': 'train teams/s (Bnn Flipout, 1 weight sample/step, unigram_b, fwd+bwd+KL+Adam)', 'value': steps * gB / (ms * 1e-3), 'unit': UNIT,
            'workload': f'imdb-shaped synthetic teamsvecs (BASELINE configs[2]): N={N} S={S} E={E}, Bnn h=[128], nsd={args.nsd}, ns=5', 'batch_per_gpu': b,
            'steps': steps, 'ms_per_step': ms / steps, 'gpu_launches_per_step': launches / steps, 'precision': prec, 'last_loss': loss, 'cuda_graphs': bool(eng.use_graphs),
            'output_layer_tflops_if_alone': flops / (ms / steps * 1e-3) / 1e12}


def batch_sweep(args, tv, splits, dev, world, rank, sync, dist):
    """the same Fnn workload at larger batches (b = 4096, 16384 teams per GPU).  At the reference's default b=1000 the dense
    Adam pass over 8.96 M parameters (251 MB of HBM traffic per step) is a fixed cost per STEP; larger batches amortise it."""
    import torch
    from opentf_b200 import _lib
    from opentf_b200.engine import Engine
    N, S = tv['skill'].shape; E = tv['member'].shape[1]
    rows = np.asarray(splits['folds'][0]['train'])
    out = []
    for b in (4096, 16384):
        gB = b * world
        if len(rows) < gB: continue
        eng = Engine(S, [128], E, dev, precision=args.precision, tpw=10, tnw=1, nsd=args.nsd, ns=5, seed=0, max_batch=b)
        eng.world, eng.rank = world, rank
        if world > 1 and os.environ.get('NTF_DP_EXCHANGE', 'peer') == 'peer': eng.attach_peers()
        eng.stage(tv['skill'], tv['member'])
        torch.manual_seed(0)
        lin = [torch.nn.Linear(S, 128), torch.nn.Linear(128, E)]
        for m in lin: torch.nn.init.xavier_uniform_(m.weight)
        eng.load_state_dict({f'layers.{i}.{n}': getattr(m, n).detach() for i, m in enumerate(lin) for n in ('weight', 'bias')})
        sp = eng.split(rows[np.random.default_rng(0).permutation(len(rows))])
        nb = sp.n // gB
        steps = max(1, min(args.steps, 100))

        def one(i):
            g0 = (i % nb) * gB
            eng.step(sp, g0 + rank * b, b, True, lr=1e-3, loss_slot=i % nb, loss_scale=1.0 / gB, gbatch=(g0, gB))

        for i in range(nb + 3): one(i)  # (captures every batch's graph, then warm)
        sync()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(steps): one(i)
        e1.record()
        sync()
        t = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1: dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
        out.append({'batch_per_gpu': b, 'value': steps * gB / (ms * 1e-3), 'unit': UNIT, 'ms_per_step': ms / steps, 'steps': steps,
                    'precision': 'tf32' if eng.precision == _lib.NTF_TF32 else 'fp32', 'output_layer_tflops_if_alone': 6.0 * 128 * E * b / (ms / steps * 1e-3) / 1e12})
        del eng, sp
        torch.cuda.empty_cache()
    return out


def staging_leg(args, tv, splits, dev, sync, pk):
    """SURVEY 8(f-4) on the bench workload's teamsvecs: member^T . skill (team.py:302-341, test teams skipped), the multi-hot rows from id lists
    (team.py:148-173) and the skill-coverage loop (metric.py:44-73), device resident, CUDA events; beside each the reference's own CPU code path
    (oracle/staging_oracle.py: scipy's product / the per-team loop) on a bounded sample on this host.  HBM/L2-bound integer work: algorithmic bytes
    over time over the measured copy bandwidth."""
    import time, numpy as np, scipy.sparse as sp_, torch
    from opentf_b200 import ops
    from oracle import staging_oracle as SO
    M, S = sp_.csr_matrix(tv['member']), sp_.csr_matrix(tv['skill'])
    M.sort_indices(); S.sort_indices()
    T, E = M.shape; Sn = S.shape[1]
    i32 = lambda a: torch.as_tensor(np.ascontiguousarray(a, dtype=np.int32), device=dev)
    mp, mi, sp0, si = i32(M.indptr), i32(M.indices), i32(S.indptr), i32(S.indices)
    flags = np.zeros(T, np.uint8); flags[np.asarray(splits['test'])] = 1
    skip = torch.as_tensor(flags, device=dev)
    ws = ops.Workspace(dev)

    def timed(fn, reps):
        for _ in range(2): r = fn()
        sync()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps): r = fn()
        e1.record()
        sync()
        return e0.elapsed_time(e1) / reps * 1e-3, r

    out = {}
    # --- member^T . skill
    t, (cp, ci, cv) = timed(lambda: ops.cooccur(mp, mi, sp0, si, E, Sn, skip, ws), 5)
    n_s = np.diff(S.indptr).astype(np.int64)
    keep = flags == 0
    pairs = int((np.diff(M.indptr).astype(np.int64) * n_s)[keep].sum())  # (team of expert, skill) visits
    nnz_m = int(np.diff(M.indptr)[keep].sum())
    nbytes = 16 * nnz_m + 2 * 2 * 4 * pairs + 8 * int(ci.numel())  # transpose (count + scatter: 8 B each per member entry), two sweeps x two passes over the pairs, the output
    t0 = time.perf_counter()
    sample = min(T, 20000)  # the scipy product is superlinear in nothing: a prefix of the teams, scaled by its own pair count
    co_cpu = SO.cooccurrence(M[:sample], S[:sample], [r for r in np.asarray(splits['test']) if r < sample])
    dt = time.perf_counter() - t0
    pairs_cpu = int((np.diff(M.indptr).astype(np.int64) * n_s)[:sample][keep[:sample]].sum())
    out['cooccurrence'] = {'what': f'member^T . skill [{E} x {Sn}] over {T} teams, {len(splits["test"])} test teams skipped', 'ms': t * 1e3, 'pairs': pairs, 'nnz': int(ci.numel()),
                           'value': pairs / t, 'unit': '(expert, skill) pairs/s', 'algorithmic_bytes': nbytes, 'achieved_gbs': nbytes / t / 1e9, 'peak_gbs': pk['hbm_gbs'],
                           'frac': nbytes / t / 1e9 / pk['hbm_gbs'], 'bound': 'hbm/l2 + shared-memory atomics',
                           'cpu_baseline': {'value': pairs_cpu / dt, 'unit': '(expert, skill) pairs/s', 'cores': 1, 'kind': 'port (oracle/staging_oracle.py: the statements of team.py:325-335 -- lil copies, rows emptied, scipy product)',
                                            'sample': f'the first {sample} teams ({dt:.1f} s)'}}
    # --- id lists -> multi-hot rows (both matrices of teamsvecs, the lists = the rows shuffled)
    ids_m, ids_s = i32(M.indices[::-1].copy()), i32(S.indices[::-1].copy())
    ptr_m, ptr_s = i32(M.indptr[-1] - M.indptr[::-1]), i32(S.indptr[-1] - S.indptr[::-1])  # (rows reversed, every row's ids descending)
    t, _ = timed(lambda: (ops.csr_from_lists(ptr_m, ids_m, E, ws), ops.csr_from_lists(ptr_s, ids_s, Sn, ws)), 5)
    n_ids = int(M.indices.shape[0] + S.indices.shape[0])
    t0 = time.perf_counter()
    sample = min(T, 2000)
    SO.rows_from_lists(M.indptr[:sample + 1], M.indices, E); SO.rows_from_lists(S.indptr[:sample + 1], S.indices, Sn)
    dt = time.perf_counter() - t0
    out['rows_from_lists'] = {'what': f'skill and member id lists of {T} teams -> sorted duplicate-free CSR rows', 'ms': t * 1e3, 'value': T / t, 'unit': 'teams/s',
                              'algorithmic_bytes': 16 * n_ids + 16 * T, 'achieved_gbs': (16 * n_ids + 16 * T) / t / 1e9, 'peak_gbs': pk['hbm_gbs'],
                              'frac': (16 * n_ids + 16 * T) / t / 1e9 / pk['hbm_gbs'], 'bound': 'latency (a warp per team, 8 ids per row)',
                              'cpu_baseline': {'value': sample / dt, 'unit': 'teams/s', 'cores': 1, 'kind': 'port (team.py:17-38,148-173: a dense one-hot row per team into a lil_matrix)',
                                               'sample': f'{sample} teams ({dt:.1f} s)'}}
    # --- skill coverage of the test teams' top-10 (random distinct scores stand in for the model's)
    test = np.asarray(splits['test'])[:4096]
    X = S[test]; X.sort_indices()
    K = 10
    idx = torch.stack([torch.randperm(E, device=dev)[:K] for _ in range(64)]).to(torch.int32).repeat((len(test) + 63) // 64, 1)[:len(test)].contiguous()
    cov = torch.empty(len(test), 3, dtype=torch.float64, device=dev)
    xp, xi = i32(X.indptr), i32(X.indices)
    t, _ = timed(lambda: ops.skill_coverage(idx, xp, xi, cp, ci, [2, 5, 10], cov), 20)
    co_host = sp_.csr_matrix((np.minimum(cv.cpu().numpy(), 255).astype(np.uint8), ci.cpu().numpy(), cp.cpu().numpy()), shape=(E, Sn))
    sample = 200
    Yd = np.zeros((sample, E), np.float32)
    ih = idx[:sample].cpu().numpy()
    for r in range(sample): Yd[r, ih[r]] = np.arange(K, 0, -1)
    t0 = time.perf_counter()
    want = SO.skill_coverage(X[:sample], Yd, co_host, '2,5,10')
    dt = time.perf_counter() - t0
    got = cov[:sample].cpu().numpy()
    ok = all(np.array_equal(got[:, j], want[k]) for j, k in enumerate((2, 5, 10)))
    out['skill_coverage'] = {'what': f'coverage@2,5,10 of {len(test)} test teams', 'ms': t * 1e3, 'value': len(test) / t, 'unit': 'teams/s', 'equals_cpu_sample': bool(ok),
                             'bound': 'latency (binary searches in the co-occurrence rows)',
                             'cpu_baseline': {'value': sample / dt, 'unit': 'teams/s', 'cores': 1, 'kind': 'port (metric.py:53-69 line by line)', 'sample': f'{sample} teams ({dt:.1f} s)'}}
    return out


def kernel_rooflines(eng, sp, b, dev, sync, pk):
    """the HBM-bound kernels of a step one by one (eager launches of the library's entry points on the engine's own buffers, CUDA
    events, warm): algorithmic bytes per launch / time / measured copy bandwidth.  At b=1000 a launch moves a few MB -- less than a microsecond
    of HBM time -- so these kernels are bound by launch + dependent-load latency, not bandwidth; the Adam pass (251 MB) is the one at the roof."""
    import torch
    from opentf_b200 import ops
    from opentf_b200._lib import NSD
    S, E, h = eng.S, eng.E, eng.hidden[0]
    ptr = sp.s_indptr[:b + 1].cpu().numpy()
    nnz = int(ptr[-1] - ptr[0]); uniq = int(torch.unique(sp.s_indices[int(ptr[0]):int(ptr[-1])]).numel())
    W0, b0 = eng.view('layers.0.weight'), eng.view('layers.0.bias')
    gW0 = eng.view('layers.0.weight', eng.grads)

    def timed(fn, reps=50):
        for _ in range(3): fn()
        sync()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps): fn()
        e1.record()
        sync()
        return e0.elapsed_time(e1) / reps * 1e-3

    peak = pk['hbm_gbs']
    legs = {
        'K1 csr_bag_fwd': (lambda: ops.csr_bag_fwd(b, sp.s_indptr.data_ptr(), sp.s_indices, W0, b0, S, h, eng.act[0]), nnz * (4 * h + 4) + b * (8 + 4 * h)),
        'K2 csr_bag_bwd (fill + hot list + reduce || hot + combine)': (lambda: ops.csr_bag_bwd(b, sp.s_indptr.data_ptr(), sp.s_indices, sp.s_ent_row, 0, eng.dz[0], S, h, gW0, eng.ws),
                                                                        nnz * (4 * h + 4) + b * 4 * h + uniq * 4 * h),
        'K4 neg_sample (unigram_b)': (lambda: ops.neg_sample(NSD['unigram_b'], 0, 1, 0, b, sp.m_indptr.data_ptr(), sp.m_indices, E, eng.ns, None, eng.neg, sp.m_indptr.data_ptr(), b), None),
        'K6 adam (whole arena)': (lambda: ops.adam_step(eng.params, eng.grads, eng.adam_m, eng.adam_v, eng.n_params, 0.0, 0.9, 0.999, 1e-8, 1), 28 * eng.n_params),
    }
    out = []
    for name, (fn, nbytes) in legs.items():
        t = timed(fn)
        row = {'kernel': name, 'us_per_launch': t * 1e6, 'batch': b}
        if nbytes is not None:
            row.update({'algorithmic_bytes': int(nbytes), 'achieved_gbs': nbytes / t / 1e9, 'peak_gbs': peak, 'frac': nbytes / t / 1e9 / peak, 'bound': 'hbm'})
        else: row.update({'bound': 'latency (binary searches in the batch CSR: O(ns log) per team, no bulk traffic)'})
        out.append(row)
    return out


def topk_sweep(args, tv, splits, dev, sync, G, lin_sd):
    """BASELINE configs[4]: top-K expert ranking over all experts, K x batch sweep (device resident; team-parallel: every rank ranks its own
    batches, the aggregate is G x one rank's rate).  K <= 1024 with 32 K <= E runs the fused kernels (the [B,E] scores never reach HBM) -- that includes
    K = 1000, the reference's default testcfg.topK; the `fused` flag of every row says which path ran.  Small batches are latency-bound (one library call per batch: ~6
    launches); the teams/s at batch >= 4096 is the throughput figure."""
    import torch
    from opentf_b200.engine import Engine
    N, S = tv['skill'].shape; E = tv['member'].shape[1]
    rows = np.asarray(splits['folds'][0]['train'])
    batches = [b for b in (1, 8, 64, 512, 4096, 32768) if b <= len(rows)]
    Bmax = max(batches)
    eng = Engine(S, [128], E, dev, precision=args.precision, nsd=args.nsd, ns=5, seed=0, max_batch=Bmax)
    eng.stage(tv['skill'], tv['member'])
    eng.load_state_dict(lin_sd)
    sp = eng.split(rows[:max(Bmax, min(len(rows), 40000))])
    out = []
    scores = None
    for K in (2, 5, 10, 20, 50, 100, 1000):
        if K > E: continue
        for ib in batches:
            fused = eng.fused_topk_ok(ib, K)
            if not fused and (scores is None or scores.shape[0] < ib): scores = torch.empty(ib, E, device=dev)
            vals, idx = torch.empty(ib, K, device=dev), torch.empty(ib, K, dtype=torch.int32, device=dev)
            span = max(1, sp.n - ib + 1)
            for r in range(2): eng.topk(sp, (r * ib) % span, ib, K, scores, vals, idx)
            sync()
            reps = int(max(3, min(200, 2e5 // max(ib, 256))))
            i0, i1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            i0.record()
            for r in range(reps): eng.topk(sp, (r * ib) % span, ib, K, scores, vals, idx)
            i1.record()
            sync()
            ms = i0.elapsed_time(i1) / reps
            out.append({'k': K, 'batch': ib, 'value': ib * G / (ms * 1e-3), 'unit': 'teams/s', 'us_per_call': ms * 1e3, 'fused': bool(fused)})
    del eng, sp, scores
    torch.cuda.empty_cache()
    return out


class HostBatches:
    """the host side of the end-to-end leg: global batches are packed from the host teamsvecs CSR into pinned blocks (engine.HostPacker)."""

    def __init__(self, tv, train_rows, b, rank, G):
        from opentf_b200.engine import HostPacker, to_csr
        self.b, self.rank, self.G = b, rank, G
        self.s = to_csr(tv['skill']); self.m = to_csr(tv['member'])
        self.rows = np.asarray(train_rows, dtype=np.int32)
        gB = b * G
        # capacities: the longest global batch of the run (+ slack), so that every batch has the same device layout (one captured graph)
        ls, lm = np.diff(self.s[0])[self.rows], np.diff(self.m[0])[self.rows]
        nb = len(self.rows) // gB
        lo, hi = self._slice()  # the skill CSR travels for this rank's rows only, the member CSR for the whole global batch
        cap = lambda l, a, z: int(-(-int(l[:nb * gB].reshape(nb, gB)[:, a:z].sum(1).max() * 1.02 + 64) // 64) * 64)
        self.cap_s, self.cap_m = cap(ls, lo, hi), cap(lm, 0, gB)
        self.packer = HostPacker(self.s, self.m, gB, self.cap_s, self.cap_m)
        self.h2d_bytes = self.packer.words * 4

    def _rows(self, i):
        gB = self.b * self.G
        g0 = (i * gB) % (len(self.rows) - gB)
        return self.rows[g0:g0 + gB]

    def _slice(self):
        gB = self.b * self.G
        per = -(-gB // self.G)
        return min(gB, self.rank * per), min(gB, (self.rank + 1) * per)

    def step(self, eng, i):
        return eng.step_host(self.packer.pack(self._rows(i), *self._slice()), self.b * self.G, self.cap_s, self.cap_m, self.rank, self.G, lr=1e-3)

    def run(self, eng, first, steps):
        """`steps` consecutive batches, every one packed from the host CSR, copied to the device and stepped, the loss of EVERY step read back.
        Two steps are in flight: while the GPU works on batch i the host packs and enqueues batch i+1 (its copy runs on the library's copy
        stream), then waits for loss i (its own copy event, not the stream) -- the GPU never waits for the host as long as pack + enqueue take
        less than a step"""
        gB = self.b * self.G
        lo, hi = self._slice()
        eng.step_host(self.packer.pack(self._rows(first), lo, hi), gB, self.cap_s, self.cap_m, self.rank, self.G, lr=1e-3, sync=False, slot=0)
        for i in range(1, steps):
            blk = self.packer.pack(self._rows(first + i), lo, hi)  # (block i % 3: the copy of batch i-3 out of it completed long ago -- loss i-2 was read)
            eng.step_host(blk, gB, self.cap_s, self.cap_m, self.rank, self.G, lr=1e-3, sync=False, slot=i & 1)
            eng.step_host_loss((i - 1) & 1)
        return eng.step_host_loss((steps - 1) & 1)


if __name__ == '__main__':
    a = parse()
    if a.impl == 'reference': run_reference(a)
    else: run_ours(a)
