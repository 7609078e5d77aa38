#!/usr/bin/env python
"""bench.py -- predict_rank throughput of the B200-native hot path (BASELINE.json metric), one JSON line on stdout.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]

Workload (BASELINE.json north_star / SURVEY.md 8d, "C5 at 1M x 1M"): predict_rank top-10 over 1M users x 1M items,
n_components = 128, indicator-regime sparse features (identity + 3 random tags per row, F = 1.2 R, ~4 nnz/row),
LinearRepresentationGraph x DotProductPredictionGraph, biased, n_tastes = 1.  Synthetic, seeded.

One step = one full pass of the hot path over the batch:
    K1 users -> split operand,  K1 items -> split operand,  2 x project_biases,  pack item meta,
    K2+K3 fused tcgen05 score + top-k,  merge           [N > 1: item axis sharded, + 1 NCCL all-gather, merge]
value  = U * I / step time with the CSR inputs and the weights already resident in HBM (CUDA events, max over ranks);
e2e    = the same metric through TensorRec.predict_rank(user_features, item_features, k) with HOST scipy matrices
         (pinned): host->device copy of the CSR arrays and device->host read of the top-k inside the timed region;
roofline: the fused kernel's algorithmic flops (2*U*I*d) / its CUDA-event time against the measured bf16 peak;
cpu_baseline: the oracle (numpy/scipy restatement of the reference's TF-CPU ops) on this box's host cores, on a
         bounded user sample of the same workload.
`--impl reference` times that oracle alone (TensorFlow, the reference's only back-end, cannot be installed)."""
import argparse
import concurrent.futures
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np
import scipy.sparse as sp

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = 'predict_rank_pairs_per_s'
UNIT = 'pairs/s'


def log(*a):
    print(*a, file=sys.stderr, flush=True)


# ----------------------------------------------------------------------------------------------------- inputs
def indicator_csr(rows, seed):
    """tensorrec/util.py:88-108 (indicator regime), vectorised and seeded; float32 CSR with sorted rows."""
    rng = np.random.default_rng(seed)
    n_features = int(rows * 1.2)
    n_tags = rows * 3
    r = np.concatenate([np.arange(rows, dtype=np.int64), rng.integers(0, rows, n_tags)])
    c = np.concatenate([np.arange(rows, dtype=np.int64), rng.integers(rows, n_features, n_tags)])
    m = sp.csr_matrix((np.ones(r.shape[0], dtype=np.float32), (r, c)), shape=(rows, n_features))
    m.sum_duplicates()
    m.data[:] = 1.0
    return m


def make_weights(n_features, d, seed):
    """representation_graphs.py:35-36: normal rows, L2-normalised (float32)."""
    rng = np.random.default_rng(seed)
    w = rng.standard_normal((n_features, d), dtype=np.float32)
    w /= np.sqrt(np.einsum('ij,ij->i', w, w, dtype=np.float32))[:, None]
    return w


def make_problem(args):
    t0 = time.time()
    uf = indicator_csr(args.users, seed=0)
    itf = indicator_csr(args.items, seed=1)
    wu = make_weights(uf.shape[1], args.d, seed=2)
    wi = make_weights(itf.shape[1], args.d, seed=3)
    rng = np.random.default_rng(4)
    bu = (0.1 * rng.standard_normal(uf.shape[1])).astype(np.float32)
    bi = (0.1 * rng.standard_normal(itf.shape[1])).astype(np.float32)
    log('[bench] synthetic problem built in %.1fs: users %s nnz %d, items %s nnz %d, d=%d'
        % (time.time() - t0, uf.shape, uf.nnz, itf.shape, itf.nnz, args.d))
    return uf, itf, wu, wi, bu, bi


# ----------------------------------------------------------------------------------------------------- CPU oracle leg
def cpu_oracle_leg(uf, itf, wu, wi, bu, bi, k, budget_s, threads):
    """Times the oracle (reference semantics: SpMM, fp32 GEMM, bias adds, the literal double full sort per user,
    then the rank <= k entries) on a bounded sample of users against ALL items.  Returns (pairs_per_s, description).

    The item-side work (item representation + item biases) is done once per predict_rank call by the reference; it is
    timed once and charged to the sample in proportion sample_users / total_users."""
    from oracle import reference_ops as R
    n_users, n_items = uf.shape[0], itf.shape[0]
    t0 = time.perf_counter()
    item_repr = R.sparse_dense_matmul_fast(itf, wi)
    item_bias = np.asarray(itf @ bi, dtype=np.float32)
    t_items = time.perf_counter() - t0

    def rank_rows(block):
        order = np.argsort(-block, axis=1, kind='stable').astype(np.int32)           # recommendation_graphs.py:81
        ranks = np.argsort(order, axis=1, kind='stable').astype(np.int32) + 1         # :82
        rows, cols = np.nonzero(ranks <= k)                                           # eval.py:23,49 read only these
        top = np.empty((block.shape[0], min(k, block.shape[1])), dtype=np.int32)
        top[rows, ranks[rows, cols] - 1] = cols
        return top

    def run(u0, u1):
        sub = uf[u0:u1]
        user_repr = R.sparse_dense_matmul_fast(sub, wu)
        user_bias = np.asarray(sub @ bu, dtype=np.float32)
        scores = R.bias_prediction_dense(R.dot_product_dense(user_repr, item_repr), user_bias, item_bias)
        rows_per = max(1, (u1 - u0 + threads - 1) // threads)
        blocks = [scores[i:i + rows_per] for i in range(0, u1 - u0, rows_per)]
        with concurrent.futures.ThreadPoolExecutor(max_workers=threads) as pool:
            return np.concatenate(list(pool.map(rank_rows, blocks)))

    chunk = max(threads, 8)
    chunk = min(chunk, n_users)
    t0 = time.perf_counter()
    first_top = run(0, chunk)       # kept: the GPU result of the same users is checked against it (full item axis)
    t_chunk = time.perf_counter() - t0
    n_chunks = int(max(1, min(budget_s / max(t_chunk, 1e-3), n_users // chunk)))
    t0 = time.perf_counter()
    done = 0
    for c in range(n_chunks):
        run(c * chunk, (c + 1) * chunk)
        done += chunk
        if time.perf_counter() - t0 > budget_s:
            break
    t_users = time.perf_counter() - t0
    total = t_users + t_items * done / float(n_users)
    desc = ('%d of %d users x all %d items, d=%d: scipy CSR SpMM + numpy fp32 GEMM + per-user double stable argsort '
            '(rank_predictions) + rank<=%d selection, %d threads; item-side time charged pro rata'
            % (done, n_users, n_items, wu.shape[1], k, threads))
    return done * float(n_items) / total, desc, total, first_top


# ----------------------------------------------------------------------------------------------------- clocks
class ClockSampler(object):
    QUERY = ('index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,'
             'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,'
             'clocks_event_reasons.sw_power_cap')

    def __init__(self, gpu_index):
        self.gpu_index = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(['nvidia-smi', '--query-gpu=' + self.QUERY, '--format=csv,noheader,nounits',
                                          '-lms', '100', '-i', str(self.gpu_index)], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, smax, power, reasons = [], [], [], set()
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        for line in self.lines:
            f = [x.strip() for x in line.split(',')]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                smax.append(float(f[2]))
                power.append(float(f[3]))
            except ValueError:
                continue
            for name, flag in zip(names, f[5:9]):
                if flag.lower().startswith('active'):
                    reasons.add(name)
        if not sm:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['no samples']}
        busy = [s for s, p in zip(sm, power) if p >= 0.5 * max(power)] or sm
        return {'sm_mhz': float(np.median(busy)), 'sm_max_mhz': float(max(smax)), 'reasons': sorted(reasons),
                'power_w_max': float(max(power)), 'samples': len(sm)}


def ncu_traffic(kernel_key):
    """DRAM bytes (read + write) of one launch of the dominant kernel at the bench workload, from the committed ncu
    capture (profiles/ncu_traffic.json, written by scripts/ncu_summary.py runs); None if not captured."""
    path = os.path.join(ROOT, 'profiles', 'ncu_traffic.json')
    if os.path.exists(path):
        return json.load(open(path)).get(kernel_key)
    return None


def measured_peaks():
    path = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(path):
        p = json.load(open(path))
        return {'hbm_gbs': p['hbm_gbs'], 'tflops_burst': p['bf16_tflops'],
                'tflops_sustained': p.get('bf16_tflops_sustained', p['bf16_tflops']), 'source': 'measured'}
    return {'hbm_gbs': 6650.0, 'tflops_burst': 1590.0, 'tflops_sustained': 1400.0, 'source': 'fallback'}


# ----------------------------------------------------------------------------------------------------- GPU arm
def run_b200(args):
    import torch
    import torch.distributed as dist
    import tensorrec_b200
    from tensorrec_b200 import kernels
    from tensorrec_b200.distributed import shard_bounds, all_gather_candidates

    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    if world != args.gpus:
        log('[bench] note: --gpus %d but WORLD_SIZE=%d; using WORLD_SIZE' % (args.gpus, world))
    kernels.require_cuda()
    torch.cuda.set_device(local_rank)
    dev = torch.device('cuda', local_rank)
    if world > 1:
        os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
        dist.init_process_group('nccl', device_id=dev)

    uf, itf, wu, wi, bu, bi = make_problem(args)
    n_users, n_items, d, k = args.users, args.items, args.d, args.k
    d_pad = kernels.d_pad_for(d)
    lo, hi = shard_bounds(n_items, world, rank)       # item axis sharded over ranks
    itf_local = itf[lo:hi]
    n_local = hi - lo

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- resident inputs for `value` ------------------------------------------------------------------
    ucsr = kernels.DeviceCSR.from_scipy(uf, device=dev)
    icsr = kernels.DeviceCSR.from_scipy(itf_local, device=dev)
    wu_d, wi_d = torch.from_numpy(wu).to(dev), torch.from_numpy(wi).to(dev)
    bu_d, bi_d = torch.from_numpy(bu).to(dev), torch.from_numpy(bi).to(dev)
    ev = {'k1u': [], 'fused': []}

    use_filter = args.topk_path == 'filter' and k <= kernels.filter_max_k()
    info = {}

    def step(record=False):
        e = [torch.cuda.Event(enable_timing=True) for _ in range(4)] if record else None
        if record:
            e[0].record()
        u32, us, usc = kernels.gather_reduce(ucsr, wu_d, want_f32=use_filter, split_d_pad=d_pad)
        if record:
            e[1].record()
        i32, its, isc = kernels.gather_reduce(icsr, wi_d, want_f32=use_filter, split_d_pad=d_pad)
        ub = kernels.project_biases(ucsr, bu_d)
        ib = kernels.project_biases(icsr, bi_d)
        users = kernels.SideOperands(u32, us, usc, ub, n_users, d, d_pad)
        items = kernels.SideOperands(i32, its, isc, ib, n_local, d, d_pad)
        if use_filter:
            user_norm = kernels.operand_stats(us, usc, d_pad)
            fitems = kernels.FilterItems(items)      # stats, bias-sorted processing order, global-scale hi, bias blocks
            if record:
                e[2].record()
            cs, ci, theta, flags = kernels.score_filter(us, usc, ub, user_norm, fitems.hi, fitems.stats,
                                                        fitems.bias_pad, fitems.block_max, fitems.perm, n_users,
                                                        n_local, d_pad, k, item_id_offset=lo)
            if record:
                e[3].record()
            ts, ti, bad = kernels.rescore_topk(u32, i32, ub, ib, ci, theta, flags, user_norm, fitems.stats, k,
                                               item_id_offset=lo)
            # rows whose bound could not be certified go through the exact kernel (inside the timed step)
            info['fallback_rows'] = kernels.rerun_uncertified(users, items, bad, ts, ti, k, item_id_offset=lo)
        else:
            meta = kernels.pack_item_meta(isc, ib, n_local)
            if record:
                e[2].record()
            cs, ci = kernels.score_topk(us, usc, ub, its, meta, n_users, n_local, d_pad, k, item_id_offset=lo)
            if record:
                e[3].record()
            ts, ti = kernels.topk_merge(cs, ci, k)
        if world > 1:
            gs, gi = all_gather_candidates(ts, ti)
            ts, ti = kernels.topk_merge(gs, gi, k)
        if record:
            ev['k1u'].append((e[0], e[1]))
            ev['fused'].append((e[2], e[3]))
        return ts, ti

    n_splits = kernels.default_splits(n_users, n_local)
    launches_per_step = (10 if use_filter else 7) + (1 if world > 1 else 0)

    for _ in range(args.warmup):
        step()
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    start, end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    launches_before = tensorrec_b200._lib.launch_count
    start.record()
    for _ in range(args.steps):
        out = step(record=True)
    end.record()
    barrier()
    gpu_launches = tensorrec_b200._lib.launch_count - launches_before   # kernel-launching C-ABI calls, counted
    ms_total = start.elapsed_time(end)
    clocks = sampler.stop() if rank == 0 else None
    t = torch.tensor([ms_total], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_step = float(t.item()) / args.steps
    value = n_users * float(n_items) / (ms_step * 1e-3)
    fused_ms = float(np.mean([a.elapsed_time(b) for a, b in ev['fused']]))
    k1u_ms = float(np.mean([a.elapsed_time(b) for a, b in ev['k1u']]))
    top_items_value = out[1][:4].cpu().numpy()
    top_items_check = out[1][:1024].cpu().numpy()      # compared with the CPU oracle's ranking below (rank 0)

    # ---- e2e: the public API with host buffers ----------------------------------------------------------
    def pinned_csr(m):
        arrs = [torch.from_numpy(np.ascontiguousarray(a)).pin_memory() for a in (m.data, m.indices, m.indptr)]
        return sp.csr_matrix((arrs[0].numpy(), arrs[1].numpy(), arrs[2].numpy()), shape=m.shape), arrs

    del ucsr, icsr
    tensorrec_b200.tensorrec.TOPK_PATH = 'auto' if use_filter else 'exact'
    model = tensorrec_b200.TensorRec(n_components=d)
    model.set_weights({'linear_weights_user_0': wu, 'linear_weights_item': wi, 'feature_biases_user': bu[:, None],
                       'feature_biases_item': bi[:, None]})
    uf_host, _keep_u = pinned_csr(uf)
    itf_host, _keep_i = pinned_csr(itf_local)
    group = dist.group.WORLD if world > 1 else None

    def e2e_step():
        return model.predict_top_k(uf_host, itf_host, k, item_id_offset=lo, gather_group=group)

    for _ in range(max(1, min(args.warmup, 2))):
        e2e_step()
    barrier()
    t0 = time.perf_counter()
    s2, e2 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s2.record()
    for _ in range(args.steps):
        top = e2e_step()
    e2.record()
    barrier()
    wall = (time.perf_counter() - t0) * 1e3
    # host conversion and the blocking D2H sit between kernels: use the larger of the event and wall times
    e2e_ms = max(s2.elapsed_time(e2), wall)
    t = torch.tensor([e2e_ms], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_ms_step = float(t.item()) / args.steps
    e2e_value = n_users * float(n_items) / (e2e_ms_step * 1e-3)
    h2d = 4 * (uf.nnz * 2 + uf.shape[0] + 1 + itf_local.nnz * 2 + itf_local.shape[0] + 1)
    d2h = n_users * k * 8
    same = bool(np.array_equal(top.items[:4], top_items_value))

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    peaks = measured_peaks()
    flops = 2.0 * n_users * n_local * d                        # algorithmic flops of one fused launch (this rank)
    achieved = flops / (fused_ms * 1e-3) / 1e12
    peak = peaks['tflops_sustained']
    # K1 (users) algorithmic bytes: nnz*8 + (R+1)*4 + D*d*4 + R*(2*d_pad*2 + 4) [+ R*d*4]  (D = distinct columns)
    distinct = int(np.unique(uf.indices).shape[0])
    k1_bytes = (uf.nnz * 8 + (n_users + 1) * 4 + distinct * d * 4 + n_users * (2 * d_pad * 2 + 4)
                + (n_users * d * 4 if use_filter else 0))      # the filter path also writes the fp32 representation
    k1_gbs = k1_bytes / (k1u_ms * 1e-3) / 1e9

    cores = os.cpu_count() or 1
    cpu_value, cpu_desc, cpu_s, cpu_top = cpu_oracle_leg(uf, itf, wu, wi, bu, bi, k, args.cpu_budget, cores)
    # parity at the full item count: the reference-semantics ranking of the first users (oracle, CPU) against the GPU
    # top-k of the same users; only sub-tolerance near-ties may order differently (fp32 rounding of the two GEMMs)
    n_chk = min(cpu_top.shape[0], top_items_check.shape[0])
    agree = float((cpu_top[:n_chk] == top_items_check[:n_chk]).mean()) if n_chk else None
    same_sets = float(np.mean([set(cpu_top[i]) == set(top_items_check[i]) for i in range(n_chk)])) if n_chk else None

    result = {
        'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': world, 'steps': args.steps, 'warmup': args.warmup,
        'ms_per_step': ms_step, 'higher_is_better': True, 'scaling': 'strong', 'vs_baseline': None,
        'dtype': ('f32 (1 fp16 tcgen05 filter pass with a certified bound + exact fp32 re-scoring of the survivors)'
                  if use_filter else 'f32 (3 x fp16 split-product tcgen05 passes, fp32 accumulate)'),
        'data': 'synthetic',
        'config': {'workload': 'predict_rank top-%d, %d users x %d items, d=%d, indicator-regime features, '
                               'LinearRepr x DotProduct, biased (BASELINE configs[4] shape at 1M x 1M; SURVEY C5)'
                               % (k, n_users, n_items, d),
                   'parallelism': 'item-sharded x%d + 1 NCCL all-gather' % world if world > 1 else 'single GPU',
                   'n_splits': n_splits, 'topk_path': 'filter+rescore' if use_filter else 'exact3',
                   'fallback_rows_last_step': info.get('fallback_rows', 0), 'l2': 'inputs exceed L2 (operands %.0f MB, tables %.0f MB)'
                   % ((n_users + n_local) * 2 * d_pad * 2 / 1e6, (wu.nbytes + wi.nbytes) / 1e6)},
        'clocks': clocks,
        'e2e': {'value': e2e_value, 'unit': UNIT, 'h2d_bytes_per_step': int(h2d), 'd2h_bytes_per_step': int(d2h),
                'ms_per_step': e2e_ms_step, 'api': 'TensorRec.predict_rank(user_features, item_features, k) on pinned '
                'host CSR', 'matches_value_arm': same},
        'gpu_launches': int(gpu_launches),
        'roofline': {'kernel': ('score_filter_kernel (trk_score_filter_f16)' if use_filter
                                else 'score_tc_kernel<topk> (trk_score_topk_f16x3)'), 'bound': 'tensor',
                     'achieved': achieved, 'peak': peak, 'unit': 'TFLOP/s', 'frac': achieved / peak,
                     'traffic': ncu_traffic('score_filter_kernel@%dx%dx%d' % (n_users, n_local, d)) if use_filter
                     else ncu_traffic('score_tc_kernel@%dx%dx%d' % (n_users, n_local, d)),
                     'peak_source': peaks['source'] + ' bf16_tflops_sustained', 'ms_per_launch': fused_ms,
                     'issued_tflops': (1 if use_filter else 3) * achieved,
                     'issued_frac': (1 if use_filter else 3) * achieved / peak, 'share_of_step': fused_ms / ms_step},
        'roofline_k1': {'kernel': 'csr_gather_reduce_kernel (users)', 'bound': 'hbm', 'achieved': k1_gbs,
                        'peak': peaks['hbm_gbs'], 'unit': 'GB/s', 'frac': k1_gbs / peaks['hbm_gbs'],
                        'ms_per_launch': k1u_ms, 'algorithmic_bytes': int(k1_bytes)},
        'cpu_baseline': {'value': cpu_value, 'unit': UNIT, 'cores': cores, 'kind': 'port', 'sample': cpu_desc,
                         'seconds': cpu_s},
        'parity': {'users_checked': int(n_chk), 'items': int(n_items), 'rank_positions_equal': agree,
                   'topk_sets_equal': same_sets,
                   'against': 'oracle (numpy restatement of the reference: fp32 GEMM + double stable argsort)'},
    }
    print(json.dumps(result), flush=True)
    if world > 1:
        dist.destroy_process_group()


# ----------------------------------------------------------------------------------------------------- reference arm
def run_reference(args):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    uf, itf, wu, wi, bu, bi = make_problem(args)
    cores = os.cpu_count() or 1
    per_step_budget = max(2.0, min(args.cpu_budget, 90.0 / max(1, args.steps + args.warmup)))
    for _ in range(args.warmup):
        cpu_oracle_leg(uf, itf, wu, wi, bu, bi, args.k, per_step_budget, cores)
    values, secs, desc = [], 0.0, ''
    for _ in range(args.steps):
        v, desc, s, _ = cpu_oracle_leg(uf, itf, wu, wi, bu, bi, args.k, per_step_budget, cores)
        values.append(v)
        secs += s
    value = float(np.mean(values))
    print(json.dumps({
        'impl': 'reference', 'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': args.gpus, 'steps': args.steps,
        'warmup': args.warmup, 'ms_per_step': secs / args.steps * 1e3, 'higher_is_better': True, 'scaling': 'strong',
        'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
        'config': {'workload': 'predict_rank top-%d, %d users x %d items, d=%d, indicator-regime features, LinearRepr x '
                               'DotProduct, biased' % (args.k, args.users, args.items, args.d),
                   'note': 'numpy/scipy restatement of the reference TF-CPU semantics (TensorFlow is not installable '
                           'here); each step is a bounded user sample of the workload'},
        'cpu_baseline': {'value': value, 'unit': UNIT, 'cores': cores, 'kind': 'port', 'sample': desc},
        'e2e': {'value': value, 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'gpu_launches': 0,
    }), flush=True)


def run_dense(args):
    """Secondary measurement (BASELINE configs[1] shape, scaled to fit HBM): predict() = K1 x2 + biases + the tensor-core
    score kernel writing the dense fp32 matrix.  Bound: HBM write, U*I*4 bytes."""
    import torch
    from tensorrec_b200 import kernels
    kernels.require_cuda()
    torch.cuda.set_device(0)
    dev = torch.device('cuda', 0)
    uf, itf, wu, wi, bu, bi = make_problem(args)
    d_pad = kernels.d_pad_for(args.d)
    ucsr, icsr = kernels.DeviceCSR.from_scipy(uf, device=dev), kernels.DeviceCSR.from_scipy(itf, device=dev)
    wu_d, wi_d = torch.from_numpy(wu).to(dev), torch.from_numpy(wi).to(dev)
    bu_d, bi_d = torch.from_numpy(bu).to(dev), torch.from_numpy(bi).to(dev)
    out = torch.empty((args.users, args.items), dtype=torch.float32, device=dev)
    ev = []

    def step():
        _, us, usc = kernels.gather_reduce(ucsr, wu_d, want_f32=False, split_d_pad=d_pad)
        _, its, isc = kernels.gather_reduce(icsr, wi_d, want_f32=False, split_d_pad=d_pad)
        ub, ib = kernels.project_biases(ucsr, bu_d), kernels.project_biases(icsr, bi_d)
        meta = kernels.pack_item_meta(isc, ib, args.items)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        kernels.score_dense_tc(us, usc, ub, its, meta, args.users, args.items, d_pad, out=out)
        b.record()
        ev.append((a, b))

    for _ in range(args.warmup):
        step()
    torch.cuda.synchronize()
    del ev[:]
    s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s0.record()
    for _ in range(args.steps):
        step()
    s1.record()
    torch.cuda.synchronize()
    ms = s0.elapsed_time(s1) / args.steps
    kms = float(np.mean([a.elapsed_time(b) for a, b in ev]))
    peaks = measured_peaks()
    gbs = args.users * float(args.items) * 4 / (kms * 1e-3) / 1e9
    print(json.dumps({'metric': 'predict_pairs_per_s', 'value': args.users * float(args.items) / (ms * 1e-3),
                      'unit': UNIT, 'n_gpus': 1, 'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': ms,
                      'config': {'workload': 'predict() dense fp32 scores, %d users x %d items, d=%d (BASELINE configs[1] '
                                             'shape, user axis cut to fit HBM)' % (args.users, args.items, args.d)},
                      'roofline': {'kernel': 'score_tc_kernel<dense>', 'bound': 'hbm', 'achieved': gbs,
                                   'peak': peaks['hbm_gbs'], 'unit': 'GB/s', 'frac': gbs / peaks['hbm_gbs'],
                                   'ms_per_launch': kms}}), flush=True)


def run_full_ranks(args):
    """Secondary measurement: predict_rank() in the reference's full mode -- dense scores, then the exact int32 rank of
    every (user, item) pair (rank_predictions, tensorrec/recommendation_graphs.py:73-82) -- at a shape whose [U, I]
    matrices fit HBM.  Bound: HBM (the score matrix is written once, read by the chunk sort, the sorted keys are
    written and re-read by log2(#chunks) merge passes, the ranks are written once)."""
    import torch
    from tensorrec_b200 import kernels
    kernels.require_cuda()
    torch.cuda.set_device(0)
    dev = torch.device('cuda', 0)
    uf, itf, wu, wi, bu, bi = make_problem(args)
    d_pad = kernels.d_pad_for(args.d)
    ucsr, icsr = kernels.DeviceCSR.from_scipy(uf, device=dev), kernels.DeviceCSR.from_scipy(itf, device=dev)
    wu_d, wi_d = torch.from_numpy(wu).to(dev), torch.from_numpy(wi).to(dev)
    bu_d, bi_d = torch.from_numpy(bu).to(dev), torch.from_numpy(bi).to(dev)
    out = torch.empty((args.users, args.items), dtype=torch.float32, device=dev)
    ev = []

    def step():
        _, us, usc = kernels.gather_reduce(ucsr, wu_d, want_f32=False, split_d_pad=d_pad)
        _, its, isc = kernels.gather_reduce(icsr, wi_d, want_f32=False, split_d_pad=d_pad)
        ub, ib = kernels.project_biases(ucsr, bu_d), kernels.project_biases(icsr, bi_d)
        meta = kernels.pack_item_meta(isc, ib, args.items)
        kernels.score_dense_tc(us, usc, ub, its, meta, args.users, args.items, d_pad, out=out)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        ranks = kernels.rank_full(out)
        b.record()
        ev.append((a, b))
        return ranks

    for _ in range(args.warmup):
        step()
    torch.cuda.synchronize()
    del ev[:]
    s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s0.record()
    for _ in range(args.steps):
        ranks = step()
    s1.record()
    torch.cuda.synchronize()
    ms = s0.elapsed_time(s1) / args.steps
    kms = float(np.mean([a.elapsed_time(b) for a, b in ev]))
    # spot check against the closed form on a few rows: rank = 1 + #greater + #equal with a lower index
    rows = np.linspace(0, args.users - 1, 4).astype(np.int64)
    sc, rk = out[rows].cpu().numpy(), ranks[rows].cpu().numpy()
    for r in range(len(rows)):
        order = np.lexsort((np.arange(args.items), -sc[r].astype(np.float64)))
        expect = np.empty(args.items, dtype=np.int64)
        expect[order] = np.arange(1, args.items + 1)
        assert np.array_equal(expect, rk[r]), 'rank_full disagrees with the closed form on row %d' % rows[r]
    peaks = measured_peaks()
    pairs = args.users * float(args.items)
    alg_bytes = pairs * (4 + 4)        # scores read once, ranks written once
    gbs = alg_bytes / (kms * 1e-3) / 1e9
    print(json.dumps({'metric': 'predict_rank_full_ranks_per_s', 'value': pairs / (ms * 1e-3), 'unit': 'ranks/s',
                      'n_gpus': 1, 'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': ms,
                      'config': {'workload': 'predict_rank() full int32 ranks, %d users x %d items, d=%d (reference '
                                             'semantics: every pair ranked)' % (args.users, args.items, args.d)},
                      'roofline': {'kernel': 'rank_chunk_sort_kernel + rank_merge_pass_kernel (trk_rank_full)',
                                   'bound': 'hbm', 'achieved': gbs, 'peak': peaks['hbm_gbs'], 'unit': 'GB/s',
                                   'frac': gbs / peaks['hbm_gbs'], 'ms_per_launch': kms,
                                   'ranks_per_s_kernel': pairs / (kms * 1e-3)}}), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=5)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    ap.add_argument('--users', type=int, default=1000000)
    ap.add_argument('--items', type=int, default=1000000)
    ap.add_argument('--d', type=int, default=128)
    ap.add_argument('--k', type=int, default=10)
    ap.add_argument('--workload', default='topk', choices=['topk', 'dense', 'ranks'])
    ap.add_argument('--topk-path', default='filter', choices=['filter', 'exact'])
    ap.add_argument('--cpu-budget', type=float, default=15.0, help='seconds of CPU work for the cpu_baseline sample')
    args = ap.parse_args()
    if args.workload == 'dense':
        run_dense(args)
    elif args.workload == 'ranks':
        run_full_ranks(args)
    elif args.impl == 'reference':
        run_reference(args)
    else:
        run_b200(args)


if __name__ == '__main__':
    main()
