"""Timing probe (GPU): the filter kernel with its epilogue stages switched off (TRK_FILTER_DEBUG), per launch form, on
one synthetic shape -- e.g. the 1/8 item shard every rank of an 8-GPU run sweeps.  Prints ms per launch.
usage: python scripts/filter_probe.py [users] [items]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from tensorrec_b200 import kernels  # noqa: E402


class A:
    users = int(sys.argv[1]) if len(sys.argv) > 1 else 262144
    items = int(sys.argv[2]) if len(sys.argv) > 2 else 262144
    d, k = 128, 10


uf, itf, wu, wi, bu, bi = bench.make_problem(A)
dev = torch.device('cuda', 0)
d_pad = kernels.d_pad_for(A.d)
ucsr, icsr = kernels.DeviceCSR.from_scipy(uf, device=dev), kernels.DeviceCSR.from_scipy(itf, device=dev)
stats = torch.empty(3, device=dev)
_, us, usc, unorm = kernels.gather_reduce(ucsr, torch.from_numpy(wu).to(dev), want_f32=False, split_d_pad=d_pad,
                                          want_norm=True)
_, its, isc = kernels.gather_reduce(icsr, torch.from_numpy(wi).to(dev), want_f32=False, split_d_pad=d_pad, stats=stats)
ub = kernels.project_biases(ucsr, torch.from_numpy(bu).to(dev))
ib = kernels.project_biases(icsr, torch.from_numpy(bi).to(dev))
users = kernels.SideOperands(None, us, usc, ub, A.users, A.d, d_pad, norm=unorm)
items = kernels.SideOperands(None, its, isc, ib, A.items, A.d, d_pad, stats=stats)
f = kernels.FilterItems(items)


def timeit(fn, n=4):
    fn()
    fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / n


def run_filter():
    return kernels.score_filter(us, usc, ub, unorm, f.hi, f.stats, f.bias_pad, f.block_max, f.perm, A.users, A.items,
                                d_pad, A.k, block_bias_min=f.block_min)


def timeit_cool(fn, n=5, pause=0.4):
    """Single launches separated by idle time: the GPU is not at its power cap, the SM clock is near its maximum -- what a
    30 ms kernel inside a short multi-GPU step sees (back-to-back launches settle at ~1.45 GHz under the 1 kW cap)."""
    import time
    fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(n):
        time.sleep(pause)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return sorted(ts)[len(ts) // 2]


pairs = A.users * float(A.items)
print('shape %d users x %d items, d=%d' % (A.users, A.items, A.d))
if os.environ.get('PROBE_COOL'):
    print('single launches after idle (high clocks): full kernel %.2f ms' % timeit_cool(run_filter))
    for mode in ('4', '1', '2', '6'):
        os.environ['TRK_FILTER_DEBUG'] = mode
        print('  cool debug=%s: %.2f ms' % (mode, timeit_cool(run_filter)))
    os.environ['TRK_FILTER_DEBUG'] = '0'
    for trig in ('32', '26'):
        os.environ['TRK_FILTER_TILE_END_TRIGGER'] = trig
        print('  cool, tile-end compaction above %s entries: %.2f ms' % (trig, timeit_cool(run_filter)))
    os.environ.pop('TRK_FILTER_TILE_END_TRIGGER')
    for cl in ('1', '2'):
        os.environ['TRK_FILTER_CLUSTER'] = cl
        print('  cool, clusters of %s: %.2f ms' % (cl, timeit_cool(run_filter)))
    os.environ.pop('TRK_FILTER_CLUSTER')
ms = timeit(run_filter)
print('filter kernel: %.2f ms  %.3e pairs/s  %.0f TFLOP/s' % (ms, pairs / ms * 1e3, 2 * pairs * A.d / ms / 1e9))
for trig in ('32', '26', '22'):
    os.environ['TRK_FILTER_TILE_END_TRIGGER'] = trig
    print('filter kernel, tile-end compaction above %s entries: %.2f ms' % (trig, timeit(run_filter)))
os.environ.pop('TRK_FILTER_TILE_END_TRIGGER')
if os.environ.get('PROBE_SHORT'):
    sys.exit(0)
os.environ['TRK_FILTER_NO_WARMSTART'] = '1'
print('filter kernel, no threshold prologue at all: %.2f ms' % timeit(run_filter))
os.environ.pop('TRK_FILTER_NO_WARMSTART')
for cl in ('1', '2'):
    os.environ['TRK_FILTER_CLUSTER'] = cl
    print('filter kernel, clusters of %s: %.2f ms' % (cl, timeit(run_filter)))
    for mode in ('4', '1', '2', '6'):
        os.environ['TRK_FILTER_DEBUG'] = mode
        print('  cluster %s debug=%s: %.2f ms' % (cl, mode, timeit(run_filter)))
    os.environ['TRK_FILTER_DEBUG'] = '0'
os.environ.pop('TRK_FILTER_CLUSTER')
os.environ['TRK_FILTER_DEBUG'] = '7'
timeit(run_filter, n=1)
os.environ['TRK_FILTER_DEBUG'] = '0'
_, ci, theta = run_filter()
print('rescore_topk: %.2f ms' % timeit(lambda: kernels.rescore_topk(users, items, ci, theta, unorm, f.stats, A.k)))
top, bad = kernels.rescore_topk(users, items, ci, theta, unorm, f.stats, A.k)
print('flagged rows: %d' % int(bad.sum()))
print('device-side fallback: %.2f ms' % timeit(lambda: kernels.rerun_uncertified(users, items, bad, top, A.k)))
meta = kernels.pack_item_meta(isc, ib, A.items)
if A.users * float(A.items) <= 3e11:
    ms = timeit(lambda: kernels.score_topk(us, usc, ub, its, meta, A.users, A.items, d_pad, A.k), n=2)
    print('exact top-k (3 pass): %.2f ms  %.3e pairs/s  issued %.0f TFLOP/s' % (ms, pairs / ms * 1e3, 6 * pairs * A.d / ms / 1e9))
