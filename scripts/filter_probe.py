"""Timing probe (GPU): the filter kernel with its epilogue stages switched off (TRK_FILTER_DEBUG), the exact top-k
kernel and the dense kernel, on one synthetic shape.  Prints ms per launch."""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from tensorrec_b200 import kernels  # noqa: E402


class A:
    users, items, d, k = int(sys.argv[1]) if len(sys.argv) > 1 else 262144, int(sys.argv[2]) if len(sys.argv) > 2 else 262144, 128, 10


uf, itf, wu, wi, bu, bi = bench.make_problem(A)
dev = torch.device('cuda', 0)
d_pad = kernels.d_pad_for(A.d)
ucsr, icsr = kernels.DeviceCSR.from_scipy(uf, device=dev), kernels.DeviceCSR.from_scipy(itf, device=dev)
u32, us, usc = kernels.gather_reduce(ucsr, torch.from_numpy(wu).to(dev), want_f32=True, split_d_pad=d_pad)
i32, its, isc = kernels.gather_reduce(icsr, torch.from_numpy(wi).to(dev), want_f32=True, split_d_pad=d_pad)
ub = kernels.project_biases(ucsr, torch.from_numpy(bu).to(dev))
ib = kernels.project_biases(icsr, torch.from_numpy(bi).to(dev))
stats = torch.zeros(3, device=dev)
unorm = kernels.operand_stats(us, usc, d_pad)
kernels.operand_stats(its, isc, d_pad, want_norm=False, stats=stats)
perm = None if os.environ.get('PROBE_NO_SORT') else kernels.bias_processing_order(ib)
hi = kernels.rescale_hi_global(its, isc, stats, d_pad, perm=perm)
bias_pad, bmax = kernels.pack_item_bias(ib, A.items, stats, dev, perm=perm)
meta = kernels.pack_item_meta(isc, ib, A.items)


def timeit(fn, n=3):
    fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / n


pairs = A.users * float(A.items)
ms = timeit(lambda: kernels.score_filter(us, usc, ub, unorm, hi, stats, bias_pad, bmax, perm, A.users, A.items, d_pad, A.k))
print('filter kernel: %.2f ms  %.3e pairs/s  %.0f TFLOP/s' % (ms, pairs / ms * 1e3, 2 * pairs * A.d / ms / 1e9))
for cl in ('1', '2'):
    os.environ['TRK_FILTER_CLUSTER'] = cl
    ms = timeit(lambda: kernels.score_filter(us, usc, ub, unorm, hi, stats, bias_pad, bmax, perm, A.users, A.items, d_pad, A.k))
    print('filter kernel, clusters of %s: %.2f ms' % (cl, ms))
os.environ.pop('TRK_FILTER_CLUSTER')
for mode in ('7', '4', '1', '2', '6'):
    os.environ['TRK_FILTER_DEBUG'] = mode
    ms = timeit(lambda: kernels.score_filter(us, usc, ub, unorm, hi, stats, bias_pad, bmax, perm, A.users, A.items, d_pad, A.k))
    print('filter debug=%s: %.2f ms' % (mode, ms))
os.environ['TRK_FILTER_DEBUG'] = '0'
ms = timeit(lambda: kernels.score_topk(us, usc, ub, its, meta, A.users, A.items, d_pad, A.k))
print('exact top-k (3 pass): %.2f ms  %.3e pairs/s  issued %.0f TFLOP/s' % (ms, pairs / ms * 1e3, 6 * pairs * A.d / ms / 1e9))
cs, ci, theta, flags = kernels.score_filter(us, usc, ub, unorm, hi, stats, bias_pad, bmax, perm, A.users, A.items, d_pad, A.k)
ms = timeit(lambda: kernels.rescore_topk(u32, i32, ub, ib, ci, theta, flags, unorm, stats, A.k))
print('rescore_topk: %.2f ms' % ms)
nu = min(A.users, 32768)
out = torch.empty((nu, A.items), dtype=torch.float32, device=dev)
ms = timeit(lambda: kernels.score_dense_tc(us[:nu], usc[:nu], ub[:nu], its, meta, nu, A.items, d_pad, out=out))
print('dense (3 pass) %d x %d: %.2f ms  %.3e pairs/s  %.0f GB/s written' % (nu, A.items, ms, nu * float(A.items) / ms * 1e3, nu * float(A.items) * 4 / ms / 1e6))
