"""K1 (csr_gather_reduce) timing at the bench shape (L2 flushed between launches): output variants, and with the
re-referenced tag rows of the indicator regime pinned in L2 (trk_l2_persist_window).
usage: python scripts/k1_probe.py [rows]"""
import ctypes
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from tensorrec_b200 import kernels, _lib  # noqa: E402


class A:
    users = int(sys.argv[1]) if len(sys.argv) > 1 else 1000000
    items, d, k = 1024, 128, 10


uf, itf, wu, wi, bu, bi = bench.make_problem(A)
dev = torch.device('cuda', 0)
ucsr = kernels.DeviceCSR.from_scipy(uf, device=dev)
w = torch.from_numpy(wu).to(dev)
bu_d = torch.from_numpy(bu).to(dev)
d_pad = kernels.d_pad_for(A.d)
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
lib = _lib.load()


def timeit(fn, n=7):
    fn()
    ts = []
    for _ in range(n):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return sorted(ts)[len(ts) // 2]


nnz = uf.nnz
distinct = int(np.unique(uf.indices).shape[0])
survey = nnz * 8 + (A.users + 1) * 4 + distinct * A.d * 4 + A.users * A.d * 4
print('%d rows, nnz %d, distinct columns %d, SURVEY 8(d) bytes %.3f GB, persisting L2 capacity %.1f MB'
      % (A.users, nnz, distinct, survey / 1e9, lib.trk_l2_persist_capacity() / 1e6))


def report(name, ms):
    print('%-58s %.3f ms  %.0f GB/s on the 8(d) bytes = %.3f of 6563.9' % (name, ms, survey / ms / 1e6,
                                                                            survey / ms / 1e6 / 6563.9))


variants = [('split only', dict(want_f32=False, split_d_pad=d_pad)),
            ('split + norm (the filter path)', dict(want_f32=False, split_d_pad=d_pad, want_norm=True)),
            ('fp32 only', dict(want_f32=True)),
            ('fp32 + split', dict(want_f32=True, split_d_pad=d_pad))]
for name, kw in variants:
    report('K1 ' + name, timeit(lambda: kernels.gather_reduce(ucsr, w, **kw)))
report('project_biases', timeit(lambda: kernels.project_biases(ucsr, bu_d)))
# the tag table of the indicator regime = weight rows [users, 1.2 users): pin it
stream = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
tag_base = w.data_ptr() + A.users * A.d * 4
tag_bytes = (w.shape[0] - A.users) * A.d * 4
for ratio in (1.0, 0.6):
    rc = lib.trk_l2_persist_window(ctypes.c_void_p(tag_base), tag_bytes, ctypes.c_float(ratio), stream)
    print('persist window over the tag rows (%.0f MB, hit ratio %.1f): rc=%d %s' % (tag_bytes / 1e6, ratio, rc,
                                                                                    _lib.last_error() if rc else ''))
    for name, kw in variants[:2]:
        report('K1 ' + name + ' [tag rows persisting]', timeit(lambda: kernels.gather_reduce(ucsr, w, **kw)))
lib.trk_l2_persist_window(None, 0, ctypes.c_float(0.0), stream)
report('K1 split only [window cleared]', timeit(lambda: kernels.gather_reduce(ucsr, w, want_f32=False, split_d_pad=d_pad)))
