"""K1 (csr_gather_reduce) timing at the bench shape (cold L2), with and without the fp32 output.
usage: python scripts/k1_probe.py [rows]"""
import os
import sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from tensorrec_b200 import kernels  # noqa: E402


class A:
    users = int(sys.argv[1]) if len(sys.argv) > 1 else 1000000
    items, d, k = 1024, 128, 10


uf, itf, wu, wi, bu, bi = bench.make_problem(A)
dev = torch.device('cuda', 0)
ucsr = kernels.DeviceCSR.from_scipy(uf, device=dev)
w = torch.from_numpy(wu).to(dev)
d_pad = kernels.d_pad_for(A.d)
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)


def timeit(fn, n=5):
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
distinct = len(set(uf.indices.tolist())) if A.users <= 2000000 else 0
for want_f32 in (True, False):
    alg = nnz * 8 + (A.users + 1) * 4 + distinct * A.d * 4 + A.users * ((A.d * 4 if want_f32 else 0) + 2 * d_pad * 2 + 4)
    ms = timeit(lambda: kernels.gather_reduce(ucsr, w, want_f32=want_f32, split_d_pad=d_pad))
    print('K1 %d rows, f32 out %s: %.3f ms  %.0f GB/s algorithmic (%.2f GB)' % (A.users, want_f32, ms, alg / ms / 1e6,
                                                                                alg / 1e9))
