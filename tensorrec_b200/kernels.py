"""Host-side launch layer: torch supplies device memory and streams, every computation is a call through the C ABI
(libtensorrec_b200.so).  No function here has a CPU path; all of them raise without a CUDA device."""
import ctypes

import numpy as np
import scipy.sparse as sp
import torch

from . import _lib

import os as _os
BIAS_ORDER = _os.environ.get('TENSORREC_B200_BIAS_ORDER', 'kernel')   # 'kernel' (trk_rank_full + trk_order_from_ranks) | 'torch'

TILE_ITEMS = 256     # item tile of the tensor-core kernel (item_meta is padded to a multiple of this)
TILE_USERS = 128


def require_cuda():
    if not torch.cuda.is_available():
        raise RuntimeError('tensorrec_b200 runs its predict / predict_rank path on a CUDA device (B200, sm_100a) only; '
                           'no CUDA device is visible and there is no CPU fallback')
    return _lib.load()


def _p(t):
    return None if t is None else ctypes.c_void_p(t.data_ptr())


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def _f32(t, device):
    if isinstance(t, np.ndarray):
        t = torch.from_numpy(np.ascontiguousarray(t, dtype=np.float32))
    return t.to(device=device, dtype=torch.float32).contiguous()


def d_pad_for(n_components):
    """Width of the split-fp16 operand: n_components rounded up to the 64-element swizzle row."""
    return ((int(n_components) + 63) // 64) * 64


class DeviceCSR(object):
    """Sparse features resident in HBM as CSR: int32 indptr[rows+1], int32 col[nnz], float32 val[nnz].

    Replaces the 5-tuple (row i64, col i64, val f32, d0, d1) of tensorrec/input_utils.py:15-40 and the
    tf.SparseTensor built from it (tensorrec/tensorrec.py:285-293).  Entry order inside a row is the order the
    reference's COO conversion yields (stable row sort), duplicates are kept, values are cast to float32."""

    def __init__(self, indptr, col, val, shape):
        self.indptr, self.col, self.val = indptr, col, val
        self.shape = (int(shape[0]), int(shape[1]))

    @property
    def nnz(self):
        return int(self.col.numel())

    @staticmethod
    def host_arrays(matrix):
        """scipy sparse matrix -> (indptr i32, col i32, val f32) numpy arrays, reference entry order."""
        if not sp.issparse(matrix):
            raise ValueError('Input must be a scipy sparse matrix')
        if matrix.shape[0] >= 2 ** 31 - 1 or matrix.shape[1] >= 2 ** 31 - 1 or matrix.nnz >= 2 ** 31 - 1:
            raise ValueError('feature matrix exceeds int32 indexing')
        if isinstance(matrix, sp.csr_matrix):
            # sp.coo_matrix(csr) walks the rows in storage order: the CSR arrays already are that order
            return (np.ascontiguousarray(matrix.indptr, dtype=np.int32),
                    np.ascontiguousarray(matrix.indices, dtype=np.int32),
                    np.ascontiguousarray(matrix.data, dtype=np.float32))
        coo = matrix if isinstance(matrix, sp.coo_matrix) else sp.coo_matrix(matrix)     # input_utils.py:29-30
        order = np.argsort(coo.row, kind='stable')
        counts = np.bincount(coo.row, minlength=coo.shape[0])
        indptr = np.zeros(coo.shape[0] + 1, dtype=np.int64)
        np.cumsum(counts, out=indptr[1:])
        return (indptr.astype(np.int32), np.ascontiguousarray(coo.col[order], dtype=np.int32),
                np.ascontiguousarray(coo.data[order], dtype=np.float32))

    @staticmethod
    def host_arrays_transposed(matrix):
        """CSR arrays of the TRANSPOSE (one row per feature column), entries of a column in ascending row order and,
        within a row, in reference entry order; duplicates are kept.  K1 on these arrays is the backward of K1:
        dW = A^T . dRepr (tf.sparse_tensor_dense_matmul's gradient w.r.t. the dense operand)."""
        indptr, col, val = DeviceCSR.host_arrays(matrix)
        n_rows, n_cols = matrix.shape
        rows = np.repeat(np.arange(n_rows, dtype=np.int32), np.diff(indptr))
        order = np.argsort(col, kind='stable')
        counts = np.bincount(col, minlength=n_cols)
        indptr_t = np.zeros(n_cols + 1, dtype=np.int64)
        np.cumsum(counts, out=indptr_t[1:])
        return indptr_t.astype(np.int32), np.ascontiguousarray(rows[order]), np.ascontiguousarray(val[order])

    @classmethod
    def from_scipy_transposed(cls, matrix, device='cuda'):
        require_cuda()
        indptr, col, val = cls.host_arrays_transposed(matrix)
        up = lambda a: torch.from_numpy(a).to(device, non_blocking=True)   # noqa: E731
        return cls(up(indptr), up(col), up(val), (matrix.shape[1], matrix.shape[0]))

    @classmethod
    def from_scipy(cls, matrix, device='cuda', pin=False):
        require_cuda()
        indptr, col, val = cls.host_arrays(matrix)

        def up(a):
            t = torch.from_numpy(a)
            if pin:
                t = t.pin_memory()
            return t.to(device, non_blocking=True)

        return cls(up(indptr), up(col), up(val), matrix.shape)

    def h2d_bytes(self):
        return 4 * (self.indptr.numel() + self.col.numel() + self.val.numel())


K1_MAX_WIDTH = 512      # widest row one K1 launch covers (256 when the width is not a multiple of 4)


def gather_reduce(csr, weights, n_normalize=0, want_f32=True, split_d_pad=None, want_norm=False, stats=None):
    """K1.  Returns (repr_f32 or None, split or None, scale or None[, norm]).

    want_norm: also return the row norms (upper bounds) formed in K1's epilogue; stats: float32[3] that receives the
    max norm / max row scale (zeroed by the call).  Both feed the filter form of the fused top-k."""
    lib = require_cuda()
    rows, n_features = csr.shape
    d = int(weights.shape[1])
    if int(weights.shape[0]) != n_features:
        raise ValueError('feature matrix has %d columns but the weights have %d rows' % (n_features, weights.shape[0]))
    dev = weights.device
    if split_d_pad is None and not want_norm and stats is None and (d > K1_MAX_WIDTH or (d % 4 != 0 and d > 256)):
        return _gather_reduce_wide(csr, weights, n_normalize), None, None
    out = torch.empty((rows, d), dtype=torch.float32, device=dev) if want_f32 else None
    split = scale = norm = None
    d_pad = 0
    if split_d_pad is not None:
        d_pad = int(split_d_pad)
        split = torch.empty((rows, 2 * d_pad), dtype=torch.float16, device=dev)
        scale = torch.empty((rows,), dtype=torch.float32, device=dev)
    if want_norm:
        norm = torch.empty((rows,), dtype=torch.float32, device=dev)
    rc = lib.trk_csr_gather_reduce_f32(_p(csr.indptr), _p(csr.col), _p(csr.val), _p(weights), rows, n_features, d,
                                       int(n_normalize), _p(out), _p(split), d_pad, _p(scale), _p(norm), _p(stats),
                                       _stream())
    _lib.check(rc, 'trk_csr_gather_reduce_f32')
    if want_norm:
        return out, split, scale, norm
    return out, split, scale


def _gather_reduce_wide(csr, weights, n_normalize):
    """Rows wider than one K1 launch covers (n_components > 512, e.g. the 4 x n_components hidden layer of a
    ReLURepresentationGraph): the component axis is cut into column blocks of at most 512, one K1 launch each (every
    output element is still the same fp32 FMA chain in CSR order), normalisation afterwards over the whole row."""
    rows = csr.shape[0]
    d = int(weights.shape[1])
    out = torch.empty((rows, d), dtype=torch.float32, device=weights.device)
    step = K1_MAX_WIDTH
    for c0 in range(0, d, step):
        c1 = min(d, c0 + step)
        block, _, _ = gather_reduce(csr, weights[:, c0:c1].contiguous(), want_f32=True)
        out[:, c0:c1] = block
    for _ in range(int(n_normalize)):
        _l2_normalize_rows_any_width_(out)
    return out


def _l2_normalize_rows_any_width_(x):
    if x.shape[1] <= 1024:
        return l2_normalize_rows_(x)
    # tf.nn.l2_normalize: x * rsqrt(max(sum x^2, 1e-12)); rows this wide are outside every kernel's row shape
    x.mul_(torch.rsqrt(torch.clamp((x * x).sum(dim=1, keepdim=True), min=1e-12)))
    return x


def split_f32(repr_f32, n_normalize=0, d_pad=None):
    lib = require_cuda()
    rows, d = repr_f32.shape
    d_pad = d_pad_for(d) if d_pad is None else int(d_pad)
    split = torch.empty((rows, 2 * d_pad), dtype=torch.float16, device=repr_f32.device)
    scale = torch.empty((rows,), dtype=torch.float32, device=repr_f32.device)
    rc = lib.trk_split_f32_to_f16x2(_p(repr_f32), rows, d, int(n_normalize), _p(split), d_pad, _p(scale), _stream())
    _lib.check(rc, 'trk_split_f32_to_f16x2')
    return split, scale


def l2_normalize_rows_(x):
    lib = require_cuda()
    rc = lib.trk_l2_normalize_rows_f32(_p(x), x.shape[0], x.shape[1], _stream())
    _lib.check(rc, 'trk_l2_normalize_rows_f32')
    return x


def project_biases(csr, feature_biases):
    lib = require_cuda()
    if int(feature_biases.numel()) != csr.shape[1]:
        raise ValueError('feature matrix has %d columns but there are %d feature biases'
                         % (csr.shape[1], feature_biases.numel()))
    out = torch.empty((csr.shape[0],), dtype=torch.float32, device=feature_biases.device)
    rc = lib.trk_csr_project_biases_f32(_p(csr.indptr), _p(csr.col), _p(csr.val), _p(feature_biases), csr.shape[0],
                                        _p(out), _stream())
    _lib.check(rc, 'trk_csr_project_biases_f32')
    return out


def score_exact(user_repr, item_repr, user_bias=None, item_bias=None, mode=0, attention_repr=None, out=None):
    """K2 on CUDA cores (exact fp32).  user_repr [T, U, d] or [U, d]; returns [U, I] float32."""
    lib = require_cuda()
    if user_repr.dim() == 2:
        user_repr = user_repr.unsqueeze(0)
    user_repr = user_repr.contiguous()
    n_tastes, n_users, d = user_repr.shape
    n_items = item_repr.shape[0]
    if item_repr.shape[1] != d:
        raise ValueError('user and item representations differ in n_components (%d vs %d)' % (d, item_repr.shape[1]))
    if out is None:
        out = torch.empty((n_users, n_items), dtype=torch.float32, device=user_repr.device)
    # grid.y carries 64-row user tiles (<= 65535 per launch): block the user axis for very tall inputs
    max_rows = 65535 * 64
    for u0 in range(0, max(n_users, 1), max_rows):
        u1 = min(n_users, u0 + max_rows)
        ur = user_repr[:, u0:u1].contiguous() if (u0 > 0 or u1 < n_users) else user_repr
        ub = None if user_bias is None else user_bias[u0:u1]
        if attention_repr is not None:
            ar = attention_repr[:, u0:u1].contiguous()
            rc = lib.trk_score_attention_f32(_p(ur), _p(ar), _p(item_repr), _p(ub), _p(item_bias), _p(out[u0:u1]),
                                             u1 - u0, n_items, d, n_tastes, _stream())
            _lib.check(rc, 'trk_score_attention_f32')
        else:
            rc = lib.trk_score_f32(_p(ur), _p(item_repr), _p(ub), _p(item_bias), _p(out[u0:u1]), u1 - u0, n_items, d,
                                   n_tastes, int(mode), _stream())
            _lib.check(rc, 'trk_score_f32')
    return out


def rank_full(scores):
    """K3 (full): the reference's rank_predictions on a dense [U, I] float32 matrix -> int32 ranks."""
    lib = require_cuda()
    scores = scores.contiguous()
    n_users, n_items = scores.shape
    ranks = torch.empty((n_users, n_items), dtype=torch.int32, device=scores.device)
    need = int(lib.trk_rank_full_workspace_bytes(n_users, n_items))
    ws = torch.empty((max(need, 8) // 8,), dtype=torch.int64, device=scores.device) if need else None
    rc = lib.trk_rank_full(_p(scores), _p(ranks), n_users, n_items, _p(ws), need, _stream())
    _lib.check(rc, 'trk_rank_full')
    return ranks


def padded_items(n_items):
    return ((int(n_items) + TILE_ITEMS - 1) // TILE_ITEMS) * TILE_ITEMS


def pack_item_meta(item_scale, item_bias, n_items):
    lib = require_cuda()
    n_pad = padded_items(n_items)
    meta = torch.empty((n_pad, 2), dtype=torch.float32, device=item_scale.device)
    rc = lib.trk_pack_item_meta(_p(item_scale), _p(item_bias), n_items, _p(meta), n_pad, _stream())
    _lib.check(rc, 'trk_pack_item_meta')
    return meta


def topk_max_k(d_pad):
    return int(require_cuda().trk_score_topk_max_k(int(d_pad)))


def default_splits(n_users, n_items):
    """Item-range splits so that (user blocks x splits) covers every SM about twice when there are few users."""
    n_sm = torch.cuda.get_device_properties(torch.cuda.current_device()).multi_processor_count
    n_ub = (n_users + TILE_USERS - 1) // TILE_USERS
    n_tiles = (n_items + TILE_ITEMS - 1) // TILE_ITEMS
    if n_ub >= n_sm:
        return 1
    return int(max(1, min(n_tiles, 256, (2 * n_sm + n_ub - 1) // n_ub)))


def score_topk(user_split, user_scale, user_bias, item_split, item_meta, n_users, n_items, d_pad, k, n_splits=None,
               item_id_offset=0, n_users_live=None):
    """K2+K3 fused.  Returns (cand_score [U, n_splits, k] f32, cand_item [U, n_splits, k] i32).
    n_users_live: device int32 tensor; only its first element's worth of user rows is processed."""
    lib = require_cuda()
    if n_splits is None:
        n_splits = default_splits(n_users, n_items)
    dev = user_split.device
    cand_score = torch.empty((n_users, n_splits, k), dtype=torch.float32, device=dev)
    cand_item = torch.empty((n_users, n_splits, k), dtype=torch.int32, device=dev)
    rc = lib.trk_score_topk_f16x3(_p(user_split), _p(user_scale), _p(user_bias), _p(item_split), _p(item_meta),
                                  n_users, n_items, int(d_pad), int(k), int(n_splits), int(item_id_offset),
                                  _p(cand_score), _p(cand_item), _p(n_users_live), _stream())
    _lib.check(rc, 'trk_score_topk_f16x3')
    return cand_score, cand_item


def score_dense_tc(user_split, user_scale, user_bias, item_split, item_meta, n_users, n_items, d_pad, out=None):
    lib = require_cuda()
    if out is None:
        out = torch.empty((n_users, n_items), dtype=torch.float32, device=user_split.device)
    rc = lib.trk_score_dense_f16x3(_p(user_split), _p(user_scale), _p(user_bias), _p(item_split), _p(item_meta),
                                   n_users, n_items, int(d_pad), _p(out), out.stride(0), _stream())
    _lib.check(rc, 'trk_score_dense_f16x3')
    return out


class PackedTopK(object):
    """Top-k result of a batch of users in the layout the multi-GPU exchange sends: int32 [n_users, 2k], row u =
    k scores (float32 bits) then k item ids.  `scores` / `items` are views."""

    def __init__(self, n_users, k, device, buf=None):
        self.k = int(k)
        self.buf = torch.empty((int(n_users), 2 * self.k), dtype=torch.int32, device=device) if buf is None else buf

    @property
    def n_users(self):
        return int(self.buf.shape[0])

    @property
    def scores(self):
        return self.buf[:, :self.k].view(torch.float32)

    @property
    def items(self):
        return self.buf[:, self.k:]

    def score_ptr(self):
        return ctypes.c_void_p(self.buf.data_ptr())

    def item_ptr(self):
        return ctypes.c_void_p(self.buf.data_ptr() + 4 * self.k)


def topk_merge(cand_score, cand_item, k_out, out=None, n_users_live=None):
    """[U, L, k_in] candidate lists -> PackedTopK [U, k_out] (top scores, top item ids)."""
    lib = require_cuda()
    n_users, n_lists, k_in = cand_score.shape
    cand_score, cand_item = cand_score.contiguous(), cand_item.contiguous()
    if out is None:
        out = PackedTopK(n_users, k_out, cand_score.device)
    rc = lib.trk_topk_merge(_p(cand_score), _p(cand_item), n_users, n_lists, k_in, int(k_out), n_lists * k_in, k_in,
                            out.score_ptr(), out.item_ptr(), 2 * out.k, _p(n_users_live), 0, _stream())
    _lib.check(rc, 'trk_topk_merge')
    return out


def topk_merge_received(recv, n_users, n_lists, k, dedup=False):
    """Merge of int32 [n_lists, n_users, 2k] -> PackedTopK [n_users, k]: the exchange receive buffer (list l = the
    candidates rank l found for THIS rank's user slice), or -- with dedup -- the per-taste results of a
    mixture-of-tastes model (list t = the top-k of taste t; an item named by several tastes keeps its best score)."""
    lib = require_cuda()
    out = PackedTopK(n_users, k, recv.device)
    if n_users == 0:
        return out
    base = recv.data_ptr()
    rc = lib.trk_topk_merge(ctypes.c_void_p(base), ctypes.c_void_p(base + 4 * k), n_users, int(n_lists), int(k), int(k),
                            2 * k, n_users * 2 * k, out.score_ptr(), out.item_ptr(), 2 * k, None, 1 if dedup else 0,
                            _stream())
    _lib.check(rc, 'trk_topk_merge')
    return out


# ---------------------------------------------------------------------------------------------------------------
# filter form of the fused top-k: 1 tensor pass + exact fp32 re-scoring of the survivors
# ---------------------------------------------------------------------------------------------------------------
def filter_max_k():
    return int(require_cuda().trk_score_filter_max_k())


def filter_list_width():
    return int(require_cuda().trk_score_filter_list_width())


def operand_stats(split, scale, d_pad, want_norm=True, stats=None):
    """Row norms (upper bounds) of a split operand and, if `stats` (zeroed float32[3]) is given, the global max norm /
    max row scale by device-side atomic max."""
    lib = require_cuda()
    rows = split.shape[0]
    norm = torch.empty((rows,), dtype=torch.float32, device=split.device) if want_norm else None
    rc = lib.trk_operand_stats(_p(split), _p(scale), rows, int(d_pad), _p(norm), _p(stats), _stream())
    _lib.check(rc, 'trk_operand_stats')
    return norm


def rescale_hi_global(split, scale, stats, d_pad, perm=None):
    lib = require_cuda()
    rows = split.shape[0]
    out = torch.empty((rows, int(d_pad)), dtype=torch.float16, device=split.device)
    rc = lib.trk_rescale_hi_global(_p(split), _p(scale), _p(stats), _p(perm), rows, int(d_pad), _p(out), _stream())
    _lib.check(rc, 'trk_rescale_hi_global')
    return out


def pack_item_bias(item_bias, n_items, stats, device, perm=None, want_min=False):
    """Returns (bias in processing order padded with -inf, max bias per block of 128 positions[, min bias per block])."""
    lib = require_cuda()
    n_pad = padded_items(n_items)
    out = torch.empty((n_pad,), dtype=torch.float32, device=device)
    block_max = torch.empty((n_pad // 128,), dtype=torch.float32, device=device)
    block_min = torch.empty((n_pad // 128,), dtype=torch.float32, device=device) if want_min else None
    rc = lib.trk_pack_item_bias(_p(item_bias), _p(perm), n_items, _p(out), n_pad, _p(stats), _p(block_max),
                                _p(block_min), _stream())
    _lib.check(rc, 'trk_pack_item_bias')
    if want_min:
        return out, block_max, block_min
    return out, block_max


def bias_processing_order(item_bias):
    """Items by DESCENDING bias (stable, so the order is deterministic): int32 perm[position] = item index.
    Highest biases first: the running k-th best rises early, and every later block starts below it by its bias gap."""
    if item_bias is None:
        return None
    if BIAS_ORDER == 'torch':
        return torch.sort(item_bias, descending=True, stable=True).indices.to(torch.int32)
    # own kernels: the reference ranks of the 1 x I bias row (K3: value descending, lower index first on ties), inverted
    lib = require_cuda()
    n = int(item_bias.numel())
    ranks = rank_full(item_bias.contiguous().view(1, n))
    order = torch.empty((n,), dtype=torch.int32, device=item_bias.device)
    rc = lib.trk_order_from_ranks(_p(ranks), n, _p(order), _stream())
    _lib.check(rc, 'trk_order_from_ranks')
    return order


def score_filter(user_split, user_scale, user_bias, user_norm, item_hi, item_stats, item_bias_pad, block_bias_max,
                 item_perm, n_users, n_items, d_pad, k, n_splits=None, item_id_offset=0, block_bias_min=None):
    """Filter pass.  Returns (cand_score, cand_item [U, n_splits, 16], theta [U, n_splits])."""
    lib = require_cuda()
    if n_splits is None:
        n_splits = default_splits(n_users, n_items)
    dev = user_split.device
    width = filter_list_width()      # one list of `width` candidates per (user, split)
    cand_s = torch.empty((n_users, n_splits, width), dtype=torch.float32, device=dev)
    cand_i = torch.empty((n_users, n_splits, width), dtype=torch.int32, device=dev)
    theta = torch.empty((n_users, n_splits), dtype=torch.float32, device=dev)
    rc = lib.trk_score_filter_f16(_p(user_split), _p(user_scale), _p(user_bias), _p(user_norm), _p(item_hi),
                                  _p(item_stats), _p(item_bias_pad), _p(block_bias_max), _p(block_bias_min),
                                  _p(item_perm), n_users, n_items, int(d_pad), int(k), int(n_splits),
                                  int(item_id_offset), _p(cand_s), _p(cand_i), _p(theta), _stream())
    _lib.check(rc, 'trk_score_filter_f16')
    return cand_s, cand_i, theta


def rescore_topk(users, items, cand_item, theta, user_norm, item_stats, k, item_id_offset=0, out=None):
    """Survivors of the filter -> (PackedTopK [U, k], flags int32 [U]); flags mark users the certificate rejects."""
    lib = require_cuda()
    n_users = users.n_rows
    n_lists = theta.numel() // max(n_users, 1)
    width = cand_item.shape[-1]
    dev = users.split.device
    if out is None:
        out = PackedTopK(n_users, k, dev)
    out_f = torch.empty((n_users,), dtype=torch.int32, device=dev)
    rc = lib.trk_rescore_topk_split(_p(users.split), _p(users.scale), _p(items.split), _p(items.scale), _p(users.bias),
                                    _p(items.bias), _p(cand_item), _p(theta), _p(user_norm), _p(item_stats), n_users,
                                    items.n_rows, int(users.d_pad), n_lists, width, int(k), int(item_id_offset),
                                    out.score_ptr(), out.item_ptr(), 2 * out.k, _p(out_f), _stream())
    _lib.check(rc, 'trk_rescore_topk_split')
    return out, out_f


class _HostResults(object):
    """Device -> host copies of results land in page-locked buffers (a pageable destination costs a staging copy and
    page faults: ~3x the time of the PCIe transfer for the 80 MB top-k of 1M users).  The returned numpy arrays ARE the
    pinned buffers; a buffer is recycled only after the array handed out for it (and every view of it) has been
    garbage collected, so results never alias."""

    max_pinned_bytes = 1 << 30   # larger results (a dense [U, I] matrix) use an ordinary pageable copy

    def __init__(self, max_idle_bytes=2 << 30):
        self._idle = []          # [(pinned tensor, weakref to the ndarray handed out)]
        self._max_idle_bytes = max_idle_bytes

    def _take(self, shape, dtype):
        keep, found = [], None
        for buf, ref in self._idle:
            if ref() is not None:
                keep.append((buf, ref))
            elif found is None and buf.dtype == dtype and tuple(buf.shape) == tuple(shape):
                found = buf
            else:
                keep.append((buf, ref))
        self._idle = keep
        if found is None:
            found = torch.empty(tuple(shape), dtype=dtype, pin_memory=True)
        return found

    def fetch(self, *tensors):
        """numpy copies of CUDA tensors, transferred together (one synchronisation)."""
        import weakref
        bufs = []
        for t in tensors:
            t = t.detach().contiguous()
            nbytes = t.numel() * t.element_size()
            if nbytes == 0 or nbytes > self.max_pinned_bytes:
                bufs.append(t.cpu().numpy())      # empty, or too large to page-lock: ordinary pageable copy
                continue
            buf = self._take(t.shape, t.dtype)
            buf.copy_(t, non_blocking=True)
            bufs.append(buf)
        torch.cuda.current_stream().synchronize()
        out = []
        for buf in bufs:
            if isinstance(buf, np.ndarray):
                out.append(buf)
                continue
            arr = buf.numpy()
            self._idle.append((buf, weakref.ref(arr)))
            out.append(arr)
        # bound what sits in the pool once its arrays are gone
        idle_bytes, keep = 0, []
        for buf, ref in reversed(self._idle):
            nbytes = buf.numel() * buf.element_size()
            if ref() is None and idle_bytes + nbytes > self._max_idle_bytes:
                continue
            if ref() is None:
                idle_bytes += nbytes
            keep.append((buf, ref))
        self._idle = list(reversed(keep))
        return out


_host_results = _HostResults()


def to_host(*tensors):
    """CUDA tensors -> numpy arrays through page-locked buffers (see _HostResults)."""
    require_cuda()
    out = _host_results.fetch(*tensors)
    return out[0] if len(out) == 1 else tuple(out)


class SideOperands(object):
    """Everything the score kernels need from one side (users or items), all resident on the device.
    norm: row norms (users, filter path); stats: float32[3] max norm / max scale / max |bias| (items, filter path)."""

    def __init__(self, repr_f32, split, scale, bias, n_rows, d, d_pad, norm=None, stats=None):
        self.repr_f32, self.split, self.scale, self.bias = repr_f32, split, scale, bias
        self.n_rows, self.d, self.d_pad = n_rows, d, d_pad
        self.norm, self.stats = norm, stats

    def rows(self, r0, r1):
        """The operands of rows [r0, r1) (views)."""
        cut = lambda t: None if t is None else t[r0:r1]   # noqa: E731
        return SideOperands(cut(self.repr_f32), cut(self.split), cut(self.scale), cut(self.bias), r1 - r0, self.d,
                            self.d_pad, norm=cut(self.norm), stats=self.stats)


def topk_exact(users, items, k, n_splits=None, item_id_offset=0, out=None, n_users_live=None):
    """Exact 3-pass fused kernel + merge -> PackedTopK [U, k]."""
    meta = pack_item_meta(items.scale, items.bias, items.n_rows)
    cs, ci = score_topk(users.split, users.scale, users.bias, items.split, meta, users.n_rows, items.n_rows,
                        users.d_pad, k, n_splits=n_splits, item_id_offset=item_id_offset, n_users_live=n_users_live)
    return topk_merge(cs, ci, k, out=out, n_users_live=n_users_live)


class FilterItems(object):
    """Item-side inputs of the filter kernel, derived once per call from the K1 outputs (items.stats: max norm / max
    scale already formed in K1's epilogue; operands from a user-defined graph get them from trk_operand_stats)."""

    def __init__(self, items):
        dev = items.split.device
        if items.stats is not None:
            self.stats = items.stats
        else:
            self.stats = torch.zeros((3,), dtype=torch.float32, device=dev)
            operand_stats(items.split, items.scale, items.d_pad, want_norm=False, stats=self.stats)
        self.perm = bias_processing_order(items.bias)
        self.hi = rescale_hi_global(items.split, items.scale, self.stats, items.d_pad, perm=self.perm)
        self.bias_pad, self.block_max, self.block_min = pack_item_bias(items.bias, items.n_rows, self.stats, dev,
                                                                       perm=self.perm, want_min=True)


def filter_and_rescore(users, items, fitems, user_norm, k, n_splits=None, item_id_offset=0, out=None):
    """(PackedTopK [U, k], flags [U]) -- flags mark users the certificate did not cover."""
    _, ci, theta = score_filter(users.split, users.scale, users.bias, user_norm, fitems.hi, fitems.stats,
                                fitems.bias_pad, fitems.block_max, fitems.perm, users.n_rows, items.n_rows,
                                users.d_pad, k, n_splits=n_splits, item_id_offset=item_id_offset,
                                block_bias_min=fitems.block_min)
    return rescore_topk(users, items, ci, theta, user_norm, fitems.stats, k, item_id_offset=item_id_offset, out=out)


def fallback_capacity(n_users):
    """Rows the device-side fallback can hold (the exact kernel is launched over this many rows and skips the unused
    ones): 1/8 of the batch, at least 1024, whole 128-row user blocks.  More flagged rows than this = a tie-heavy
    batch; the host layer then re-runs the whole batch through the exact kernel."""
    cap = min(int(n_users), max(1024, int(n_users) // 8))
    return ((cap + 127) // 128) * 128


FALLBACK_SMALL_ROWS = 1024     # the small re-scoring tier: at most this many flagged rows, many item splits


def _ptr_at(t, index):
    return ctypes.c_void_p(t.data_ptr() + index * t.element_size())


def rerun_uncertified(users, items, bad, top, k, item_id_offset=0):
    """Users flagged by the certificate go through the exact kernel WITHOUT a host round trip: the flagged rows are
    compacted on the device, their operands gathered into a fixed-capacity buffer, the exact kernel runs over that buffer
    with the device-side count and the rows are scattered back into `top`.  Two tiers are launched, exactly one does
    work (decided on the device): up to FALLBACK_SMALL_ROWS rows (the normal case: ~0.01 % of the users) with as many
    item splits as it takes to fill the machine from one or two user blocks, or up to `capacity` rows with few splits.
    Returns (counters, capacity); counters[0] = flagged rows, > capacity means overflow (the caller checks it at its
    next synchronisation)."""
    lib = require_cuda()
    dev = users.split.device
    cap = fallback_capacity(users.n_rows)
    small = min(cap, FALLBACK_SMALL_ROWS)
    idx = torch.empty((cap,), dtype=torch.int32, device=dev)
    counters = torch.empty((4,), dtype=torch.int32, device=dev)
    rc = lib.trk_select_flagged_rows(_p(bad), users.n_rows, _p(idx), cap, _p(counters), _stream())
    _lib.check(rc, 'trk_select_flagged_rows')
    sub_split = torch.empty((cap, 2 * users.d_pad), dtype=torch.float16, device=dev)
    sub_scale = torch.empty((cap,), dtype=torch.float32, device=dev)
    sub_bias = None if users.bias is None else torch.empty((cap,), dtype=torch.float32, device=dev)
    rc = lib.trk_gather_operand_rows(_p(idx), _p(counters), cap, small, _p(users.split), _p(users.scale),
                                     _p(users.bias), int(users.d_pad), _p(sub_split), _p(sub_scale), _p(sub_bias),
                                     _stream())
    _lib.check(rc, 'trk_gather_operand_rows')
    sub = SideOperands(None, sub_split, sub_scale, sub_bias, cap, users.d, users.d_pad)
    tiers = [(sub.rows(0, small), small, 2, default_splits(2 * TILE_USERS, items.n_rows))]
    if cap > small:
        tiers.append((sub, cap, 3, None))
    for tier_rows, n_rows, slot, n_splits in tiers:
        live = counters[slot:slot + 1]
        exact = topk_exact(tier_rows, items, k, n_splits=n_splits, item_id_offset=item_id_offset, n_users_live=live)
        rc = lib.trk_scatter_topk_rows(_p(idx), _ptr_at(counters, slot), n_rows, exact.score_ptr(), exact.item_ptr(),
                                       2 * exact.k, int(k), top.score_ptr(), top.item_ptr(), 2 * top.k, _stream())
        _lib.check(rc, 'trk_scatter_topk_rows')
    return counters, cap


def topk_filter(users, items, k, n_splits=None, item_id_offset=0, fitems=None):
    """Filter form: one tensor pass + re-scoring from the split operands; users whose error bound cannot be certified
    (buffer overflow under massive ties, bound violated) are re-run through the exact kernel on the device.
    Returns (PackedTopK, counters device int32[2], capacity)."""
    user_norm = users.norm if users.norm is not None else operand_stats(users.split, users.scale, users.d_pad)
    if fitems is None:
        fitems = FilterItems(items)
    top, bad = filter_and_rescore(users, items, fitems, user_norm, k, n_splits=n_splits, item_id_offset=item_id_offset)
    counters, cap = rerun_uncertified(users, items, bad, top, k, item_id_offset=item_id_offset)
    return top, counters, cap
