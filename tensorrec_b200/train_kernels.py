"""The sampled-rank training step on hand-written kernels (BASELINE config #4).

For the model family the reference's WMRB examples use -- LinearRepresentationGraph on both sides, DotProductPredictionGraph,
WMRBLossGraph or BalancedWMRBLossGraph, one taste, no attention -- one Adam step is, through the C ABI:

    K1  trk_csr_gather_reduce_f32      user / item representations             (tensorrec/representation_graphs.py:40)
        trk_csr_project_biases_f32     projected biases                        (tensorrec/recommendation_graphs.py:4-19)
        trk_f32_to_bf16                [bf16 form] representations rounded once, halving the gather traffic of the step
        trk_sample_items               n_sampled_items item ids per user       (tensorrec/util.py:12-21)
        trk_wmrb_step                  serial predictions of the interactions and of the samples, WMRB loss, and the
                                       gradient with respect to representations and projected biases, fused
    K1^T trk_csr_gather_reduce_f32     on the transposed CSR: weight gradients (the gradient of sparse_tensor_dense_matmul)
        trk_csr_project_biases_f32     on the transposed CSR: feature-bias gradients
        trk_adam_step_f32              L2 term + Adam moments + parameter step (tensorrec/tensorrec.py:487-489)

Every other model family trains through the torch-autograd mirror of the reference's graph functions
(TensorRec._training_losses); TENSORREC_B200_TRAIN_PATH=torch forces that path."""
import ctypes
import os

import numpy as np
import scipy.sparse as sp
import torch

from . import _lib, kernels
from .kernels import _p, _stream

TRAIN_PATH = os.environ.get('TENSORREC_B200_TRAIN_PATH', 'auto')       # 'auto' | 'torch'
TRAIN_DTYPE = os.environ.get('TENSORREC_B200_TRAIN_DTYPE', 'f32')      # 'f32' | 'bf16' (representations only)

ADAM_BETA1, ADAM_BETA2, ADAM_EPSILON = 0.9, 0.999, 1e-8                # tf.train.AdamOptimizer defaults


def sample_items_device(n_items, n_users, n_sampled_items, replace, seed, step, device):
    """int32 [n_users, n_sampled_items] on the device (trk_sample_items)."""
    lib = kernels.require_cuda()
    if (not replace) and n_sampled_items > n_items:
        raise ValueError("Cannot take a larger sample than population when 'replace=False'")
    out = torch.empty((int(n_users), int(n_sampled_items)), dtype=torch.int32, device=device)
    rc = lib.trk_sample_items(int(n_users), int(n_items), int(n_sampled_items), 1 if replace else 0,
                              ctypes.c_uint64(int(seed) & (2 ** 64 - 1)), ctypes.c_uint32(int(step) & 0xffffffff),
                              _p(out), _stream())
    _lib.check(rc, 'trk_sample_items')
    return out


def sample_items_host(n_items, n_users, n_sampled_items, replace, seed, step):
    """The same sample computed on the host from the same Philox stream (trk_sample_stream_u64): the statement of what
    the device kernel must produce, used by the tests.  Pure Python: small sizes only."""
    lib = _lib.load()
    out = np.empty((n_users, n_sampled_items), dtype=np.int32)
    for u in range(n_users):
        chosen = []
        for j in range(n_sampled_items):
            r = int(lib.trk_sample_stream_u64(ctypes.c_uint64(seed), ctypes.c_uint32(step), ctypes.c_uint32(u),
                                              ctypes.c_uint32(j)))
            if replace:
                chosen.append((r * n_items) >> 64)
            else:
                top = n_items - n_sampled_items + j            # Floyd: t uniform in [0, top]
                t = (r * (top + 1)) >> 64
                chosen.append(top if t in chosen else t)
        out[u] = chosen
    return out


def eligible(model):
    """Does the kernel training step cover this model?"""
    from .loss_graphs import WMRBLossGraph, BalancedWMRBLossGraph
    from .prediction_graphs import DotProductPredictionGraph
    from .representation_graphs import LinearRepresentationGraph
    return (TRAIN_PATH != 'torch'
            and type(model.user_repr_graph_factory) is LinearRepresentationGraph
            and type(model.item_repr_graph_factory) is LinearRepresentationGraph
            and type(model.prediction_graph_factory) is DotProductPredictionGraph
            and type(model.loss_graph_factory) in (WMRBLossGraph, BalancedWMRBLossGraph)
            and model.n_tastes == 1 and model.attention_graph_factory is None
            and model.n_components % 4 == 0 and 4 <= model.n_components <= 512)


class WmrbStep(object):
    """State of the kernel training path of one model: Adam moments per weight, the step counter of the sampler's
    stream and of Adam's bias correction."""

    def __init__(self, model, device, seed=None, bf16=None):
        self.model, self.device = model, device
        self.seed = int(np.random.SeedSequence().generate_state(2, dtype=np.uint32).view(np.uint64)[0]) \
            if seed is None else int(seed)
        self.bf16 = (TRAIN_DTYPE == 'bf16') if bf16 is None else bool(bf16)
        self.t = 0
        self.beta_powers = (np.float32(1.0), np.float32(1.0))    # beta1^t, beta2^t as float32 variables (TensorFlow)
        self.moments = {}            # weight name -> (m, v)
        self.last = {}
        self.marks = None            # bench.py: list that receives (phase name, CUDA event) pairs of a step

    def _mark(self, name):
        if self.marks is not None:
            e = torch.cuda.Event(enable_timing=True)
            e.record()
            self.marks.append((name, e))

    # -- weights ------------------------------------------------------------------------------------------
    def _weight(self, name, shape, init):
        store = self.model._variables
        if name not in store:
            store[name] = init().to(self.device).requires_grad_(True)
        w = store[name]
        if w.device != self.device:
            w = w.detach().to(self.device).requires_grad_(True)
            store[name] = w
        if tuple(w.shape) != tuple(shape):
            raise ValueError('weight {!r} has shape {} but the inputs need {}'.format(name, tuple(w.shape), shape))
        return w

    def _weights(self, n_user_features, n_item_features):
        d = self.model.n_components

        def normal_rows(n):       # representation_graphs.py:35-36: random_normal rows, L2-normalised
            def init():
                w = torch.randn(n, d, dtype=torch.float32)
                return w * torch.rsqrt(torch.clamp((w * w).sum(dim=1, keepdim=True), min=1e-12))
            return init

        names = ['linear_weights_item', 'linear_weights_user_0']      # creation order of the reference's graph
        ws = {'linear_weights_item': self._weight('linear_weights_item', (n_item_features, d), normal_rows(n_item_features)),
              'linear_weights_user_0': self._weight('linear_weights_user_0', (n_user_features, d),
                                                    normal_rows(n_user_features))}
        if self.model.biased:         # recommendation_graphs.py:11: zeros
            for name, n in (('feature_biases_user', n_user_features), ('feature_biases_item', n_item_features)):
                ws[name] = self._weight(name, (n, 1), lambda n=n: torch.zeros(n, 1, dtype=torch.float32))
                names.append(name)
        return names, ws

    # -- one step -----------------------------------------------------------------------------------------
    def step(self, interactions_in, user_in, item_in, n_sampled_items, learning_rate, l2, samples=None):
        """One Adam step on sum(WMRB loss) + l2 * sum_w 0.5 |w|^2.  `l2` is the coefficient of the L2 term in the
        SUMMED loss (the reference adds alpha * reg to every element of its loss vector, tensorrec.py:488, so it is
        n_positive_interactions * batched_alpha).  Returns the device tensors of the step (loss, pred_serial)."""
        lib = kernels.require_cuda()
        from .loss_graphs import BalancedWMRBLossGraph
        dev, d = self.device, self.model.n_components
        n_users, n_items = user_in.shape[0], item_in.shape[0]
        ucsr, icsr = user_in.device_csr(dev), item_in.device_csr(dev)
        ucsr_t, icsr_t = user_in.device_csr_t(dev), item_in.device_csr_t(dev)
        inter = interactions_in.device_csr(dev)
        names, ws = self._weights(user_in.shape[1], item_in.shape[1])
        w_user, w_item = ws['linear_weights_user_0'].detach(), ws['linear_weights_item'].detach()

        self._mark('start')
        # forward: representations and projected biases (K1)
        user_repr, _, _ = kernels.gather_reduce(ucsr, w_user, want_f32=True)
        item_repr, _, _ = kernels.gather_reduce(icsr, w_item, want_f32=True)
        ub = ib = None
        if self.model.biased:
            ub = kernels.project_biases(ucsr, ws['feature_biases_user'].detach().reshape(-1))
            ib = kernels.project_biases(icsr, ws['feature_biases_item'].detach().reshape(-1))
        repr_u, repr_i = user_repr, item_repr
        if self.bf16:
            repr_u = torch.empty(user_repr.shape, dtype=torch.bfloat16, device=dev)
            repr_i = torch.empty(item_repr.shape, dtype=torch.bfloat16, device=dev)
            _lib.check(lib.trk_f32_to_bf16(_p(user_repr), user_repr.numel(), _p(repr_u), _stream()), 'trk_f32_to_bf16')
            _lib.check(lib.trk_f32_to_bf16(_p(item_repr), item_repr.numel(), _p(repr_i), _stream()), 'trk_f32_to_bf16')

        self._mark('representations')
        if samples is None:
            samples = sample_items_device(n_items, n_users, n_sampled_items,
                                          self.model.loss_graph_factory.is_sampled_with_replacement, self.seed, self.t, dev)
        weight_sum = None
        if type(self.model.loss_graph_factory) is BalancedWMRBLossGraph:
            weight_sum = interactions_in.positive_item_sums(dev)

        self._mark('sampler')
        nnz = inter.nnz
        loss = torch.empty((nnz,), dtype=torch.float32, device=dev)
        pred = torch.empty((nnz,), dtype=torch.float32, device=dev)
        coef = torch.empty((nnz,), dtype=torch.float32, device=dev)
        d_user_repr = torch.empty((n_users, d), dtype=torch.float32, device=dev)
        d_item_repr = torch.zeros((n_items, d), dtype=torch.float32, device=dev)
        d_ub = torch.empty((n_users,), dtype=torch.float32, device=dev) if self.model.biased else None
        d_ib = torch.zeros((n_items,), dtype=torch.float32, device=dev) if self.model.biased else None
        rc = lib.trk_wmrb_step(_p(repr_u), _p(repr_i), 1 if self.bf16 else 0, _p(ub), _p(ib), _p(inter.indptr),
                               _p(inter.col), _p(inter.val), _p(weight_sum), _p(samples), n_users, n_items, d,
                               int(samples.shape[1]), _p(loss), _p(pred), _p(coef), _p(d_user_repr), _p(d_ub),
                               _p(d_item_repr), _p(d_ib), _stream())
        _lib.check(rc, 'trk_wmrb_step')
        self._mark('wmrb_step')

        # backward through the sparse x dense products: K1 on the transposed CSR
        grads = {'linear_weights_user_0': kernels.gather_reduce(ucsr_t, d_user_repr, want_f32=True)[0],
                 'linear_weights_item': kernels.gather_reduce(icsr_t, d_item_repr, want_f32=True)[0]}
        if self.model.biased:
            grads['feature_biases_user'] = kernels.project_biases(ucsr_t, d_ub)
            grads['feature_biases_item'] = kernels.project_biases(icsr_t, d_ib)
        self.last = {'loss': loss, 'pred_serial': pred, 'grads': grads, 'samples': samples, 'inter_val': inter.val}

        self._mark('weight_gradients')
        # Adam
        self.t += 1
        f32 = np.float32
        self.beta_powers = (f32(self.beta_powers[0] * f32(ADAM_BETA1)), f32(self.beta_powers[1] * f32(ADAM_BETA2)))
        lr_t = float(f32(learning_rate) * np.sqrt(f32(1.0) - self.beta_powers[1]) / (f32(1.0) - self.beta_powers[0]))
        for name in names:
            w = ws[name]
            if name not in self.moments:
                self.moments[name] = (torch.zeros_like(w, requires_grad=False), torch.zeros_like(w, requires_grad=False))
            m, v = self.moments[name]
            rc = lib.trk_adam_step_f32(_p(w), _p(grads[name]), _p(m), _p(v), w.numel(), ctypes.c_float(lr_t),
                                       ctypes.c_float(ADAM_BETA1), ctypes.c_float(ADAM_BETA2),
                                       ctypes.c_float(ADAM_EPSILON), ctypes.c_float(l2), _stream())
            _lib.check(rc, 'trk_adam_step_f32')
        self._mark('adam')
        return loss, pred


def positive_item_sums(matrix, n_items):
    """BalancedWMRBLossGraph's listening_sum_per_item (tensorrec/loss_graphs.py:201): sum of the positive interaction
    values per item, float32 (host side, once per interaction matrix)."""
    coo = matrix if isinstance(matrix, sp.coo_matrix) else sp.coo_matrix(matrix)
    data = coo.data.astype(np.float32)
    mask = data > 0.0
    out = np.zeros(n_items, dtype=np.float32)
    np.add.at(out, coo.col[mask], data[mask])
    return out
