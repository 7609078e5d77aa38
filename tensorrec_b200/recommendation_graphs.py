"""Glue functions of the recommendation graph (API of tensorrec/recommendation_graphs.py).

Two evaluation modes share these names:
  * training (torch autograd on tensors that require grad): plain differentiable torch ops;
  * prediction (no gradient): numpy arrays / tensors are moved to the CUDA device and evaluated by the hand-written
    kernels through the C ABI -- rank_predictions -> trk_rank_full, relative_cosine / dense predictions ->
    trk_l2_normalize_rows_f32 + trk_score_f32, project_biases -> trk_csr_project_biases_f32.  No CPU fallback."""
import numpy as np
import scipy.sparse as sp
import torch

from .session_management import get_variable


def _needs_grad(*tensors):
    """True -> some input carries a gradient: evaluate with differentiable torch ops (the training step).
    False -> plain arrays / gradient-free tensors: evaluate with the kernels on the CUDA device (no CPU path)."""
    from .session_management import in_training_step
    if in_training_step():
        return True
    for t in tensors:
        if isinstance(t, torch.Tensor):
            if t.requires_grad and torch.is_grad_enabled():
                return True
    return False


def _dev(x, dtype=torch.float32):
    from . import kernels
    kernels.require_cuda()
    if isinstance(x, (list, tuple)):
        x = np.asarray(x)
    if isinstance(x, np.ndarray):
        x = torch.from_numpy(np.ascontiguousarray(x))
    return x.to(device='cuda', dtype=dtype).contiguous()


class _DeferredProjection(object):
    """reduce_sum(features @ biases, 1) of a graph node that is evaluated on demand (tf semantics: values assigned to
    the bias variable after the node was built are seen)."""

    def __init__(self, tf_features, tf_feature_biases):
        self.tf_features, self.tf_feature_biases = tf_features, tf_feature_biases

    def value(self):
        return torch.sum(torch.sparse.mm(self.tf_features, self.tf_feature_biases.detach()), dim=1)

    def eval(self, session=None, feed_dict=None):
        return self.value().cpu().numpy()


def project_biases(tf_features, n_features):
    """recommendation_graphs.py:4-19.  Returns (feature_biases [n_features, 1], projected_biases [n_rows]).

    tf_features: torch sparse tensor (training) or kernels.DeviceCSR / scipy matrix (prediction)."""
    from . import kernels
    if isinstance(tf_features, torch.Tensor):
        from .session_management import in_training_step
        from .sparse_ops import sparse_dense_matmul
        tf_feature_biases = get_variable('feature_biases_{}x{}'.format(tf_features.shape[0], n_features),
                                         lambda: torch.zeros([n_features, 1], device=tf_features.device))
        if in_training_step():
            return tf_feature_biases, torch.sum(sparse_dense_matmul(tf_features, tf_feature_biases), dim=1)
        # built outside a training step (the reference's own test does: build, assign bias values, THEN evaluate,
        # test/test_recommendation_graphs.py:20-40): the projection is evaluated when it is asked for
        return tf_feature_biases, _DeferredProjection(tf_features, tf_feature_biases)
    if sp.issparse(tf_features):
        tf_features = kernels.DeviceCSR.from_scipy(tf_features)
    tf_feature_biases = torch.zeros([n_features, 1], device='cuda')
    return tf_feature_biases, kernels.project_biases(tf_features, tf_feature_biases.view(-1))


def project_biases_with(tf_features, feature_biases):
    """project_biases with given bias values (what the reference test does via assign, test_recommendation_graphs.py:
    35-36), evaluated by trk_csr_project_biases_f32."""
    from . import kernels
    if sp.issparse(tf_features):
        tf_features = kernels.DeviceCSR.from_scipy(tf_features)
    return kernels.project_biases(tf_features, _dev(feature_biases).view(-1))


def split_sparse_tensor_indices(tf_sparse_tensor, n_dimensions):
    """recommendation_graphs.py:22-30."""
    indices = tf_sparse_tensor._indices() if not tf_sparse_tensor.is_coalesced() else tf_sparse_tensor.indices()
    return (indices[i] for i in range(n_dimensions))


def bias_prediction_dense(tf_prediction, tf_projected_user_biases, tf_projected_item_biases):
    """recommendation_graphs.py:33-41: pred + ub[:, None] + ib[None, :], left to right."""
    if not _needs_grad(tf_prediction, tf_projected_user_biases, tf_projected_item_biases):
        tf_prediction = _dev(tf_prediction)
        tf_projected_user_biases = _dev(tf_projected_user_biases)
        tf_projected_item_biases = _dev(tf_projected_item_biases)
    return tf_prediction + tf_projected_user_biases.unsqueeze(1) + tf_projected_item_biases.unsqueeze(0)


def bias_prediction_serial(tf_prediction_serial, tf_projected_user_biases, tf_projected_item_biases, tf_x_user,
                           tf_x_item):
    """recommendation_graphs.py:44-57 (training step: torch ops; numpy inputs are accepted like in the reference's
    tests)."""
    def tensor(x, like=None, index=False):
        if not isinstance(x, torch.Tensor):
            x = torch.as_tensor(np.asarray(x))
        if index:
            x = x.to(torch.long)
        return x.to(like.device) if like is not None else x

    prediction = tensor(tf_prediction_serial)
    user_biases, item_biases = tensor(tf_projected_user_biases, prediction), tensor(tf_projected_item_biases, prediction)
    x_user, x_item = tensor(tf_x_user, prediction, index=True), tensor(tf_x_item, prediction, index=True)
    return prediction + user_biases[x_user].to(prediction.dtype) + item_biases[x_item].to(prediction.dtype)


def densify_sampled_item_predictions(tf_sample_predictions_serial, tf_n_sampled_items, tf_n_users):
    """recommendation_graphs.py:60-70."""
    if not isinstance(tf_sample_predictions_serial, torch.Tensor):
        tf_sample_predictions_serial = torch.as_tensor(np.asarray(tf_sample_predictions_serial))
    return tf_sample_predictions_serial.reshape(int(tf_n_users), int(tf_n_sampled_items))


def rank_predictions(tf_prediction):
    """recommendation_graphs.py:73-82: the double tf.nn.top_k == rank by (score desc, index asc), int32, 1-based.
    Evaluated by the K3 kernel (trk_rank_full)."""
    from . import kernels
    if isinstance(tf_prediction, torch.Tensor):
        tf_prediction = tf_prediction.detach()
    scores = _dev(tf_prediction)
    if scores.dim() == 1:
        scores = scores.unsqueeze(0)
    return kernels.rank_full(scores)


def collapse_mixture_of_tastes(tastes_predictions, tastes_attentions):
    """recommendation_graphs.py:85-109."""
    grad = _needs_grad(*tastes_predictions) or (tastes_attentions is not None and _needs_grad(*tastes_attentions))
    if not grad:
        tastes_predictions = [_dev(p) for p in tastes_predictions]
        tastes_attentions = None if tastes_attentions is None else [_dev(a) for a in tastes_attentions]
    stacked_predictions = torch.stack(list(tastes_predictions))
    if tastes_attentions is not None:
        softmax_attentions = torch.softmax(torch.stack(list(tastes_attentions)), dim=0)
        return torch.sum(stacked_predictions * softmax_attentions, dim=0)
    return torch.max(stacked_predictions, dim=0).values


def relative_cosine(tf_tensor_1, tf_tensor_2):
    """recommendation_graphs.py:112-121: cosine of every row of tensor_1 against every row of tensor_2."""
    if _needs_grad(tf_tensor_1, tf_tensor_2):
        def norm(x):
            return x * torch.rsqrt(torch.clamp(torch.sum(x * x, dim=1, keepdim=True), min=1e-12))
        return norm(tf_tensor_1) @ norm(tf_tensor_2).t()
    from . import kernels
    t1 = kernels.l2_normalize_rows_(_dev(tf_tensor_1).clone())
    t2 = kernels.l2_normalize_rows_(_dev(tf_tensor_2).clone())
    return kernels.score_exact(t1, t2)


def predict_similar_items(prediction_graph_factory, tf_item_representation, tf_similar_items_ids):
    """recommendation_graphs.py:124-137."""
    if not isinstance(tf_item_representation, torch.Tensor) or not tf_item_representation.requires_grad:
        tf_item_representation = _dev(tf_item_representation)
    ids = torch.as_tensor(np.asarray(tf_similar_items_ids), dtype=torch.long, device=tf_item_representation.device)
    gathered_items = tf_item_representation[ids]
    return prediction_graph_factory.connect_dense_prediction_graph(tf_user_representation=gathered_items,
                                                                   tf_item_representation=tf_item_representation)
