"""Prediction plugin graphs: (user_repr, item_repr) -> scores (API of tensorrec/prediction_graphs.py).

`connect_dense_prediction_graph` / `connect_serial_prediction_graph` keep the reference's names and argument
meaning (prediction_graphs.py:11-40).  They accept torch tensors (differentiable: this is what the training step
runs) -- and, for the dense form, CUDA tensors or numpy arrays are scored by the hand-written kernels when no
gradient is needed (the predict / predict_rank hot path; TensorRec lowers built-in graphs by `b200_kind`)."""
import numpy as np
import torch


def _l2_normalize(x, eps=1e-12):
    return x * torch.rsqrt(torch.clamp(torch.sum(x * x, dim=1, keepdim=True), min=eps))


def _needs_grad(*tensors):
    """True -> some input carries a gradient: differentiable torch ops (training step); False -> kernels on CUDA."""
    from .session_management import in_training_step
    if in_training_step():
        return True
    for t in tensors:
        if isinstance(t, torch.Tensor):
            if t.requires_grad and torch.is_grad_enabled():
                return True
    return False


def _as_device_f32(x):
    """numpy / CPU tensor -> float32 CUDA tensor for the kernel path (raises without a CUDA device)."""
    from . import kernels
    kernels.require_cuda()
    if isinstance(x, np.ndarray):
        x = torch.from_numpy(np.ascontiguousarray(x, dtype=np.float32))
    return x.to(device='cuda', dtype=torch.float32).contiguous()


def _serial_operands(tf_user_representation, tf_item_representation, tf_x_user, tf_x_item):
    """The serial forms are part of the training step (torch ops); plain numpy arrays -- what the reference's own
    tests pass in -- are accepted too and become tensors on the representations' device."""
    def tensor(x, index=False, like=None):
        if not isinstance(x, torch.Tensor):
            x = torch.as_tensor(np.asarray(x))
            if not index and x.dtype == torch.float64:
                x = x.to(torch.float32)
        if index:
            x = x.to(torch.long)
        return x.to(like.device) if like is not None else x

    users = tensor(tf_user_representation)
    items = tensor(tf_item_representation, like=users)
    return users, items, tensor(tf_x_user, index=True, like=users), tensor(tf_x_item, index=True, like=users)


class AbstractPredictionGraph(object):
    b200_kind = None

    def connect_dense_prediction_graph(self, tf_user_representation, tf_item_representation):
        pass

    def connect_serial_prediction_graph(self, tf_user_representation, tf_item_representation, tf_x_user, tf_x_item):
        pass


class DotProductPredictionGraph(AbstractPredictionGraph):
    """prediction = user_repr . item_repr (prediction_graphs.py:43-55)."""
    b200_kind = 'dot'

    def connect_dense_prediction_graph(self, tf_user_representation, tf_item_representation):
        if _needs_grad(tf_user_representation, tf_item_representation):
            return tf_user_representation @ tf_item_representation.t()
        from . import kernels
        return kernels.score_exact(_as_device_f32(tf_user_representation), _as_device_f32(tf_item_representation))

    def connect_serial_prediction_graph(self, tf_user_representation, tf_item_representation, tf_x_user, tf_x_item):
        users, items, x_user, x_item = _serial_operands(tf_user_representation, tf_item_representation, tf_x_user,
                                                        tf_x_item)
        return torch.sum(users[x_user] * items[x_item], dim=1)


class CosineSimilarityPredictionGraph(AbstractPredictionGraph):
    """prediction = cos(user_repr, item_repr) (prediction_graphs.py:58-72)."""
    b200_kind = 'cosine'

    def connect_dense_prediction_graph(self, tf_user_representation, tf_item_representation):
        from .recommendation_graphs import relative_cosine
        return relative_cosine(tf_tensor_1=tf_user_representation, tf_tensor_2=tf_item_representation)

    def connect_serial_prediction_graph(self, tf_user_representation, tf_item_representation, tf_x_user, tf_x_item):
        users, items, x_user, x_item = _serial_operands(tf_user_representation, tf_item_representation, tf_x_user,
                                                        tf_x_item)
        return torch.sum(_l2_normalize(users)[x_user] * _l2_normalize(items)[x_item], dim=1)


class EuclideanSimilarityPredictionGraph(AbstractPredictionGraph):
    """prediction = -sqrt(max(|u|^2 - 2 u.i + |i|^2, 1e-16)) (prediction_graphs.py:75-117)."""
    b200_kind = 'euclidean'
    epsilon = 1e-16

    def connect_dense_prediction_graph(self, tf_user_representation, tf_item_representation):
        if _needs_grad(tf_user_representation, tf_item_representation):
            r_user = torch.sum(tf_user_representation ** 2, 1, keepdim=True)
            r_item = torch.sum(tf_item_representation ** 2, 1, keepdim=True)
            distance = r_user - 2.0 * (tf_user_representation @ tf_item_representation.t()) + r_item.t()
            return -1.0 * torch.sqrt(torch.clamp(distance, min=self.epsilon))
        from . import kernels
        return kernels.score_exact(_as_device_f32(tf_user_representation), _as_device_f32(tf_item_representation),
                                   mode=1)

    def connect_serial_prediction_graph(self, tf_user_representation, tf_item_representation, tf_x_user, tf_x_item):
        users, items, x_user, x_item = _serial_operands(tf_user_representation, tf_item_representation, tf_x_user,
                                                        tf_x_item)
        distance = torch.clamp(torch.sum((users[x_user] - items[x_item]) ** 2, dim=1), min=self.epsilon)
        return -1.0 * torch.sqrt(distance)
