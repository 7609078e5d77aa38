"""A `tensorflow` stand-in for code written against the reference's plugin surface.

NOT TensorFlow: the handful of TF 1.x names the reference's tests, README and examples use inside plugin graphs, mapped
eagerly onto torch tensors (the define-by-run contract of tensorrec_b200: weights are created through
session_management.get_variable, see tensorrec_b200/representation_graphs.py).  There is no graph and no session;
`tensor.eval(session=...)` returns the value as a numpy array.  Ops take the device of their tensor operands.

Covered: tf.SparseTensor, tf.Variable(initial_value, name=...), tf.random_normal, tf.sparse_tensor_dense_matmul,
tf.matmul, tf.nn.tanh / relu / sigmoid / l2_normalize, tf.reduce_mean / reduce_sum, tf.abs / square / sqrt / maximum,
tf.global_variables_initializer, tf.reset_default_graph, tf.Tensor.  Importing this module also restores the numpy
aliases the reference's (numpy-1 era) tests use: np.mat, np.int."""
import numpy as np
import torch

from tensorrec_b200.session_management import get_variable, next_anonymous_name

Tensor = torch.Tensor
float32, int64, int32 = torch.float32, torch.int64, torch.int32

if not hasattr(np, 'mat'):
    np.mat = np.asmatrix
for _alias, _target in (('int', int), ('float', float), ('bool', bool)):
    if not hasattr(np, _alias):
        setattr(np, _alias, _target)


def _eval(self, session=None, feed_dict=None):
    """tf.Tensor.eval: there is no deferred graph, the tensor already holds its value."""
    return self.detach().cpu().numpy()


def _assign(self, value, use_locking=None, name=None):
    """tf.Variable.assign: executed immediately (returns None where TF returns an op to run)."""
    with torch.no_grad():
        self.copy_(torch.as_tensor(np.asarray(value), dtype=self.dtype).reshape(self.shape))
    return None


torch.Tensor.eval = _eval
torch.Tensor.assign = _assign


def _t(x, like=None):
    if isinstance(x, torch.Tensor):
        return x
    t = torch.as_tensor(np.asarray(x))
    if t.dtype == torch.float64:
        t = t.to(torch.float32)
    return t.to(like.device) if like is not None else t


def SparseTensor(indices, values, dense_shape):
    """tf.SparseTensor(indices [nnz, n_dims], values [nnz], dense_shape): an (uncoalesced) torch sparse COO tensor."""
    idx = torch.as_tensor(np.asarray(indices), dtype=torch.long).t().contiguous()
    return torch.sparse_coo_tensor(idx, _t(values), size=tuple(int(d) for d in dense_shape), is_coalesced=False,
                                   check_invariants=False)


def Variable(initial_value, name=None, dtype=None, trainable=True):
    """tf.Variable: a trainable tensor registered under its name in the model's variable scope (created once, like a
    graph variable; plugin methods run on every training step).  An unnamed variable gets a name that is stable across
    steps: 'Variable_<plugin call>_<n-th anonymous variable of that call>' (session_management.name_scope)."""
    if name is None:
        name = next_anonymous_name()
    return get_variable(name, lambda: _t(initial_value).to(torch.float32))


def random_normal(shape, mean=0.0, stddev=1.0, dtype=None, seed=None, name=None):
    return torch.randn(*[int(s) for s in shape]) * float(stddev) + float(mean)


def sparse_tensor_dense_matmul(sp_a, b, name=None):
    from tensorrec_b200.sparse_ops import sparse_dense_matmul
    return sparse_dense_matmul(sp_a, _t(b).to(sp_a.device))


def matmul(a, b, transpose_a=False, transpose_b=False, name=None):
    a = _t(a)
    b = _t(b, like=a)
    return (a.t() if transpose_a else a) @ (b.t() if transpose_b else b)


def reduce_mean(x, axis=None, keep_dims=False, name=None):
    x = _t(x)
    return torch.mean(x) if axis is None else torch.mean(x, dim=axis, keepdim=keep_dims)


def reduce_sum(x, axis=None, keep_dims=False, name=None):
    x = _t(x)
    return torch.sum(x) if axis is None else torch.sum(x, dim=axis, keepdim=keep_dims)


def abs(x, name=None):          # noqa: A001  (the TF name)
    return torch.abs(_t(x))


def square(x, name=None):
    return _t(x) ** 2


def sqrt(x, name=None):
    return torch.sqrt(_t(x))


def maximum(x, y, name=None):
    x = _t(x)
    return torch.maximum(x, _t(y, like=x).to(x.dtype))


def global_variables_initializer():
    """Variables are initialised when they are created; kept so `session.run(tf.global_variables_initializer())` reads."""
    return None


def reset_default_graph():
    """No default graph exists; the model-level state lives in tensorrec.session_management."""
    return None


class nn(object):
    tanh = staticmethod(lambda x, name=None: torch.tanh(_t(x)))
    relu = staticmethod(lambda x, name=None: torch.relu(_t(x)))
    sigmoid = staticmethod(lambda x, name=None: torch.sigmoid(_t(x)))

    @staticmethod
    def l2_normalize(x, axis=None, epsilon=1e-12, name=None, dim=None):
        x = _t(x)
        ax = axis if axis is not None else dim
        return x * torch.rsqrt(torch.clamp(torch.sum(x * x, dim=ax, keepdim=True), min=epsilon))
