"""Process-wide execution context.

The reference keeps one global tf.Session (tensorrec/session_management.py:3-21); tests reset it with
set_session(None).  Here the "session" is the CUDA device + stream the kernels are launched on, plus the variable
scope (name -> parameter tensor) that plugin graphs create their weights in -- the define-by-run analogue of the
TF graph's variable collection."""
import contextlib
import threading

import torch


class Session(object):
    def __init__(self, device=None):
        if device is None:
            device = torch.device('cuda', torch.cuda.current_device()) if torch.cuda.is_available() \
                else torch.device('cpu')
        self.device = torch.device(device)

    @property
    def is_cuda(self):
        return self.device.type == 'cuda'

    def run(self, fetches, feed_dict=None):
        """tf.Session.run for code written against the reference: nothing is deferred here, so fetching a tensor just
        returns its value as a numpy array (lists / tuples element-wise; ops such as an initializer are None)."""
        if isinstance(fetches, (list, tuple)):
            return type(fetches)(self.run(f) for f in fetches)
        if isinstance(fetches, torch.Tensor):
            return fetches.detach().cpu().numpy()
        return fetches


_session = None


def get_session():
    global _session
    if _session is None:
        _session = Session()
    return _session


def set_session(session):
    global _session
    _session = session


# ---- variable scope ------------------------------------------------------------------------------------
_scope = threading.local()


@contextlib.contextmanager
def variable_scope(store):
    """Makes `store` (an ordered dict name -> tensor) the target of get_variable() inside the block."""
    previous = getattr(_scope, 'store', None)
    _scope.store = store
    try:
        yield store
    finally:
        _scope.store = previous


@contextlib.contextmanager
def name_scope(label):
    """Names the plugin call that is about to run ('item', 'user_0', 'attn_1', 'loss', ...).  Variables a plugin creates
    WITHOUT a name (tf.Variable(initial_value) in code written against the reference) are registered as
    'Variable_<label>_<n>', n counting the anonymous creations inside this call: the same call creates the same names
    on every training step and at predict time, whatever ran before it (the reference gets this from building its graph
    once; define-by-run code re-runs the plugin methods every step)."""
    previous = getattr(_scope, 'label', None), getattr(_scope, 'anonymous', 0)
    _scope.label, _scope.anonymous = str(label), 0
    try:
        yield
    finally:
        _scope.label, _scope.anonymous = previous


def next_anonymous_name():
    _scope.anonymous = getattr(_scope, 'anonymous', 0) + 1
    return 'Variable_{}_{}'.format(getattr(_scope, 'label', None) or 'anonymous', _scope.anonymous)


@contextlib.contextmanager
def training_step():
    """Inside a training step every graph function evaluates with differentiable torch ops, also when no input
    happens to carry a gradient (weight-free representation graphs): the kernel path is the predict path."""
    previous = getattr(_scope, 'training', False)
    _scope.training = True
    try:
        yield
    finally:
        _scope.training = previous


def in_training_step():
    return getattr(_scope, 'training', False)


def get_variable(name, initializer):
    """Returns the trainable tensor registered under `name`, creating it with initializer() on first use.

    Plugin graphs must create their weights through this function: `connect_*` methods run on every training step
    (define-by-run), and a weight created any other way would be re-initialised each step."""
    store = getattr(_scope, 'store', None)
    if store is None:
        value = initializer()
        return value.requires_grad_(True) if value.is_floating_point() else value
    if name not in store:
        value = initializer().detach().clone()
        store[name] = value.requires_grad_(True)
    return store[name]
