"""Representation plugin graphs: features -> latent representation (API of tensorrec/representation_graphs.py).

Contract (same method name, arguments and return value as the reference, representation_graphs.py:9-23):
    connect_representation_graph(tf_features, n_components, n_features, node_name_ending) -> (repr, [weights])
`tf_features` is a torch sparse tensor [n_rows, n_features] (the stand-in for tf.SparseTensor); the method runs on
every training step (define-by-run), so weights MUST be created through session_management.get_variable(name, init).

`b200_kind` tells TensorRec which hand-written kernel evaluates the graph on the predict / predict_rank hot path:
'linear' and 'normalized_linear' lower to the CSR gather-reduce kernel K1 (trk_csr_gather_reduce_f32).  Graphs
without a kind are evaluated by running this method under torch.no_grad() on the CUDA device and handing the dense
result to the score / rank kernels."""
import torch

from .session_management import get_variable


def _l2_normalize(x, eps=1e-12):
    """tf.nn.l2_normalize(x, 1): x * rsqrt(max(sum(x^2), eps))."""
    return x * torch.rsqrt(torch.clamp(torch.sum(x * x, dim=1, keepdim=True), min=eps))


def _random_normal(shape, stddev, like):
    return torch.randn(*shape, device=like.device, dtype=torch.float32) * stddev


def _sparse_dense_matmul(tf_features, weights):
    """tf.sparse_tensor_dense_matmul (differentiable w.r.t. the dense operand): K1 forward, K1 on the transposed CSR
    backward when the features live on the CUDA device (sparse_ops.py)."""
    from .sparse_ops import sparse_dense_matmul
    return sparse_dense_matmul(tf_features, weights)


class AbstractRepresentationGraph(object):
    """Declared like the reference (py2-style metaclass there): instantiating the abstract class does not raise."""
    b200_kind = None

    def connect_representation_graph(self, tf_features, n_components, n_features, node_name_ending):
        pass


class LinearRepresentationGraph(AbstractRepresentationGraph):
    """repr = features @ W[n_features, n_components]; W initialised as L2-normalised normal rows
    (representation_graphs.py:32-43)."""
    b200_kind = 'linear'

    @staticmethod
    def weight_name(node_name_ending):
        return 'linear_weights_{}'.format(node_name_ending)

    def connect_representation_graph(self, tf_features, n_components, n_features, node_name_ending):
        weights = get_variable(self.weight_name(node_name_ending),
                               lambda: _l2_normalize(_random_normal([n_features, n_components], 1.0, tf_features)))
        return _sparse_dense_matmul(tf_features, weights), [weights]


class NormalizedLinearRepresentationGraph(LinearRepresentationGraph):
    """Linear representation followed by a row L2-normalisation (representation_graphs.py:53-58)."""
    b200_kind = 'normalized_linear'

    def connect_representation_graph(self, tf_features, n_components, n_features, node_name_ending):
        tf_repr, weights = super(NormalizedLinearRepresentationGraph, self).connect_representation_graph(
            tf_features=tf_features, n_components=n_components, n_features=n_features,
            node_name_ending=node_name_ending)
        return _l2_normalize(tf_repr), weights


class FeaturePassThroughRepresentationGraph(AbstractRepresentationGraph):
    """The features are the representation (representation_graphs.py:61-74)."""

    def connect_representation_graph(self, tf_features, n_components, n_features, node_name_ending):
        if n_components != n_features:
            raise ValueError('{} requires n_features and n_components to be equal. Either adjust n_components or use a '
                             'different representation graph. n_features = {}, n_components = {}'.format(
                                 self.__class__.__name__, n_features, n_components))
        return tf_features.to_dense(), []


class WeightedFeaturePassThroughRepresentationGraph(FeaturePassThroughRepresentationGraph):
    """Pass-through with one weight per feature (representation_graphs.py:77-89)."""

    def connect_representation_graph(self, tf_features, n_components, n_features, node_name_ending):
        dense_repr, _ = super(WeightedFeaturePassThroughRepresentationGraph, self).connect_representation_graph(
            tf_features=tf_features, n_components=n_components, n_features=n_features,
            node_name_ending=node_name_ending)
        # the reference multiplies by tf.ones (a constant, not a tf.Variable): it is returned for regularisation
        # but never trained; same here
        weights = torch.ones([1, n_components], device=dense_repr.device)
        return dense_repr * weights, [weights]


class ReLURepresentationGraph(AbstractRepresentationGraph):
    """One ReLU hidden layer of size relu_size (default 4 * n_components) (representation_graphs.py:92-124)."""

    def __init__(self, relu_size=None):
        self.relu_size = relu_size

    def connect_representation_graph(self, tf_features, n_components, n_features, node_name_ending):
        relu_size = 4 * n_components if self.relu_size is None else self.relu_size
        relu_weights = get_variable('relu_weights_{}'.format(node_name_ending),
                                    lambda: _random_normal([n_features, relu_size], .5, tf_features))
        relu_biases = get_variable('relu_biases_{}'.format(node_name_ending),
                                   lambda: torch.zeros([1, relu_size], device=tf_features.device))
        linear_weights = get_variable('linear_weights_{}'.format(node_name_ending),
                                      lambda: _random_normal([relu_size, n_components], .5, tf_features))
        hidden = torch.relu(_sparse_dense_matmul(tf_features, relu_weights) + relu_biases)
        return hidden @ linear_weights, [relu_weights, linear_weights, relu_biases]


class AbstractKerasRepresentationGraph(AbstractRepresentationGraph):
    """The reference drives Keras layers here (representation_graphs.py:127-159).  Keras is a TensorFlow front end;
    in this build `create_layers` returns torch.nn.Module layers, applied in order to the dense features."""

    def connect_representation_graph(self, tf_features, n_components, n_features, node_name_ending):
        key = 'keras_layers_{}'.format(node_name_ending)
        if not hasattr(self, '_layers'):
            self._layers = {}
        if key not in self._layers:
            layers = self.create_layers(n_features=n_features, n_components=n_components)
            self._layers[key] = [layer.to(tf_features.device) if hasattr(layer, 'to') else layer for layer in layers]
        last_layer = tf_features.to_dense()
        weights = []
        for layer in self._layers[key]:
            last_layer = layer(last_layer)
            if hasattr(layer, 'parameters'):
                weights.extend(list(layer.parameters()))
        return last_layer, weights

    def create_layers(self, n_features, n_components):
        pass
