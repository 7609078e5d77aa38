"""The TensorRec model class: same constructor, fit / fit_partial / predict / predict_rank / predict_* / save / load
signatures and error behaviour as tensorrec/tensorrec.py, with the predict / predict_rank hot path evaluated by
hand-written sm_100a kernels through the C ABI (include/tensorrec_b200.h):

    sparse features --K1 trk_csr_gather_reduce_f32--> representations (fp32 and/or split-fp16 operand)
                    --trk_csr_project_biases_f32--> user / item biases
    predict():       K2 trk_score_dense_f16x3 (tcgen05) or trk_score_f32 (exact fp32, any shape)  -> [U, I] float32
    predict_rank():  ... + K3 trk_rank_full                                                      -> [U, I] int32
    predict_rank(k): K2+K3 fused trk_score_topk_f16x3 + trk_topk_merge (+ one NCCL all-gather when the item axis is
                     sharded over GPUs)                                                           -> top-k ids, scores

Training (fit) is outside that path: it is a torch-autograd step over the plugin graphs' differentiable forms
(SURVEY.md 8f rank 1)."""
import collections
import logging
import os
import pickle
from itertools import cycle

import numpy as np
import scipy.sparse as sp
import torch

from . import kernels
from .errors import (
    ModelNotBiasedException, ModelNotFitException, ModelWithoutAttentionException, BatchNonSparseInputException
)
from .input_utils import SparseInput, TensorRecDataset
from .loss_graphs import AbstractLossGraph, RMSELossGraph
from .prediction_graphs import (
    AbstractPredictionGraph, DotProductPredictionGraph, CosineSimilarityPredictionGraph,
    EuclideanSimilarityPredictionGraph
)
from .recommendation_graphs import (
    split_sparse_tensor_indices, bias_prediction_dense, bias_prediction_serial, densify_sampled_item_predictions,
    collapse_mixture_of_tastes
)
from .representation_graphs import (
    AbstractRepresentationGraph, LinearRepresentationGraph, NormalizedLinearRepresentationGraph
)
from .session_management import get_session, variable_scope, get_variable, name_scope
from .util import sample_items, calculate_batched_alpha

TopK = collections.namedtuple('TopK', ['items', 'scores'])
TopK.__doc__ = """predict_rank(k=...) result: items int32 [n_users, k] = the item ids holding reference ranks 1..k (in
rank order), scores float32 [n_users, k].  Slots beyond n_items hold id 2**31-1 / score -inf."""

_BUILTIN_REPR = (LinearRepresentationGraph, NormalizedLinearRepresentationGraph)
_BUILTIN_PRED = (DotProductPredictionGraph, CosineSimilarityPredictionGraph, EuclideanSimilarityPredictionGraph)

# 'auto': tensor cores whenever the shape allows; 'exact': always the fp32 CUDA-core kernel; 'tensor': insist.
SCORE_PATH = os.environ.get('TENSORREC_B200_SCORE_PATH', 'auto')
# predict_rank(k) on the tensor path: 'auto' = 1-pass filter + exact fp32 re-scoring when k allows, 'exact' = always the
# 3-pass split-product kernel
TOPK_PATH = os.environ.get('TENSORREC_B200_TOPK_PATH', 'auto')


def _names_argument(method, name):
    """Does `method` declare `name` as an explicit parameter (not just **kwargs)?"""
    import inspect
    try:
        return name in inspect.signature(method).parameters
    except (TypeError, ValueError):
        return True


class _Hook(object):
    """Placeholder stored in the tf_* attributes once the model is built (the reference stores TF nodes there and
    tests `self.tf_prediction is None` to detect an unfitted model, tensorrec.py:654)."""

    def __init__(self, name):
        self.name = name

    def __repr__(self):
        return '<graph hook {}>'.format(self.name)


class TensorRec(object):

    def __init__(self,
                 n_components=100,
                 n_tastes=1,
                 user_repr_graph=LinearRepresentationGraph(),
                 item_repr_graph=LinearRepresentationGraph(),
                 attention_graph=None,
                 prediction_graph=DotProductPredictionGraph(),
                 loss_graph=RMSELossGraph(),
                 biased=True,):
        """A TensorRec recommendation model (arguments as tensorrec/tensorrec.py:28-61)."""
        # Arg Check (tensorrec.py:69-88)
        if (n_components is None) or (n_tastes is None) or (user_repr_graph is None) or (item_repr_graph is None) \
                or (prediction_graph is None) or (loss_graph is None):
            raise ValueError("All arguments to TensorRec() must be non-None")
        if n_components < 1:
            raise ValueError("n_components must be >= 1")
        if n_tastes < 1:
            raise ValueError("n_tastes must be >= 1")
        if not isinstance(user_repr_graph, AbstractRepresentationGraph):
            raise ValueError("user_repr_graph must inherit AbstractRepresentationGraph")
        if not isinstance(item_repr_graph, AbstractRepresentationGraph):
            raise ValueError("item_repr_graph must inherit AbstractRepresentationGraph")
        if not isinstance(prediction_graph, AbstractPredictionGraph):
            raise ValueError("prediction_graph must inherit AbstractPredictionGraph")
        if not isinstance(loss_graph, AbstractLossGraph):
            raise ValueError("loss_graph must inherit AbstractLossGraph")
        if attention_graph is not None:
            if not isinstance(attention_graph, AbstractRepresentationGraph):
                raise ValueError("attention_graph must be None or inherit AbstractRepresentationGraph")
            if n_tastes == 1:
                raise ValueError("attention_graph must be None if n_tastes == 1")

        self.n_components = n_components
        self.n_tastes = n_tastes
        self.user_repr_graph_factory = user_repr_graph
        self.item_repr_graph_factory = item_repr_graph
        self.attention_graph_factory = attention_graph
        self.prediction_graph_factory = prediction_graph
        self.loss_graph_factory = loss_graph
        self.biased = biased

        # graph hook attribute names, as the reference declares them (tensorrec.py:100-124)
        self.graph_tensor_hook_attr_names = [
            'tf_user_representation', 'tf_item_representation', 'tf_prediction_serial', 'tf_prediction', 'tf_rankings',
            'tf_predict_similar_items', 'tf_rank_similar_items',
            'tf_basic_loss', 'tf_weight_reg_loss', 'tf_loss',
            'tf_learning_rate', 'tf_alpha', 'tf_sample_indices', 'tf_n_sampled_items', 'tf_similar_items_ids',
        ]
        if self.biased:
            self.graph_tensor_hook_attr_names += ['tf_projected_user_biases', 'tf_projected_item_biases']
        if self.attention_graph_factory is not None:
            self.graph_tensor_hook_attr_names += ['tf_user_attention_representation']
        self.graph_operation_hook_attr_names = ['tf_optimizer']
        self.graph_iterator_hook_attr_names = ['tf_user_feature_iterator', 'tf_item_feature_iterator',
                                               'tf_interaction_iterator']
        self._break_graph_hooks()

        self.n_user_features = None
        self.n_item_features = None
        self._variables = collections.OrderedDict()   # name -> trainable tensor (the model's weights)
        self._optimizer = None
        self._optimizer_params = None
        self._stepped = False          # has any training step completed?

    # ------------------------------------------------------------------------------------------------
    # graph hooks (tensorrec.py:136-183).  Only their None-ness carries meaning here.
    # ------------------------------------------------------------------------------------------------
    def _all_hook_names(self):
        return (self.graph_tensor_hook_attr_names + self.graph_operation_hook_attr_names +
                self.graph_iterator_hook_attr_names)

    def _break_graph_hooks(self):
        for name in self._all_hook_names():
            self.__setattr__(name, None)

    def _attach_graph_hooks(self):
        for name in self._all_hook_names():
            self.__setattr__(name, _Hook(name))

    # ------------------------------------------------------------------------------------------------
    # input handling (tensorrec.py:185-263, util.py:34-58)
    # ------------------------------------------------------------------------------------------------
    @staticmethod
    def _inputs_from_raw(raw_input):
        """scipy sparse matrix / TensorRecDataset / list of those -> list of SparseInput."""
        def ok(v):
            return sp.issparse(v) or isinstance(v, (TensorRecDataset, SparseInput))

        def wrap(v):
            return v if isinstance(v, SparseInput) else SparseInput(v)

        if ok(raw_input):
            return [wrap(raw_input)]
        if isinstance(raw_input, list) and len(raw_input) > 0 and all(ok(v) for v in raw_input):
            return [wrap(v) for v in raw_input]
        if isinstance(raw_input, str) or (isinstance(raw_input, list) and raw_input and
                                          all(isinstance(v, str) for v in raw_input)):
            # TFRecord path(s) (tensorrec.py:203-215): every record of every file is one batch
            from .input_utils import create_tensorrec_dataset_from_tfrecord
            paths = [raw_input] if isinstance(raw_input, str) else raw_input
            return [wrap(ds) for path in paths for ds in create_tensorrec_dataset_from_tfrecord(path)]
        raise ValueError('Input must be a scipy sparse matrix, an iterable of scipy sprase matrices, or a TensorFlow '
                         'Dataset')

    @classmethod
    def _single_input(cls, raw_input, what):
        inputs = cls._inputs_from_raw(raw_input)
        if len(inputs) != 1:
            raise ValueError('{} must be one matrix at predict time (got a list of {})'.format(what, len(inputs)))
        return inputs[0]

    def _create_batched_inputs(self, interactions, user_features, item_features, user_batch_size=None):
        if user_batch_size is not None:
            if (not sp.issparse(interactions)) or (not sp.issparse(user_features)):
                raise BatchNonSparseInputException()
            if not isinstance(interactions, sp.csr_matrix):
                interactions = sp.csr_matrix(interactions)
            if not isinstance(user_features, sp.csr_matrix):
                user_features = sp.csr_matrix(user_features)
            n_users = user_features.shape[0]
            interactions_batched, user_features_batched = [], []
            start_batch = 0
            while start_batch < n_users:
                end_batch = min(start_batch + user_batch_size, n_users)
                interactions_batched.append(interactions[start_batch:end_batch])
                user_features_batched.append(user_features[start_batch:end_batch])
                start_batch = end_batch
            interactions, user_features = interactions_batched, user_features_batched

        int_in = self._inputs_from_raw(interactions)
        uf_in = self._inputs_from_raw(user_features)
        if_in = self._inputs_from_raw(item_features)
        if len(int_in) != len(uf_in):
            raise ValueError('Number of batches in user_features and interactions must be equal.')
        if (len(if_in) > 1) and (len(if_in) != len(uf_in)):
            raise ValueError('Number of batches in item_features must be 1 or equal to the number of batches in '
                             'user_features.')
        return [batch for batch in zip(int_in, uf_in, cycle(if_in))]

    # ------------------------------------------------------------------------------------------------
    # training step (define-by-run form of _build_tf_graph, tensorrec.py:270-492)
    # ------------------------------------------------------------------------------------------------
    def _training_losses(self, interactions, user_features, item_features, n_sampled_items, device):
        from .session_management import training_step
        with training_step():
            return self._training_losses_impl(interactions, user_features, item_features, n_sampled_items, device)

    def _training_losses_impl(self, interactions, user_features, item_features, n_sampled_items, device):
        tf_user_features = user_features.torch_sparse(device)
        tf_item_features = item_features.torch_sparse(device)
        tf_interactions = interactions.torch_sparse(device)
        n_users, n_items = user_features.shape[0], item_features.shape[0]
        loss_graph = self.loss_graph_factory
        tf_weights = []

        with name_scope('item'):
            item_repr, item_weights = self.item_repr_graph_factory.connect_representation_graph(
                tf_features=tf_item_features, n_components=self.n_components, n_features=self.n_item_features,
                node_name_ending='item')
        tf_weights.extend(item_weights)

        tf_x_user, tf_x_item = split_sparse_tensor_indices(tf_sparse_tensor=tf_interactions, n_dimensions=2)
        if loss_graph.is_sample_based:
            sample_indices = torch.from_numpy(sample_items(n_items, n_users, n_sampled_items,
                                                           replace=loss_graph.is_sampled_with_replacement)).to(device)
            tf_x_user_sample, tf_x_item_sample = sample_indices[:, 0], sample_indices[:, 1]

        pred_graph = self.prediction_graph_factory
        tastes_predictions, tastes_prediction_serials, tastes_sample_prediction_serials = [], [], []
        with_attention = self.attention_graph_factory is not None
        tastes_attentions = [] if with_attention else None
        tastes_attention_serials = [] if with_attention else None
        tastes_sample_attention_serials = [] if with_attention else None

        for taste in range(self.n_tastes):
            with name_scope('user_{}'.format(taste)):
                user_repr, user_weights = self.user_repr_graph_factory.connect_representation_graph(
                    tf_features=tf_user_features, n_components=self.n_components, n_features=self.n_user_features,
                    node_name_ending='user_{}'.format(taste))
            tf_weights.extend(user_weights)

            if with_attention:
                with name_scope('attn_{}'.format(taste)):
                    attention_repr, attention_weights = self.attention_graph_factory.connect_representation_graph(
                        tf_features=tf_user_features, n_components=self.n_components,
                        n_features=self.n_user_features, node_name_ending='attn_{}'.format(taste))
                tf_weights.extend(attention_weights)
                if loss_graph.is_dense:
                    tastes_attentions.append(pred_graph.connect_dense_prediction_graph(
                        tf_user_representation=attention_repr, tf_item_representation=item_repr))
                tastes_attention_serials.append(pred_graph.connect_serial_prediction_graph(
                    tf_user_representation=attention_repr, tf_item_representation=item_repr,
                    tf_x_user=tf_x_user, tf_x_item=tf_x_item))
                if loss_graph.is_sample_based:
                    # the reference feeds the USER representation here, not the attention one (tensorrec.py:367-372)
                    tastes_sample_attention_serials.append(pred_graph.connect_serial_prediction_graph(
                        tf_user_representation=user_repr, tf_item_representation=item_repr,
                        tf_x_user=tf_x_user_sample, tf_x_item=tf_x_item_sample))

            if loss_graph.is_dense:
                tastes_predictions.append(pred_graph.connect_dense_prediction_graph(
                    tf_user_representation=user_repr, tf_item_representation=item_repr))
            tastes_prediction_serials.append(pred_graph.connect_serial_prediction_graph(
                tf_user_representation=user_repr, tf_item_representation=item_repr,
                tf_x_user=tf_x_user, tf_x_item=tf_x_item))
            if loss_graph.is_sample_based:
                tastes_sample_prediction_serials.append(pred_graph.connect_serial_prediction_graph(
                    tf_user_representation=user_repr, tf_item_representation=item_repr,
                    tf_x_user=tf_x_user_sample, tf_x_item=tf_x_item_sample))

        tf_prediction = None
        if loss_graph.is_dense:
            tf_prediction = collapse_mixture_of_tastes(tastes_predictions, tastes_attentions if with_attention else None)
        tf_prediction_serial = collapse_mixture_of_tastes(tastes_prediction_serials, tastes_attention_serials)
        tf_sample_predictions_serial = None
        if loss_graph.is_sample_based:
            tf_sample_predictions_serial = collapse_mixture_of_tastes(tastes_sample_prediction_serials,
                                                                      tastes_sample_attention_serials)

        if self.biased:
            user_feature_biases = get_variable('feature_biases_user',
                                               lambda: torch.zeros([self.n_user_features, 1], device=device))
            item_feature_biases = get_variable('feature_biases_item',
                                               lambda: torch.zeros([self.n_item_features, 1], device=device))
            projected_user_biases = torch.sum(torch.sparse.mm(tf_user_features, user_feature_biases), dim=1)
            projected_item_biases = torch.sum(torch.sparse.mm(tf_item_features, item_feature_biases), dim=1)
            tf_weights.append(user_feature_biases)
            tf_weights.append(item_feature_biases)
            if tf_prediction is not None:
                tf_prediction = bias_prediction_dense(tf_prediction, projected_user_biases, projected_item_biases)
            tf_prediction_serial = bias_prediction_serial(tf_prediction_serial, projected_user_biases,
                                                          projected_item_biases, tf_x_user, tf_x_item)
            if tf_sample_predictions_serial is not None:
                tf_sample_predictions_serial = bias_prediction_serial(
                    tf_sample_predictions_serial, projected_user_biases, projected_item_biases,
                    tf_x_user_sample, tf_x_item_sample)

        # loss-graph kwargs with the reference's visibility rules (tensorrec.py:463-482)
        loss_graph_kwargs = {
            'tf_prediction_serial': tf_prediction_serial,
            'tf_interactions_serial': tf_interactions._values(),
            'tf_interactions': tf_interactions,
            'tf_n_users': n_users,
            'tf_n_items': n_items,
        }
        if loss_graph.is_dense:
            # In the reference tf_rankings is a graph node that costs nothing unless the loss graph consumes it; here it
            # is a full K3 sort of [n_users, n_items] per step, so it is evaluated only for a loss graph that names it
            tf_rankings = None
            if tf_prediction.is_cuda and _names_argument(loss_graph.connect_loss_graph, 'tf_rankings'):
                # ranks come from the K3 kernel; they carry no gradient (as tf.nn.top_k)
                tf_rankings = kernels.rank_full(tf_prediction.detach().contiguous())
            loss_graph_kwargs.update({'tf_prediction': tf_prediction, 'tf_rankings': tf_rankings})
        if loss_graph.is_sample_based:
            loss_graph_kwargs.update({
                'tf_sample_predictions': densify_sampled_item_predictions(
                    tf_sample_predictions_serial=tf_sample_predictions_serial,
                    tf_n_sampled_items=n_sampled_items, tf_n_users=n_users),
                'tf_n_sampled_items': n_sampled_items})

        with name_scope('loss'):
            basic_loss = loss_graph.connect_loss_graph(**loss_graph_kwargs)
        weight_reg_loss = sum(0.5 * torch.sum(w * w) for w in tf_weights)      # sum of tf.nn.l2_loss
        return basic_loss, weight_reg_loss, tf_prediction_serial, tf_weights

    def fit(self, interactions, user_features, item_features, epochs=100, learning_rate=0.1, alpha=0.00001,
            verbose=False, user_batch_size=None, n_sampled_items=None):
        """Constructs the model on first use and fits it (arguments as tensorrec/tensorrec.py:494-526)."""
        self.fit_partial(interactions=interactions, user_features=user_features, item_features=item_features,
                         epochs=epochs, learning_rate=learning_rate, alpha=alpha, verbose=verbose,
                         user_batch_size=user_batch_size, n_sampled_items=n_sampled_items)

    def fit_partial(self, interactions, user_features, item_features, epochs=1, learning_rate=0.1,
                    alpha=0.00001, verbose=False, user_batch_size=None, n_sampled_items=None):
        """One or more epochs of Adam on the loss graph (tensorrec/tensorrec.py:539-634)."""
        device = get_session().device

        if self.loss_graph_factory.is_sample_based:
            if (n_sampled_items is None) or (n_sampled_items <= 0):
                raise ValueError("n_sampled_items must be an integer >0")
        if (n_sampled_items is not None) and (not self.loss_graph_factory.is_sample_based):
            logging.warning('n_sampled_items was specified, but the loss graph is not sample-based')

        if verbose:
            logging.info('Processing interaction and feature data')
        batches = self._create_batched_inputs(interactions=interactions, user_features=user_features,
                                              item_features=item_features, user_batch_size=user_batch_size)

        first_build = self.tf_prediction is None
        if first_build:
            # feature counts are learned from the first batch and cannot change afterwards (tensorrec.py:598-605)
            self.n_user_features = batches[0][1].shape[1]
            self.n_item_features = batches[0][2].shape[1]
            self._attach_graph_hooks()

        batched_alpha = calculate_batched_alpha(num_batches=len(batches), alpha=alpha)
        if verbose:
            logging.info('Beginning fitting')
        try:
            self._fit_epochs(batches, epochs, learning_rate, alpha, batched_alpha, verbose, n_sampled_items, device)
            first_build = False
        finally:
            if first_build and not getattr(self, '_stepped', False):
                # the very first step failed (e.g. n_sampled_items > n_items): the model is still unbuilt -- predict()
                # must keep raising ModelNotFitException and a later fit may use other feature counts
                self._break_graph_hooks()
                self.n_user_features = self.n_item_features = None
                self._variables.clear()
                self._optimizer = None

    def _fit_epochs(self, batches, epochs, learning_rate, alpha, batched_alpha, verbose, n_sampled_items, device):
        from . import train_kernels
        on_kernels = device.type == 'cuda' and train_kernels.eligible(self) and \
            (n_sampled_items is None or n_sampled_items <= 2048)
        for epoch in range(epochs):
            for batch, (int_in, uf_in, if_in) in enumerate(batches):
                if uf_in.shape[1] != self.n_user_features or if_in.shape[1] != self.n_item_features:
                    raise ValueError('feature matrices have {} / {} columns but the model was built for {} / {}'.format(
                        uf_in.shape[1], if_in.shape[1], self.n_user_features, self.n_item_features))
                if on_kernels:
                    # the sampled-rank step on hand-written kernels (train_kernels.py)
                    if getattr(self, '_wmrb_step', None) is None or self._wmrb_step.device != device:
                        self._wmrb_step = train_kernels.WmrbStep(self, device)
                    n_pos = int_in.n_positive
                    loss_vec, serial_predictions = self._wmrb_step.step(
                        int_in, uf_in, if_in, n_sampled_items, learning_rate, l2=n_pos * batched_alpha)
                    self._stepped = True
                    if verbose:
                        mean_loss = float(loss_vec.sum()) / max(n_pos, 1)
                        mean_pred = float(torch.mean(serial_predictions))
                        wr = sum(0.5 * float(torch.sum(w.detach() * w.detach())) for w in self._variables.values())
                        logging.info('EPOCH {} BATCH {} loss = {}, weight_reg_l2_loss = {}, mean_pred = {}'.format(
                            epoch, batch, mean_loss, alpha * wr, mean_pred))
                    continue
                with variable_scope(self._variables):
                    basic_loss, wr_loss, serial_predictions, tf_weights = self._training_losses(
                        int_in, uf_in, if_in, n_sampled_items, device)
                loss = basic_loss + batched_alpha * wr_loss
                params = [w for w in tf_weights if w.requires_grad and w.is_leaf]
                self._ensure_optimizer(params, learning_rate)
                self._optimizer.zero_grad(set_to_none=True)
                loss.sum().backward()       # tf.gradients of a vector loss (WMRB) is the gradient of its sum
                self._optimizer.step()
                self._stepped = True
                if verbose:
                    mean_loss = float(torch.mean(basic_loss.detach()))
                    mean_pred = float(torch.mean(serial_predictions.detach()))
                    weight_reg_l2_loss = alpha * float(wr_loss.detach())
                    logging.info('EPOCH {} BATCH {} loss = {}, weight_reg_l2_loss = {}, mean_pred = {}'.format(
                        epoch, batch, mean_loss, weight_reg_l2_loss, mean_pred))

    def _ensure_optimizer(self, params, learning_rate):
        ids = tuple(id(p) for p in params)
        if self._optimizer is None or self._optimizer_params != ids:
            self._optimizer = torch.optim.Adam(params, lr=learning_rate)     # tf.train.AdamOptimizer defaults
            self._optimizer_params = ids
        for group in self._optimizer.param_groups:
            group['lr'] = learning_rate

    # ------------------------------------------------------------------------------------------------
    # weights
    # ------------------------------------------------------------------------------------------------
    def get_weights(self):
        """name -> numpy array of every model weight (linear_weights_item, linear_weights_user_<t>,
        linear_weights_attn_<t>, feature_biases_user, feature_biases_item, ...)."""
        return collections.OrderedDict((k, v.detach().cpu().numpy().copy()) for k, v in self._variables.items())

    def set_weights(self, weights, n_user_features=None, n_item_features=None):
        """Injects weights (dict name -> array).  Marks the model as built, so predict* work without fit -- the
        parity-test hook: the reference never seeds its initialiser (representation_graphs.py:35), so values after
        fit() are unpinned."""
        device = get_session().device
        for name, value in weights.items():
            t = torch.as_tensor(np.asarray(value, dtype=np.float32)).to(device).clone()
            self._variables[name] = t.requires_grad_(True)
        self._optimizer = None
        self._wmrb_step = None          # Adam moments of the kernel training path belong to the replaced weights
        if n_user_features is None and 'linear_weights_user_0' in self._variables:
            n_user_features = self._variables['linear_weights_user_0'].shape[0]
        if n_item_features is None and 'linear_weights_item' in self._variables:
            n_item_features = self._variables['linear_weights_item'].shape[0]
        self.n_user_features = n_user_features if n_user_features is not None else self.n_user_features
        self.n_item_features = n_item_features if n_item_features is not None else self.n_item_features
        self._attach_graph_hooks()

    def _var(self, name, device):
        if name not in self._variables:
            raise RuntimeError('weight {!r} does not exist; fit the model or inject it with set_weights()'.format(name))
        v = self._variables[name]
        if v.device != device:     # fitted on another device: move once and keep
            v = v.detach().to(device).requires_grad_(True)
            self._variables[name] = v
            self._optimizer = None
        return v.detach()

    # ------------------------------------------------------------------------------------------------
    # the predict / predict_rank hot path
    # ------------------------------------------------------------------------------------------------
    @staticmethod
    def _cuda_device():
        kernels.require_cuda()
        return torch.device('cuda', torch.cuda.current_device())

    def _check_features(self, sparse_input, n_features, side):
        if n_features is not None and sparse_input.shape[1] != n_features:
            raise ValueError('{} feature matrix has {} columns but the model was built for {}'.format(
                side, sparse_input.shape[1], n_features))

    def _represent(self, graph, sparse_input, n_features, node_name_ending, device, extra_normalize=0,
                   want_f32=True, split_d_pad=None, want_norm=False, stats=None):
        """One representation on the device: (repr_f32 | None, split | None, scale | None[, norm])."""
        if type(graph) in _BUILTIN_REPR:
            weights = self._var(LinearRepresentationGraph.weight_name(node_name_ending), device)
            n_norm = (1 if graph.b200_kind == 'normalized_linear' else 0) + extra_normalize
            return kernels.gather_reduce(sparse_input.device_csr(device), weights, n_normalize=n_norm,
                                         want_f32=want_f32, split_d_pad=split_d_pad, want_norm=want_norm, stats=stats)
        # user-defined / non-linear plugin: run its own forward on the device, then hand the dense rows to the kernels
        with torch.no_grad(), variable_scope(self._variables), name_scope(node_name_ending):
            for k in list(self._variables):
                self._var(k, device)
            dense, _ = graph.connect_representation_graph(
                tf_features=sparse_input.torch_sparse(device), n_components=self.n_components, n_features=n_features,
                node_name_ending=node_name_ending)
        dense = dense.detach().to(torch.float32).contiguous().clone()
        for _ in range(extra_normalize):
            kernels.l2_normalize_rows_(dense)
        split = scale = None
        if split_d_pad is not None:
            split, scale = kernels.split_f32(dense, n_normalize=0, d_pad=split_d_pad)
        if stats is not None:
            stats.zero_()
        if want_norm or stats is not None:
            norm = kernels.operand_stats(split, scale, split_d_pad, want_norm=want_norm, stats=stats)
            if want_norm:
                return (dense if want_f32 else None), split, scale, norm
        return (dense if want_f32 else None), split, scale

    def _projected_biases(self, sparse_input, name, device):
        return kernels.project_biases(sparse_input.device_csr(device), self._var(name, device).reshape(-1))

    def _tensor_path_ok(self, allow_tastes=False):
        """Can the tcgen05 kernels evaluate this model?  allow_tastes: the fused top-k also covers n_tastes > 1 without
        attention (one sweep per taste, then a de-duplicating merge: the prediction is the maximum over the tastes)."""
        if SCORE_PATH == 'exact':
            return False
        ok = (type(self.prediction_graph_factory) in (DotProductPredictionGraph, CosineSimilarityPredictionGraph)
              and (self.n_tastes == 1 or allow_tastes) and self.attention_graph_factory is None
              and kernels.d_pad_for(self.n_components) <= 128)
        if SCORE_PATH == 'tensor' and not ok:
            raise RuntimeError('TENSORREC_B200_SCORE_PATH=tensor but this model cannot use the tcgen05 kernel')
        return ok

    def _side_operands(self, side, sparse_in, device, for_filter=False, taste=0):
        """One side ('user' or 'item') as kernels.SideOperands: split-fp16 operand + scale, projected biases and -- for
        the filter form of the fused top-k -- the row norms (users) / the global statistics (items), all from K1."""
        extra = 1 if type(self.prediction_graph_factory) is CosineSimilarityPredictionGraph else 0
        d_pad = kernels.d_pad_for(self.n_components)
        is_user = side == 'user'
        graph = self.user_repr_graph_factory if is_user else self.item_repr_graph_factory
        n_features = self.n_user_features if is_user else self.n_item_features
        stats = norm = None
        if for_filter and not is_user:
            stats = torch.empty((3,), dtype=torch.float32, device=device)
        out = self._represent(graph, sparse_in, n_features, 'user_{}'.format(taste) if is_user else 'item', device, extra,
                              want_f32=False, split_d_pad=d_pad, want_norm=for_filter and is_user, stats=stats)
        split, scale = out[1], out[2]
        if for_filter and is_user:
            norm = out[3]
        bias = None
        if self.biased:
            bias = self._projected_biases(sparse_in, 'feature_biases_user' if is_user else 'feature_biases_item', device)
        return kernels.SideOperands(None, split, scale, bias, sparse_in.shape[0], self.n_components, d_pad, norm=norm,
                                    stats=stats)

    def _tensor_operands(self, user_in, item_in, device):
        return self._side_operands('user', user_in, device), self._side_operands('item', item_in, device)

    def _score_plan(self, item_in, device):
        """Item-side work of the dense prediction, done once per call: returns score(user_block_in, out=None) ->
        float32 [rows, n_items] on the device.  Tensor cores (split-product kernel) when the model allows, the exact
        CUDA-core kernel (tastes, attention, Euclidean, wide rows) or the plugin's own dense form otherwise."""
        n_items = item_in.shape[0]
        if self._tensor_path_ok():
            items = self._side_operands('item', item_in, device)
            meta = kernels.pack_item_meta(items.scale, items.bias, n_items)

            def score(block_in, out=None):
                users = self._side_operands('user', block_in, device)
                return kernels.score_dense_tc(users.split, users.scale, users.bias, items.split, meta,
                                              block_in.shape[0], n_items, users.d_pad, out=out)
            return score

        pred_graph = self.prediction_graph_factory
        builtin = type(pred_graph) in _BUILTIN_PRED
        extra = 1 if (builtin and pred_graph.b200_kind == 'cosine') else 0
        item_repr, _, _ = self._represent(self.item_repr_graph_factory, item_in, self.n_item_features, 'item', device,
                                          extra)
        item_bias = self._projected_biases(item_in, 'feature_biases_item', device) if self.biased else None

        def score(block_in, out=None):
            user_reprs = torch.stack([self._represent(self.user_repr_graph_factory, block_in, self.n_user_features,
                                                      'user_{}'.format(t), device, extra)[0]
                                      for t in range(self.n_tastes)])
            attention_reprs = None
            if self.attention_graph_factory is not None:
                attention_reprs = torch.stack([self._represent(self.attention_graph_factory, block_in,
                                                               self.n_user_features, 'attn_{}'.format(t), device,
                                                               extra)[0]
                                               for t in range(self.n_tastes)])
            user_bias = self._projected_biases(block_in, 'feature_biases_user', device) if self.biased else None
            if builtin and not (pred_graph.b200_kind == 'euclidean' and attention_reprs is not None):
                mode = 1 if pred_graph.b200_kind == 'euclidean' else 0
                return kernels.score_exact(user_reprs, item_repr, user_bias, item_bias, mode=mode,
                                           attention_repr=attention_reprs, out=out)
            # user-defined prediction graph: its own dense form per taste, then the reference's collapse + bias order
            with torch.no_grad():
                preds = [pred_graph.connect_dense_prediction_graph(tf_user_representation=user_reprs[t],
                                                                   tf_item_representation=item_repr)
                         for t in range(self.n_tastes)]
                atts = None
                if attention_reprs is not None:
                    atts = [pred_graph.connect_dense_prediction_graph(tf_user_representation=attention_reprs[t],
                                                                      tf_item_representation=item_repr)
                            for t in range(self.n_tastes)]
                pred = collapse_mixture_of_tastes(preds, atts)
                if self.biased:
                    pred = bias_prediction_dense(pred, user_bias, item_bias)
            pred = pred.to(torch.float32).contiguous()
            if out is not None:
                out.copy_(pred)
                return out
            return pred
        return score

    def _predict_device(self, user_in, item_in, device):
        """tf_prediction: dense float32 scores [n_users, n_items] on the device."""
        self._check_features(user_in, self.n_user_features, 'user')
        self._check_features(item_in, self.n_item_features, 'item')
        n_users, n_items = user_in.shape[0], item_in.shape[0]
        if n_users == 0 or n_items == 0:
            return torch.zeros((n_users, n_items), dtype=torch.float32, device=device)
        return self._score_plan(item_in, device)(user_in)

    # a dense [n_users, n_items] result beyond this many bytes is produced in user blocks (BASELINE config #2,
    # 1M x 100K = 400 GB, exceeds the 180 GB of HBM): two device buffers + two page-locked host buffers of this size
    PREDICT_BLOCK_BYTES = 4 << 30

    def _user_blocks(self, user_in, n_items, user_batch_size):
        n_users = user_in.shape[0]
        if user_batch_size is None:
            user_batch_size = max(128, (self.PREDICT_BLOCK_BYTES // max(4 * n_items, 1)) // 128 * 128)
        step = max(1, int(user_batch_size))
        if step >= n_users:
            return [(0, n_users, user_in)]
        csr = user_in.matrix if isinstance(user_in.matrix, sp.csr_matrix) else sp.csr_matrix(user_in.matrix)
        return [(u0, min(n_users, u0 + step), SparseInput(csr[u0:min(n_users, u0 + step)]))
                for u0 in range(0, n_users, step)]

    def predict_batches(self, user_features, item_features, user_batch_size=None):
        """predict() as a stream of user blocks: yields (u0, u1, scores float32 ndarray [u1 - u0, n_items]).

        For results that fit neither HBM nor host memory at once (BASELINE config #2: 1M x 100K = 400 GB).  The item side
        is computed once; user blocks are scored into two alternating device buffers and copied into two alternating
        page-locked host buffers on a copy stream, so the device->host transfer of block b overlaps the kernels of block
        b + 1 (the path is bound by the host link, not by HBM).  The yielded array IS the page-locked buffer: it stays
        valid until the generator is advanced twice."""
        if self.tf_prediction is None:
            raise ModelNotFitException(method='predict')
        device = self._cuda_device()
        user_in = self._single_input(user_features, 'user_features')
        item_in = self._single_input(item_features, 'item_features')
        self._check_features(user_in, self.n_user_features, 'user')
        self._check_features(item_in, self.n_item_features, 'item')
        n_users, n_items = user_in.shape[0], item_in.shape[0]
        if n_users == 0 or n_items == 0:
            yield 0, n_users, np.zeros((n_users, n_items), dtype=np.float32)
            return
        blocks = self._user_blocks(user_in, n_items, user_batch_size)
        rows = max(u1 - u0 for u0, u1, _ in blocks)
        score = self._score_plan(item_in, device)
        n_buf = min(2, len(blocks))
        # the staging buffers (page-locking gigabytes takes seconds) are kept for the next call of the same shape
        key = (rows, n_items, n_buf, str(device))
        cached = getattr(self, '_stream_buffers', None)
        if cached is None or cached[0] != key:
            self._stream_buffers = None
            cached = (key,
                      [torch.empty((rows, n_items), dtype=torch.float32, device=device) for _ in range(n_buf)],
                      [torch.empty((rows, n_items), dtype=torch.float32, pin_memory=True) for _ in range(n_buf)])
            self._stream_buffers = cached
        dev_buf, host_buf = cached[1], cached[2]
        compute = torch.cuda.current_stream()
        copier = torch.cuda.Stream(device=device)
        copied = [None] * n_buf       # event: the copy out of dev_buf[j] / into host_buf[j] has finished
        pending = None
        for b, (u0, u1, block_in) in enumerate(blocks):
            j = b % n_buf
            if copied[j] is not None:
                compute.wait_event(copied[j])          # the kernels of this block overwrite dev_buf[j]
            score(block_in, out=dev_buf[j][:u1 - u0])
            done = torch.cuda.Event()
            done.record(compute)
            with torch.cuda.stream(copier):
                copier.wait_event(done)
                host_buf[j][:u1 - u0].copy_(dev_buf[j][:u1 - u0], non_blocking=True)
                copied[j] = torch.cuda.Event()
                copied[j].record(copier)
            if pending is not None:                    # hand out block b - 1 while block b is being computed / copied
                pj, p0, p1 = pending
                copied[pj].synchronize()
                yield p0, p1, host_buf[pj][:p1 - p0].numpy()
            pending = (j, u0, u1)
        pj, p0, p1 = pending
        copied[pj].synchronize()
        yield p0, p1, host_buf[pj][:p1 - p0].numpy()

    def predict(self, user_features, item_features, out=None, user_batch_size=None):
        """Scores for every user x item pair: float32 ndarray [n_users, n_items] (tensorrec/tensorrec.py:636-664).

        Results larger than PREDICT_BLOCK_BYTES (or any result when user_batch_size / out is given) are produced in user
        blocks (predict_batches) and assembled in `out` -- a caller-provided float32 array [n_users, n_items], e.g. a
        numpy.memmap when the matrix exceeds host memory -- or in a new array."""
        if self.tf_prediction is None:
            raise ModelNotFitException(method='predict')
        device = self._cuda_device()
        user_in = self._single_input(user_features, 'user_features')
        item_in = self._single_input(item_features, 'item_features')
        n_users, n_items = user_in.shape[0], item_in.shape[0]
        streamed = (out is not None or user_batch_size is not None or
                    4 * n_users * n_items > self.PREDICT_BLOCK_BYTES)
        if not streamed:
            return kernels.to_host(self._predict_device(user_in, item_in, device))
        if out is None:
            out = np.empty((n_users, n_items), dtype=np.float32)
        elif tuple(out.shape) != (n_users, n_items) or out.dtype != np.float32:
            raise ValueError('out must be a float32 array of shape ({}, {})'.format(n_users, n_items))
        for u0, u1, block in self.predict_batches(user_in, item_in, user_batch_size=user_batch_size):
            out[u0:u1] = block
        return out

    def predict_rank(self, user_features, item_features, k=None):
        """Ranks for every user x item pair: int32 ndarray [n_users, n_items], 1 = best, ties by lower item index
        (tensorrec/tensorrec.py:705-733).  With k (an addition for shapes whose rank matrix cannot be materialised)
        only the entries with rank <= k are produced, as a TopK(items, scores) -- see predict_top_k."""
        if self.tf_prediction is None:
            raise ModelNotFitException(method='predict_rank')
        if k is not None:
            return self.predict_top_k(user_features, item_features, k)
        device = self._cuda_device()
        user_in = self._single_input(user_features, 'user_features')
        item_in = self._single_input(item_features, 'item_features')
        scores = self._predict_device(user_in, item_in, device)
        if scores.numel() == 0:
            return np.zeros(tuple(scores.shape), dtype=np.int32)
        return kernels.to_host(kernels.rank_full(scores))

    def predict_top_k(self, user_features, item_features, k, item_id_offset=0, gather_group=None, to_host=True,
                      gather='all', user_batch_size=None):
        """The k best items per user in reference rank order, without materialising the score matrix.

        Single GPU: K2+K3 fused kernel (filter form: one tensor pass + re-scoring of the survivors; users the
        certificate rejects go through the exact kernel on the device) -> TopK(items, scores) for every user.
        Item axis sharded over ranks (`gather_group` = a torch.distributed process group whose ranks each pass THEIR
        rows of item_features and the global id of the first one as item_id_offset): one all-to-all of the per-shard
        top-k -- rank r receives the candidates of ITS slice of the users from every shard and merges them -- and,
        with gather='all', one all-gather of the merged slices so that every rank returns all users.  gather='slice'
        returns this rank's users only (rows `last_topk_info['user_rows']`).
        user_batch_size: users are processed in blocks of this many rows (bounds device memory at 10M+ users)."""
        if self.tf_prediction is None:
            raise ModelNotFitException(method='predict_rank')
        device = self._cuda_device()
        user_in = self._single_input(user_features, 'user_features')
        item_in = self._single_input(item_features, 'item_features')
        self._check_features(user_in, self.n_user_features, 'user')
        self._check_features(item_in, self.n_item_features, 'item')
        n_users, n_items = user_in.shape[0], item_in.shape[0]
        k = int(k)
        if k < 1:
            raise ValueError('k must be >= 1')
        if gather not in ('all', 'slice'):
            raise ValueError("gather must be 'all' or 'slice'")
        if n_users == 0:
            return TopK(np.zeros((0, k), np.int32), np.zeros((0, k), np.float32))
        from . import distributed

        fused = (self._tensor_path_ok(allow_tastes=True) and n_items > 0 and
                 k <= kernels.topk_max_k(kernels.d_pad_for(self.n_components)))
        use_filter = fused and TOPK_PATH != 'exact' and k <= kernels.filter_max_k()
        info = self.last_topk_info = {'path': 'filter' if use_filter else ('exact3' if fused else 'dense+rank'),
                                      'fallback_rows': 0}
        items = fitems = None
        if fused:
            items = self._side_operands('item', item_in, device, for_filter=use_filter)
            if use_filter:
                fitems = kernels.FilterItems(items)

        if user_batch_size is None or user_batch_size >= n_users:
            blocks = [(0, n_users, user_in)]
        else:
            step = max(1, int(user_batch_size))
            csr = user_in.matrix if isinstance(user_in.matrix, sp.csr_matrix) else sp.csr_matrix(user_in.matrix)
            blocks = [(u0, min(n_users, u0 + step), SparseInput(csr[u0:min(n_users, u0 + step)]))
                      for u0 in range(0, n_users, step)]

        def run_taste(block_in, taste, force_exact):
            users = self._side_operands('user', block_in, device, for_filter=use_filter and not force_exact, taste=taste)
            if use_filter and not force_exact:
                return kernels.topk_filter(users, items, k, item_id_offset=item_id_offset, fitems=fitems)
            return kernels.topk_exact(users, items, k, item_id_offset=item_id_offset), None, 0

        def run_block(block_in, force_exact=False):
            """-> (PackedTopK of the block, [(device counters | None, capacity)] of its sweeps)"""
            if not fused:
                return self._topk_from_dense(block_in, item_in, k, item_id_offset, device), [(None, 0)]
            if self.n_tastes == 1:
                top, cnt, cap = run_taste(block_in, 0, force_exact)
                return top, [(cnt, cap)]
            # mixture of tastes (no attention): prediction = max over tastes (recommendation_graphs.py:107), so the top-k
            # lies in the union of the per-taste top-k lists: one fused sweep per taste, then a de-duplicating merge
            per_taste = [run_taste(block_in, t, force_exact) for t in range(self.n_tastes)]
            stacked = torch.stack([top.buf for top, _, _ in per_taste]).contiguous()          # [T, U_block, 2k]
            merged = kernels.topk_merge_received(stacked, block_in.shape[0], self.n_tastes, k, dedup=True)
            return merged, [(cnt, cap) for _, cnt, cap in per_taste]

        def exchange(top, u0, u1):
            if gather_group is None:
                return top, np.arange(u0, u1)
            merged, (lo, hi) = distributed.exchange_and_merge(top, gather_group)
            if gather == 'all':
                return distributed.all_gather_rows(merged, u1 - u0, gather_group), np.arange(u0, u1)
            return merged, np.arange(u0 + lo, u0 + hi)

        results, counters, rows = [], [], []
        for (u0, u1, block_in) in blocks:
            top, sweeps = run_block(block_in)
            top, user_rows = exchange(top, u0, u1)
            results.append(top)
            counters.append(sweeps)
            rows.append(user_rows)

        # one synchronisation for the whole call: how many rows the certificate rejected per sweep (device counters);
        # a block with more rejected rows than the device-side fallback holds is re-run through the exact kernel
        live = [c for sweeps in counters for c, _ in sweeps if c is not None]
        if live:
            counts = iter(torch.stack([c[0] for c in live]).cpu().numpy().tolist())
            overflow = []
            for b, sweeps in enumerate(counters):
                for c, cap in sweeps:
                    if c is None:
                        continue
                    n_bad = next(counts)
                    info['fallback_rows'] += min(n_bad, cap)
                    if n_bad > cap and b not in overflow:
                        overflow.append(b)
            if gather_group is not None:     # every rank must take the same decision: the exchange is collective
                overflow = distributed.union_of_indices(overflow, len(blocks), gather_group, device)
            for b in overflow:
                u0, u1, block_in = blocks[b]
                top, _ = run_block(block_in, force_exact=True)
                results[b], rows[b] = exchange(top, u0, u1)
            info['overflow_blocks'] = len(overflow)
        info['user_rows'] = rows[0] if len(rows) == 1 else np.concatenate(rows)
        top_s = results[0].scores if len(results) == 1 else torch.cat([r.scores for r in results])
        top_i = results[0].items if len(results) == 1 else torch.cat([r.items for r in results])
        if not to_host:
            return TopK(top_i, top_s)
        return TopK(*kernels.to_host(top_i, top_s))

    def _topk_from_dense(self, user_in, item_in, k, item_id_offset, device):
        """Any model the fused kernel does not cover: dense scores -> exact full ranks -> the rank <= k entries."""
        n_users, n_items = user_in.shape[0], item_in.shape[0]
        top = kernels.PackedTopK(n_users, k, device)
        top.scores.fill_(float('-inf'))
        top.items.fill_(2 ** 31 - 1)
        if n_items > 0:
            scores = self._predict_device(user_in, item_in, device)
            ranks = kernels.rank_full(scores).long()
            sel = ranks <= k
            rows, cols = sel.nonzero(as_tuple=True)
            pos = ranks[rows, cols] - 1
            top.scores[rows, pos] = scores[rows, cols]
            top.items[rows, pos] = (cols + item_id_offset).to(torch.int32)
        return top

    def predict_similar_items(self, item_features, item_ids, n_similar):
        """tensorrec/tensorrec.py:666-703: for each id, the n_similar (item_id, score) pairs of highest prediction
        between that item's representation and every item's."""
        if self.tf_prediction is None:
            raise ModelNotFitException(method='predict_similar_items')
        device = self._cuda_device()
        item_in = self._single_input(item_features, 'item_features')
        pred_graph = self.prediction_graph_factory
        builtin = type(pred_graph) in _BUILTIN_PRED
        extra = 1 if (builtin and pred_graph.b200_kind == 'cosine') else 0
        item_repr, _, _ = self._represent(self.item_repr_graph_factory, item_in, self.n_item_features, 'item', device,
                                          extra)
        ids = torch.as_tensor(np.asarray(item_ids), dtype=torch.long, device=device)
        gathered = item_repr[ids].contiguous()
        if builtin:
            sims = kernels.score_exact(gathered, item_repr, mode=1 if pred_graph.b200_kind == 'euclidean' else 0)
        else:
            with torch.no_grad():
                sims = pred_graph.connect_dense_prediction_graph(tf_user_representation=gathered,
                                                                 tf_item_representation=item_repr)
        sims = sims.cpu().numpy()
        results = []
        for i in range(len(item_ids)):
            item_sims = sims[i]
            best = np.argpartition(item_sims, -n_similar)[-n_similar:]
            results.append(sorted(zip(best, item_sims[best]), key=lambda x: -x[1]))
        return results

    def predict_user_representation(self, user_features):
        """[n_users, n_components] (or [n_tastes, n_users, n_components] when n_tastes > 1) (tensorrec.py:735-762)."""
        if self.tf_prediction is None:
            raise ModelNotFitException(method='predict_user_representation')
        device = self._cuda_device()
        user_in = self._single_input(user_features, 'user_features')
        self._check_features(user_in, self.n_user_features, 'user')
        user_repr = torch.stack([self._represent(self.user_repr_graph_factory, user_in, self.n_user_features,
                                                 'user_{}'.format(t), device)[0]
                                 for t in range(self.n_tastes)]).cpu().numpy()
        if self.n_tastes == 1:
            user_repr = np.sum(user_repr, axis=0)
        return user_repr

    def predict_user_attention_representation(self, user_features):
        """tensorrec.py:764-793."""
        if self.tf_prediction is None:
            raise ModelNotFitException(method='predict_user_attention_representation')
        if self.attention_graph_factory is None:
            raise ModelWithoutAttentionException()
        device = self._cuda_device()
        user_in = self._single_input(user_features, 'user_features')
        attn = torch.stack([self._represent(self.attention_graph_factory, user_in, self.n_user_features,
                                            'attn_{}'.format(t), device)[0]
                            for t in range(self.n_tastes)]).cpu().numpy()
        if self.n_tastes == 1:
            attn = np.sum(attn, axis=0)
        return attn

    def predict_item_representation(self, item_features):
        """[n_items, n_components] (tensorrec.py:795-816)."""
        if self.tf_prediction is None:
            raise ModelNotFitException(method='predict_item_representation')
        device = self._cuda_device()
        item_in = self._single_input(item_features, 'item_features')
        self._check_features(item_in, self.n_item_features, 'item')
        return self._represent(self.item_repr_graph_factory, item_in, self.n_item_features, 'item',
                               device)[0].cpu().numpy()

    def predict_user_bias(self, user_features):
        """[n_users] (tensorrec.py:818-842)."""
        if self.tf_prediction is None:
            raise ModelNotFitException(method='predict_user_bias')
        if not self.biased:
            raise ModelNotBiasedException(actor='user')
        device = self._cuda_device()
        user_in = self._single_input(user_features, 'user_features')
        return self._projected_biases(user_in, 'feature_biases_user', device).cpu().numpy()

    def predict_item_bias(self, item_features):
        """[n_items] (tensorrec.py:844-868)."""
        if self.tf_prediction is None:
            raise ModelNotFitException(method='predict_item_bias')
        if not self.biased:
            raise ModelNotBiasedException(actor='item')
        device = self._cuda_device()
        item_in = self._single_input(item_features, 'item_features')
        return self._projected_biases(item_in, 'feature_biases_item', device).cpu().numpy()

    # ------------------------------------------------------------------------------------------------
    # persistence (tensorrec.py:870-917): weights as .npz beside the pickled Python object
    # ------------------------------------------------------------------------------------------------
    def __getstate__(self):
        state = dict(self.__dict__)
        state['_variables'] = collections.OrderedDict()
        state['_optimizer'] = None
        state['_optimizer_params'] = None
        state['_was_fit'] = self.tf_prediction is not None
        state.pop('_stream_buffers', None)
        state['_wmrb_step'] = None
        for name in self._all_hook_names():
            state[name] = None
        return state

    def save_model(self, directory_path):
        if self.tf_prediction is None:
            raise ModelNotFitException(method='save_model')
        if not os.path.exists(directory_path):
            os.makedirs(directory_path)
        np.savez(os.path.join(directory_path, 'tensorrec_session.npz'), **self.get_weights())
        with open(os.path.join(directory_path, 'tensorrec.pkl'), 'wb') as file:
            pickle.dump(file=file, obj=self)

    @classmethod
    def load_model(cls, directory_path):
        with open(os.path.join(directory_path, 'tensorrec.pkl'), 'rb') as file:
            model = pickle.load(file=file)
        with np.load(os.path.join(directory_path, 'tensorrec_session.npz')) as data:
            weights = collections.OrderedDict((k, data[k]) for k in data.files)
        model.set_weights(weights, n_user_features=model.n_user_features, n_item_features=model.n_item_features)
        return model
