"""Loss plugin graphs (API of tensorrec/loss_graphs.py): class flags and connect_loss_graph(**kwargs) keep the
reference's names and visibility rules (tensorrec.py:463-485).  Losses belong to the training step, which is NOT
the predict / predict_rank hot path: they are plain differentiable torch code (SURVEY.md 8f, rank 1 lists the
fused sampled-rank training kernels as the next row)."""
import math

import torch


class AbstractLossGraph(object):
    # If True, dense prediction results will be passed to the loss function
    is_dense = False
    # If True, randomly sampled predictions will be passed to the loss function
    is_sample_based = False
    # If True, and if is_sample_based is True, predictions will be sampled with replacement
    is_sampled_with_replacement = False

    def connect_loss_graph(self, tf_prediction_serial, tf_interactions_serial, tf_interactions, tf_n_users, tf_n_items,
                           tf_prediction, tf_rankings, tf_sample_predictions, tf_n_sampled_items):
        pass


def _moments(x):
    """tf.nn.moments: mean and biased variance."""
    mean = torch.mean(x)
    return mean, torch.mean((x - mean) ** 2)


def _normal_overlap_loss(positive_predictions, negative_predictions):
    """1 - Normal(neg_mean - pos_mean, sqrt(neg_var + pos_var)).cdf(0)  (loss_graphs.py:85-98)."""
    pos_mean, pos_var = _moments(positive_predictions)
    neg_mean, neg_var = _moments(negative_predictions)
    loc = neg_mean - pos_mean
    scale = torch.sqrt(neg_var + pos_var)
    cdf0 = 0.5 * (1.0 + torch.erf((0.0 - loc) / (scale * math.sqrt(2.0))))
    return 1.0 - cdf0


class RMSELossGraph(AbstractLossGraph):
    """Root mean square error on the given interactions (loss_graphs.py:53-59)."""

    def connect_loss_graph(self, tf_prediction_serial, tf_interactions_serial, **kwargs):
        return torch.sqrt(torch.mean((tf_interactions_serial - tf_prediction_serial) ** 2))


class RMSEDenseLossGraph(AbstractLossGraph):
    """RMSE against the dense interaction matrix, missing entries as 0 (loss_graphs.py:62-72)."""
    is_dense = True

    def connect_loss_graph(self, tf_interactions, tf_prediction, **kwargs):
        error = tf_interactions.to_dense() - tf_prediction
        return torch.sqrt(torch.mean(error ** 2))


class SeparationLossGraph(AbstractLossGraph):
    """Overlap of the normal fits of positive / non-positive interaction predictions (loss_graphs.py:75-98)."""

    def connect_loss_graph(self, tf_prediction_serial, tf_interactions_serial, **kwargs):
        positive = tf_prediction_serial[tf_interactions_serial > 0.0]
        negative = tf_prediction_serial[tf_interactions_serial <= 0.0]
        return _normal_overlap_loss(positive, negative)


class SeparationDenseLossGraph(AbstractLossGraph):
    """Separation loss over the dense matrix, non-interacted items as negatives (loss_graphs.py:101-134)."""
    is_dense = True

    def connect_loss_graph(self, tf_prediction, tf_interactions, **kwargs):
        interactions_serial = tf_interactions.to_dense().reshape(-1)
        prediction_serial = tf_prediction.reshape(-1)
        positive = prediction_serial[interactions_serial > 0.0]
        negative = prediction_serial[interactions_serial <= 0.0]
        return _normal_overlap_loss(positive, negative)


class WMRBLossGraph(AbstractLossGraph):
    """Sampled weighted-margin-rank-batch loss (loss_graphs.py:137-180)."""
    is_sample_based = True

    def connect_loss_graph(self, tf_prediction_serial, tf_interactions, tf_sample_predictions, tf_n_items,
                           tf_n_sampled_items, **kwargs):
        return self.weighted_margin_rank_batch(tf_prediction_serial=tf_prediction_serial,
                                               tf_interactions=tf_interactions,
                                               tf_sample_predictions=tf_sample_predictions,
                                               tf_n_items=tf_n_items,
                                               tf_n_sampled_items=tf_n_sampled_items)

    @staticmethod
    def _positive(tf_interactions):
        indices = tf_interactions._indices()
        values = tf_interactions._values()
        mask = values > 0.0
        return mask, indices[:, mask], values[mask]

    def weighted_margin_rank_batch(self, tf_prediction_serial, tf_interactions, tf_sample_predictions, tf_n_items,
                                   tf_n_sampled_items):
        mask, positive_indices, _ = self._positive(tf_interactions)
        positive_predictions = tf_prediction_serial[mask]                          # [n_positive]
        mapped_samples = tf_sample_predictions[positive_indices[0]]                # [n_positive, n_sampled]
        summation_term = torch.clamp(1.0 - positive_predictions.unsqueeze(1) + mapped_samples, min=0.0)
        sampled_margin_rank = (float(tf_n_items) / float(tf_n_sampled_items)) * torch.sum(summation_term, dim=1)
        return torch.log(sampled_margin_rank + 1.0)


class BalancedWMRBLossGraph(WMRBLossGraph):
    """WMRB weighted by interaction magnitude / per-item interaction mass (loss_graphs.py:183-227)."""

    def weighted_margin_rank_batch(self, tf_prediction_serial, tf_interactions, tf_sample_predictions, tf_n_items,
                                   tf_n_sampled_items):
        mask, positive_indices, positive_values = self._positive(tf_interactions)
        n_items_total = tf_interactions.shape[1]
        listening_sum_per_item = torch.zeros(n_items_total, device=positive_values.device).index_add_(
            0, positive_indices[1], positive_values)
        gathered_sums = listening_sum_per_item[positive_indices[1]]
        positive_predictions = tf_prediction_serial[mask]
        mapped_samples = tf_sample_predictions[positive_indices[0]]
        summation_term = torch.clamp(1.0 - positive_predictions.unsqueeze(1) + mapped_samples, min=0.0)
        sampled_margin_rank = ((float(tf_n_items) / float(tf_n_sampled_items)) * torch.sum(summation_term, dim=1)
                               * positive_values / gathered_sums)
        return torch.log(sampled_margin_rank + 1.0)
