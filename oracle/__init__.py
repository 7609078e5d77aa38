"""
oracle/ -- TEST INFRASTRUCTURE ONLY.

A CPU (numpy/scipy) restatement of the reference's predict / predict_rank arithmetic
(jfkirk/tensorrec @ 80690737).  It is the checker for the CUDA path, never the product:
only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import it.  Nothing under ``tensorrec_b200/`` imports this package.

Parity pinning status (see DESIGN.md "Oracle"):
  * pinned by the reference's own known-answer tests (ported as data in tests/golden/):
    dot / cosine dense prediction, project_biases, bias_prediction_dense, rank_predictions,
    collapse_mixture_of_tastes (max and attention softmax), predict_similar_items; and, for the training
    step, the serial dot / cosine / euclidean predictions, bias_prediction_serial,
    densify_sampled_item_predictions, split_sparse_tensor_indices.
  * PARITY UNPINNED (no reference test holds a number, and TensorFlow -- the reference's only
    numeric back-end -- is not installable here, so the reference cannot be run):
    LinearRepresentationGraph / NormalizedLinearRepresentationGraph values for d > 1,
    tie-breaking inside rank_predictions (taken from tf.nn.top_k's documented contract: lower index
    first), end-to-end predict()/predict_rank() after fit().
"""
from .reference_ops import (  # noqa: F401
    coo_from_sparse,
    sparse_dense_matmul,
    l2_normalize,
    linear_representation,
    normalized_linear_representation,
    dot_product_dense,
    cosine_dense,
    euclidean_dense,
    dot_product_serial,
    cosine_serial,
    euclidean_serial,
    split_sparse_tensor_indices,
    bias_prediction_serial,
    densify_sampled_item_predictions,
    collapse_mixture_of_tastes,
    project_biases,
    bias_prediction_dense,
    rank_predictions,
    rank_predictions_closed_form,
    top_k_from_scores,
    top_k_from_scores_fast,
    predict,
    predict_rank,
    predict_similar_items,
    OracleModel,
)
from . import loss_ops  # noqa: F401,E402  (training-step losses; parity unpinned)
