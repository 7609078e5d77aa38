"""Ranking metrics over predict_rank output (API of tensorrec/eval.py).

Consumers of the hot path, host-side numpy/scipy.  The reference builds `predicted_ranks * positive.A` dense
(eval.py:22,48) and only ever tests `rank <= k` (eval.py:23,49,68-69); that is what makes the top-k output of
predict_rank(k=...) sufficient: every function here accepts either the full int32 rank matrix or a TopK result."""
import numpy as np
import scipy.sparse as sp


def _ranks_of_positives(predicted_ranks, positive):
    """csr matrix with the predicted rank of every positive test interaction (0 entries are not stored).

    For a TopK input, positives outside the top-k get rank n_items + 1 (any value > k would do)."""
    positive = sp.csr_matrix(positive)
    if hasattr(predicted_ranks, 'items') and hasattr(predicted_ranks, 'scores'):       # TopK
        items = np.asarray(predicted_ranks.items)
        n_users, n_items = positive.shape
        k = items.shape[1]
        rows = np.repeat(np.arange(n_users), k)
        valid = (items.reshape(-1) >= 0) & (items.reshape(-1) < n_items)
        rank_of = sp.csr_matrix((np.tile(np.arange(1, k + 1), n_users)[valid],
                                 (rows[valid], items.reshape(-1)[valid])), shape=positive.shape)
        inside = positive.multiply(rank_of)                    # rank where the positive is in the top-k
        outside = positive - positive.multiply(rank_of > 0)    # positives not in the top-k
        return sp.csr_matrix(inside + outside * (n_items + 1))
    return sp.csr_matrix(np.asarray(predicted_ranks) * positive.toarray())


def precision_at_k(predicted_ranks, test_interactions, k=10, preserve_rows=False):
    """eval.py:7-30."""
    positive_test_interactions = sp.csr_matrix(test_interactions > 0)
    ranks = _ranks_of_positives(predicted_ranks, positive_test_interactions)
    ranks.data = np.less(ranks.data, (k + 1)).astype(ranks.data.dtype)
    precision = np.squeeze(np.array(ranks.sum(axis=1))).astype(float) / k
    if not preserve_rows:
        precision = precision[positive_test_interactions.getnnz(axis=1) > 0]
    return precision


def recall_at_k(predicted_ranks, test_interactions, k=10, preserve_rows=False):
    """eval.py:33-58."""
    positive_test_interactions = sp.csr_matrix(test_interactions > 0)
    ranks = _ranks_of_positives(predicted_ranks, positive_test_interactions)
    ranks.data = np.less(ranks.data, (k + 1)).astype(ranks.data.dtype)
    retrieved = np.squeeze(positive_test_interactions.getnnz(axis=1))
    hit = np.squeeze(np.array(ranks.sum(axis=1)))
    if not preserve_rows:
        hit = hit[positive_test_interactions.getnnz(axis=1) > 0]
        retrieved = retrieved[positive_test_interactions.getnnz(axis=1) > 0]
    return hit.astype(float) / retrieved.astype(float)


def _setup_ndcg(predicted_ranks, test_interactions, k=10):
    """eval.py:61-72."""
    test_interactions = sp.csr_matrix(test_interactions)
    pos_inter = sp.csr_matrix(test_interactions > 0)
    ror = _ranks_of_positives(predicted_ranks, pos_inter).astype(np.float64)
    relevance = sp.csr_matrix(test_interactions.multiply(pos_inter)).astype(np.float64)
    ror.sort_indices()
    relevance.sort_indices()
    k_mask = np.less(ror.data, k + 1)
    ror_at_k = np.maximum(np.multiply(ror.data, k_mask), 1)
    return relevance, k_mask, ror, ror_at_k


def _idcg(hits, k=10):
    """eval.py:75-78."""
    sorted_hits = hits[np.argsort(-hits)][:min(len(hits), k)]
    return np.sum((2 ** sorted_hits - 1) / np.log2(np.arange(len(sorted_hits)) + 2))


def _dcg(relevance, k_mask, ror_at_k, ror):
    """eval.py:81-87."""
    numer = (2 ** np.multiply(relevance.data, k_mask)) - 1
    denom = np.log2(ror_at_k + 1)
    ror.data = numer / denom
    return ror.sum(axis=1).flatten()


def ndcg_at_k(predicted_ranks, test_interactions, k=10, preserve_rows=False):
    """eval.py:89-117."""
    relevance, k_mask, ranks_of_relevant, ror_at_k = _setup_ndcg(predicted_ranks, test_interactions, k)
    dcg = np.asarray(_dcg(relevance, k_mask, ror_at_k, ranks_of_relevant))[0]
    idcg = np.apply_along_axis(_idcg, 1, relevance.toarray())
    with np.errstate(divide='ignore', invalid='ignore'):
        ndcg = dcg / idcg
    if not preserve_rows:
        positive_test_interactions = sp.csr_matrix(sp.csr_matrix(test_interactions) > 0)
        ndcg = ndcg[positive_test_interactions.getnnz(axis=1) > 0]
    return ndcg


def f1_score_at_k(predicted_ranks, test_interactions, k=10, preserve_rows=False):
    """eval.py:120-148."""
    p_at_k = precision_at_k(predicted_ranks=predicted_ranks, test_interactions=test_interactions, k=k,
                            preserve_rows=preserve_rows)
    r_at_k = recall_at_k(predicted_ranks=predicted_ranks, test_interactions=test_interactions, k=k,
                         preserve_rows=preserve_rows)
    mean_p, mean_r = np.mean(p_at_k), np.mean(r_at_k)
    return (2.0 * mean_p * mean_r) / (mean_p + mean_r)


def fit_and_eval(model, user_features, item_features, train_interactions, test_interactions, fit_kwargs, recall_k=30,
                 precision_k=5, ndcg_k=30):
    """eval.py:151-166."""
    model.fit(user_features=user_features, item_features=item_features, interactions=train_interactions, **fit_kwargs)
    predicted_ranks = model.predict_rank(user_features=user_features, item_features=item_features)
    p_at_k = precision_at_k(predicted_ranks, test_interactions, k=precision_k)
    r_at_k = recall_at_k(predicted_ranks, test_interactions, k=recall_k)
    n_at_k = ndcg_at_k(predicted_ranks, test_interactions, k=ndcg_k)
    p_at_k_insample = precision_at_k(predicted_ranks, train_interactions, k=precision_k)
    r_at_k_insample = recall_at_k(predicted_ranks, train_interactions, k=recall_k)
    n_at_k_insample = ndcg_at_k(predicted_ranks, train_interactions, k=ndcg_k)
    return (np.mean(r_at_k), np.mean(p_at_k), np.mean(n_at_k), np.mean(r_at_k_insample), np.mean(p_at_k_insample),
            np.mean(n_at_k_insample))


def eval_random_ranks_on_dataset(interactions, recall_k=30, precision_k=5, ndcg_k=30):
    """eval.py:181-192."""
    n_users, n_items = interactions.shape
    random_guesses = np.array([np.random.choice(a=n_items, size=n_items, replace=False) + 1 for _ in range(n_users)])
    return (np.mean(recall_at_k(random_guesses, interactions, k=recall_k)),
            np.mean(precision_at_k(random_guesses, interactions, k=precision_k)),
            np.mean(ndcg_at_k(random_guesses, interactions, k=ndcg_k)))
