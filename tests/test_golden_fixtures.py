"""Committed golden vectors (tests/golden/oracle_fixtures.npz, made by tests/golden/make_golden.py).

CPU: the oracle still reproduces them (drift check).  GPU: the CUDA path, through the TensorRec class, against the
same committed vectors -- bit-exact for the integer fixture (scores, full ranks, top-k incl. ties), 1e-5 for floats."""
import os

import numpy as np
import pytest
import scipy.sparse as sp

import oracle

F = np.load(os.path.join(os.path.dirname(__file__), 'golden', 'oracle_fixtures.npz'))


def coo(prefix):
    shape = tuple(int(x) for x in F[prefix + '_shape'])
    return sp.coo_matrix((F[prefix + '_val'], (F[prefix + '_row'], F[prefix + '_col'])), shape=shape)


def test_oracle_reproduces_committed_fixtures():
    om = oracle.OracleModel([F['int_wu']], F['int_wi'], F['int_bu'], F['int_bi'])
    scores = om.predict(coo('int_uf'), coo('int_if'))
    assert np.array_equal(scores, F['int_scores'])
    assert np.array_equal(oracle.rank_predictions(scores), F['int_ranks'])
    assert np.array_equal(oracle.rank_predictions_closed_form(scores), F['int_ranks'])
    c = oracle.coo_from_sparse(coo('spmm_f'))
    assert np.array_equal(oracle.linear_representation(c, F['spmm_w']), F['spmm_linear'])
    assert np.array_equal(oracle.normalized_linear_representation(c, F['spmm_w']), F['spmm_normalized'])
    om = oracle.OracleModel(list(F['e2e_wu']), F['e2e_wi'], F['e2e_bu'], F['e2e_bi'], user_repr='normalized_linear',
                            prediction='cosine')
    pred = om.predict(coo('e2e_uf'), coo('e2e_if'))
    assert np.array_equal(pred, F['e2e_scores'])
    assert np.array_equal(oracle.rank_predictions(pred), F['e2e_ranks'])


@pytest.mark.gpu
def test_cuda_path_against_committed_fixtures():
    import tensorrec_b200 as T
    from tensorrec_b200 import kernels
    import torch
    # integer fixture: exact everywhere
    model = T.TensorRec(n_components=16)
    model.set_weights({'linear_weights_user_0': F['int_wu'], 'linear_weights_item': F['int_wi'],
                       'feature_biases_user': F['int_bu'][:, None], 'feature_biases_item': F['int_bi'][:, None]})
    uf, itf = coo('int_uf'), coo('int_if')
    assert np.array_equal(model.predict(uf, itf), F['int_scores'])
    assert np.array_equal(model.predict_rank(uf, itf), F['int_ranks'])
    top = model.predict_rank(uf, itf, k=10)
    rows = np.arange(uf.shape[0])[:, None]
    assert np.array_equal(F['int_ranks'][rows, top.items], np.tile(np.arange(1, 11), (uf.shape[0], 1)))
    assert np.array_equal(top.scores, F['int_scores'][rows, top.items])
    # SpMM edge cases through K1 (unsorted COO, duplicates, empty rows)
    csr = kernels.DeviceCSR.from_scipy(coo('spmm_f'))
    w = torch.from_numpy(F['spmm_w']).cuda()
    lin, _, _ = kernels.gather_reduce(csr, w)
    nrm, _, _ = kernels.gather_reduce(csr, w, n_normalize=1)
    assert np.allclose(lin.cpu().numpy(), F['spmm_linear'], rtol=0, atol=4e-6 * np.abs(F['spmm_linear']).max())
    assert np.allclose(nrm.cpu().numpy(), F['spmm_normalized'], rtol=0, atol=4e-6)
    # end-to-end float: 3 tastes, NormalizedLinear users, cosine, biased
    model = T.TensorRec(n_components=10, n_tastes=3,
                        user_repr_graph=T.representation_graphs.NormalizedLinearRepresentationGraph(),
                        prediction_graph=T.prediction_graphs.CosineSimilarityPredictionGraph())
    weights = {'linear_weights_item': F['e2e_wi'], 'feature_biases_user': F['e2e_bu'][:, None],
               'feature_biases_item': F['e2e_bi'][:, None]}
    for t in range(3):
        weights['linear_weights_user_%d' % t] = F['e2e_wu'][t]
    model.set_weights(weights)
    got = model.predict(coo('e2e_uf'), coo('e2e_if'))
    assert np.all(np.abs(got - F['e2e_scores']) <= 1e-5 + 2e-6)     # cosine scores: |u| = |i| = 1
    ranks = model.predict_rank(coo('e2e_uf'), coo('e2e_if'))
    assert np.array_equal(ranks, oracle.rank_predictions(got))
    assert (ranks != F['e2e_ranks']).mean() < 0.01
