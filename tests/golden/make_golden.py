"""Generates tests/golden/oracle_fixtures.npz -- committed input/output vectors of the predict / predict_rank path.

    python tests/golden/make_golden.py

The reference (jfkirk/tensorrec) cannot be imported here: tensorrec/tensorrec.py:8 imports TensorFlow 1.x, which is
not installed and not installable (no network; no py3.12 build).  These fixtures are therefore produced by the
ORACLE (oracle/reference_ops.py), which is itself pinned to the reference's own known-answer tests
(tests/golden/reference_known_answers.json).  They (a) freeze the oracle against drift and (b) give the GPU tests
fixed vectors for the three gaps the reference's tests leave open: integer-valued ties, SpMM with
duplicates / empty rows / unsorted COO, end-to-end predict with injected weights.  Inputs are stored sparse (COO
triplets) so the file stays small."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

import oracle  # noqa: E402
from tests import helpers as H  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'oracle_fixtures.npz')


def coo_triplet(prefix, m, store):
    row, col, val, d0, d1 = oracle.coo_from_sparse(m)
    store[prefix + '_row'], store[prefix + '_col'], store[prefix + '_val'] = row, col, val
    store[prefix + '_shape'] = np.array([d0, d1], dtype=np.int64)


def main():
    store = {}

    # 1. integer-valued fixture with many ties: exact scores, exact ranks (tie -> lower item id)
    uf = H.tag_features(48, 40, 4, seed=11, integer=True)
    itf = H.tag_features(300, 40, 4, seed=12, integer=True)
    wu, wi = H.linear_weights(40, 16, seed=13, integer=True), H.linear_weights(40, 16, seed=14, integer=True)
    bu, bi = H.feature_biases(40, seed=15, integer=True), H.feature_biases(40, seed=16, integer=True)
    om = oracle.OracleModel([wu], wi, bu, bi)
    scores = om.predict(uf, itf)
    coo_triplet('int_uf', uf, store)
    coo_triplet('int_if', itf, store)
    store.update(int_wu=wu, int_wi=wi, int_bu=bu, int_bi=bi, int_scores=scores,
                 int_ranks=oracle.rank_predictions(scores))
    assert (np.diff(np.sort(scores, axis=1), axis=1) == 0).mean() > 0.5, 'fixture is meant to be full of ties'

    # 2. SpMM edge cases: unsorted COO with duplicates and empty rows, d = 12, with and without normalisation
    m = H.messy_coo(37, 23, 300, seed=5)
    w = H.linear_weights(23, 12, seed=6)
    coo_triplet('spmm_f', m, store)
    store.update(spmm_w=w, spmm_linear=oracle.linear_representation(oracle.coo_from_sparse(m), w),
                 spmm_normalized=oracle.normalized_linear_representation(oracle.coo_from_sparse(m), w))

    # 3. end-to-end float: 3 tastes, NormalizedLinear users, cosine, biased
    uf = H.tag_features(20, 30, 6, seed=21)
    itf = H.tag_features(45, 30, 6, seed=22)
    wus = [H.linear_weights(30, 10, seed=23 + t) for t in range(3)]
    wi = H.linear_weights(30, 10, seed=27)
    bu, bi = H.feature_biases(30, seed=28), H.feature_biases(30, seed=29)
    om = oracle.OracleModel(wus, wi, bu, bi, user_repr='normalized_linear', prediction='cosine')
    coo_triplet('e2e_uf', uf, store)
    coo_triplet('e2e_if', itf, store)
    pred = om.predict(uf, itf)
    store.update(e2e_wu=np.stack(wus), e2e_wi=wi, e2e_bu=bu, e2e_bi=bi, e2e_scores=pred,
                 e2e_ranks=oracle.rank_predictions(pred))

    np.savez_compressed(OUT, **store)
    print('wrote', OUT, '%.1f KB' % (os.path.getsize(OUT) / 1e3))


if __name__ == '__main__':
    main()
