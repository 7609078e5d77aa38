"""Seeded synthetic inputs shared by the tests, the golden-fixture generator and bench.py.

Generators follow the reference's tensorrec/util.py:61-85 (tag regime: sp.rand features) and :88-117 (indicator
regime: identity + random tag columns), but are seeded (the reference is unseeded) -- SURVEY.md 8(d)."""
import numpy as np
import scipy.sparse as sp

F32 = np.float32


def tag_features(rows, n_features=200, per_row=20, seed=0, integer=False):
    """util.py:75-78: sp.rand(rows, n_features, density=per_row/n_features), values U[0,1) (or {1} if integer)."""
    rng = np.random.default_rng(seed)
    m = sp.random(rows, n_features, density=float(per_row) / n_features, format='csr', dtype=np.float64,
                  random_state=rng)
    if integer:
        m.data[:] = 1.0
    return m.astype(F32)


def indicator_features(rows, seed=0, tags_per_row=3):
    """util.py:90-108: identity block + tags_per_row*rows random 1.0 entries in columns [rows, 1.2*rows)."""
    rng = np.random.default_rng(seed)
    n_features = int(rows * 1.2)
    n_tag_cols = max(n_features - rows, 1)
    n_features = rows + n_tag_cols
    n_tags = rows * tags_per_row
    r = np.concatenate([np.arange(rows), rng.integers(0, rows, n_tags)])
    c = np.concatenate([np.arange(rows), rows + rng.integers(0, n_tag_cols, n_tags)])
    m = sp.csr_matrix((np.ones(r.shape[0], dtype=F32), (r, c)), shape=(rows, n_features))
    m.sum_duplicates()            # the reference writes `= 1` into a lil_matrix: duplicates collapse
    m.data[:] = 1.0
    return m.astype(F32)


def linear_weights(n_features, d, seed=2, integer=False):
    """representation_graphs.py:35-36: random_normal rows, L2-normalised (or small integers for exact fixtures)."""
    rng = np.random.default_rng(seed)
    if integer:
        return rng.integers(-2, 3, size=(n_features, d)).astype(F32)
    w = rng.standard_normal((n_features, d)).astype(F32)
    w /= np.maximum(np.linalg.norm(w, axis=1, keepdims=True), 1e-6).astype(F32)
    return w.astype(F32)


def feature_biases(n_features, seed=4, integer=False):
    rng = np.random.default_rng(seed)
    if integer:
        return rng.integers(-3, 4, size=(n_features,)).astype(F32)
    return (0.1 * rng.standard_normal(n_features)).astype(F32)


def messy_coo(rows, n_features, nnz, seed=5):
    """Unsorted COO with duplicates and empty rows -- the SpMM edge cases the reference never pins."""
    rng = np.random.default_rng(seed)
    r = rng.integers(0, rows, nnz)
    r[r % 7 == 3] = 0                       # leaves several rows empty, piles duplicates on row 0
    c = rng.integers(0, max(n_features // 3, 1), nnz)    # few columns -> many (row, col) duplicates
    v = rng.standard_normal(nnz).astype(F32)
    return sp.coo_matrix((v, (r, c)), shape=(rows, n_features))


def norm_tolerance(user_repr, item_repr, rel=1e-5):
    """|score error| bound: rel * |u| * |i| (scores cancel to ~0, so per-element relative error is meaningless)."""
    nu = np.linalg.norm(np.asarray(user_repr, dtype=np.float64), axis=-1)
    ni = np.linalg.norm(np.asarray(item_repr, dtype=np.float64), axis=-1)
    return rel * nu[..., :, None] * ni[None, :]
