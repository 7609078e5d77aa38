"""CPU pinning of the training-step oracle (oracle/loss_ops.wmrb_step_reference): its loss equals the
loss-graph restatement, and its analytic gradients equal torch autograd over the HOST MIRROR of the reference's graph
functions (tensorrec_b200.TensorRec._training_losses: serial gather-dot, bias_prediction_serial, densify, WMRB) for the
same weights and the same sampled items.  The CUDA kernels are then checked against this oracle on the GPU."""
import numpy as np
import pytest
import torch

import oracle
from oracle import loss_ops
from tests import helpers as H

import tensorrec_b200 as T
from tensorrec_b200 import util


@pytest.fixture(autouse=True)
def cpu_session():
    from tensorrec_b200 import session_management as sm
    sm.set_session(sm.Session('cpu'))
    yield
    sm.set_session(None)


def make_case(seed=0, n_users=40, n_items=55, d=12, biased=True):
    interactions, uf, itf = util.generate_dummy_data(num_users=n_users, num_items=n_items, interaction_density=.15,
                                                     num_user_features=30, num_item_features=25,
                                                     n_features_per_user=6, n_features_per_item=5, seed=seed)
    rng = np.random.default_rng(seed + 100)
    wu = (0.3 * rng.standard_normal((uf.shape[1], d))).astype(np.float32)
    wi = (0.3 * rng.standard_normal((itf.shape[1], d))).astype(np.float32)
    bu = (0.2 * rng.standard_normal(uf.shape[1])).astype(np.float32) if biased else None
    bi = (0.2 * rng.standard_normal(itf.shape[1])).astype(np.float32) if biased else None
    return interactions, uf, itf, wu, wi, bu, bi


@pytest.mark.parametrize('balanced', [False, True])
@pytest.mark.parametrize('biased', [True, False])
def test_step_oracle_equals_autograd_of_the_host_mirror(monkeypatch, balanced, biased):
    interactions, uf, itf, wu, wi, bu, bi = make_case(seed=3, biased=biased)
    n_users, n_items, n_sampled = uf.shape[0], itf.shape[0], 9
    samples = util.sample_items(n_items, n_users, n_sampled, replace=False,
                                rng=np.random.default_rng(7))[:, 1].reshape(n_users, n_sampled)
    ref = loss_ops.wmrb_step_reference(uf, itf, interactions, wu, wi, bu, bi, samples, balanced=balanced)

    # the loss-graph restatement on the oracle's own predictions
    coo = oracle.coo_from_sparse(interactions)
    fn = loss_ops.balanced_wmrb if balanced else loss_ops.wmrb
    expect = fn(ref['pred_serial'], coo, ref['sample_pred'], n_items, n_sampled)
    assert np.allclose(ref['loss'], expect, rtol=1e-6, atol=1e-6)

    # torch autograd over the host mirror, fed the same samples
    loss_graph = T.loss_graphs.BalancedWMRBLossGraph() if balanced else T.loss_graphs.WMRBLossGraph()
    model = T.TensorRec(n_components=wu.shape[1], loss_graph=loss_graph, biased=biased)
    weights = {'linear_weights_user_0': wu, 'linear_weights_item': wi}
    if biased:
        weights.update({'feature_biases_user': bu[:, None], 'feature_biases_item': bi[:, None]})
    model.set_weights(weights)
    pairs = np.stack([np.repeat(np.arange(n_users), n_sampled), samples.reshape(-1)], axis=1).astype(np.int64)
    monkeypatch.setattr(T.tensorrec, 'sample_items', lambda *a, **k: pairs)
    from tensorrec_b200.input_utils import SparseInput
    from tensorrec_b200.session_management import variable_scope
    with variable_scope(model._variables):
        basic_loss, wr_loss, pred_serial, tf_weights = model._training_losses(
            SparseInput(interactions), SparseInput(uf), SparseInput(itf), n_sampled, torch.device('cpu'))
    assert np.allclose(basic_loss.detach().numpy(), ref['loss'], rtol=2e-5, atol=2e-6)
    assert np.allclose(pred_serial.detach().numpy(), ref['pred_serial'], rtol=2e-5, atol=2e-6)
    basic_loss.sum().backward()
    grads = {k: v.grad.detach().numpy() for k, v in model._variables.items()}
    scale = max(1.0, float(np.abs(ref['d_w_user']).max()))
    assert np.allclose(grads['linear_weights_user_0'], ref['d_w_user'], rtol=1e-4, atol=1e-5 * scale)
    assert np.allclose(grads['linear_weights_item'], ref['d_w_item'], rtol=1e-4, atol=1e-5 * scale)
    if biased:
        assert np.allclose(grads['feature_biases_user'][:, 0], ref['d_b_user'], rtol=1e-4, atol=1e-5 * scale)
        assert np.allclose(grads['feature_biases_item'][:, 0], ref['d_b_item'], rtol=1e-4, atol=1e-5 * scale)


def test_adam_reference_is_tensorflow_adam():
    """Two hand-computed steps of tf.train.AdamOptimizer (m, v from zero; lr_t = lr sqrt(1 - b2^t) / (1 - b1^t))."""
    w = np.array([1.0, -2.0], np.float32)
    g = np.array([0.5, 0.25], np.float32)
    w1, m1, v1 = loss_ops.adam_reference(w, g, np.zeros(2), np.zeros(2), 1, 0.1)
    # step 1: m = 0.1 g, v = 0.001 g^2, lr_1 = 0.1 * sqrt(0.001) / 0.1 -> update = sqrt(0.001) * 0.1 g / (sqrt(0.001) |g| + 1e-8)
    assert np.allclose(w1, w - 0.1 * np.sign(g), atol=1e-6)
    w2, m2, v2 = loss_ops.adam_reference(w1, g, m1, v1, 2, 0.1, l2=0.5)
    g2 = g + 0.5 * w1
    m_exp = 0.9 * m1 + 0.1 * g2
    v_exp = 0.999 * v1 + 0.001 * g2 * g2
    lr2 = 0.1 * np.sqrt(1 - 0.999 ** 2) / (1 - 0.9 ** 2)
    assert np.allclose(w2, w1 - lr2 * m_exp / (np.sqrt(v_exp) + 1e-8), rtol=1e-6)


def test_bfloat16_rounding_helper():
    x = np.array([1.0, 1.00390625, 1.005859375, -3.1415927, 0.0, 65504.0], np.float32)
    got = loss_ops.round_to_bfloat16(x)
    assert np.array_equal(got, torch.from_numpy(x).to(torch.bfloat16).to(torch.float32).numpy())
