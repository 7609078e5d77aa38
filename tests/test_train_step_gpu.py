"""GPU parity tests of the sampled-rank training step kernels against the oracle
(oracle/loss_ops.wmrb_step_reference / adam_reference, themselves pinned on the CPU against torch autograd over the host
mirror of the reference's graph functions: tests/test_train_step_cpu.py)."""
import numpy as np
import pytest
import scipy.sparse as sp

from oracle import loss_ops
from tests import helpers as H

pytestmark = pytest.mark.gpu
F32 = np.float32


@pytest.fixture(scope='module')
def T():
    import torch
    import tensorrec_b200
    from tensorrec_b200 import kernels, session_management as sm
    kernels.require_cuda()
    torch.cuda.set_device(0)
    sm.set_session(None)
    return tensorrec_b200


def make_case(seed, n_users, n_items, d, biased=True, density=0.05):
    from tensorrec_b200 import util
    interactions, uf, itf = util.generate_dummy_data(num_users=n_users, num_items=n_items, interaction_density=density,
                                                     num_user_features=40, num_item_features=30,
                                                     n_features_per_user=6, n_features_per_item=5, seed=seed)
    rng = np.random.default_rng(seed + 100)
    wu = (0.3 * rng.standard_normal((uf.shape[1], d))).astype(F32)
    wi = (0.3 * rng.standard_normal((itf.shape[1], d))).astype(F32)
    bu = (0.2 * rng.standard_normal(uf.shape[1])).astype(F32) if biased else None
    bi = (0.2 * rng.standard_normal(itf.shape[1])).astype(F32) if biased else None
    return sp.csr_matrix(interactions), uf, itf, wu, wi, bu, bi


@pytest.mark.parametrize('replace', [True, False])
def test_sampler_equals_its_host_stream_and_is_uniform(T, replace):
    import torch
    from tensorrec_b200 import train_kernels as TK
    dev = torch.device('cuda', 0)
    got = TK.sample_items_device(37, 50, 9, replace, seed=1234, step=5, device=dev).cpu().numpy()
    assert np.array_equal(got, TK.sample_items_host(37, 50, 9, replace, seed=1234, step=5))
    assert np.array_equal(got, TK.sample_items_device(37, 50, 9, replace, 1234, 5, dev).cpu().numpy())   # deterministic
    assert not np.array_equal(got, TK.sample_items_device(37, 50, 9, replace, 1234, 6, dev).cpu().numpy())  # per step
    assert not np.array_equal(got[0], got[1])                                                             # per user
    n_items, n_users, n_s = 64, 20000, 16
    s = TK.sample_items_device(n_items, n_users, n_s, replace, seed=7, step=0, device=dev).cpu().numpy()
    assert s.shape == (n_users, n_s) and s.min() >= 0 and s.max() < n_items
    if not replace:
        assert np.all(np.diff(np.sort(s, axis=1), axis=1) > 0)              # no item twice for one user
        full = TK.sample_items_device(12, 100, 12, False, seed=3, step=0, device=dev).cpu().numpy()
        assert np.array_equal(np.sort(full, axis=1), np.tile(np.arange(12), (100, 1)))
        with pytest.raises(ValueError):
            TK.sample_items_device(5, 3, 6, False, 0, 0, dev)
    counts = np.bincount(s.reshape(-1), minlength=n_items)
    expect = n_users * n_s / n_items
    assert np.all(np.abs(counts - expect) < 6 * np.sqrt(expect))            # every item equally likely
    # every column position is uniform too (Floyd's late positions prefer late ids only conditionally)
    big = TK.sample_items_device(1000003, 3, 2000, replace, seed=9, step=1, device=dev).cpu().numpy()
    assert big.min() >= 0 and big.max() < 1000003


def run_kernel_step(T, case, n_sampled, balanced, bf16, samples=None, lr=0.05, l2=0.0, seed=11):
    import torch
    from tensorrec_b200 import train_kernels as TK
    from tensorrec_b200.input_utils import SparseInput
    interactions, uf, itf, wu, wi, bu, bi = case
    lg = T.loss_graphs.BalancedWMRBLossGraph() if balanced else T.loss_graphs.WMRBLossGraph()
    model = T.TensorRec(n_components=wu.shape[1], loss_graph=lg, biased=bu is not None)
    weights = {'linear_weights_user_0': wu, 'linear_weights_item': wi}
    if bu is not None:
        weights.update({'feature_biases_user': bu[:, None], 'feature_biases_item': bi[:, None]})
    model.set_weights(weights)
    dev = torch.device('cuda', 0)
    stepper = TK.WmrbStep(model, dev, seed=seed, bf16=bf16)
    if samples is not None:
        samples = torch.from_numpy(np.ascontiguousarray(samples, dtype=np.int32)).to(dev)
    loss, pred = stepper.step(SparseInput(interactions), SparseInput(uf), SparseInput(itf), n_sampled, lr, l2,
                              samples=samples)
    return model, stepper, loss.cpu().numpy(), pred.cpu().numpy()


def csr_order(interactions):
    """The kernel reports per-interaction values in CSR order (stable row sort of the COO order)."""
    coo = sp.coo_matrix(interactions)
    return np.argsort(coo.row, kind='stable')


@pytest.mark.parametrize('d,n_sampled', [(12, 9), (128, 33), (200, 64)])
@pytest.mark.parametrize('balanced', [False, True])
@pytest.mark.parametrize('biased', [True, False])
def test_wmrb_step_matches_the_oracle_fp32(T, d, n_sampled, balanced, biased):
    case = make_case(seed=d, n_users=300, n_items=257, d=d, biased=biased)
    interactions, uf, itf, wu, wi, bu, bi = case
    rng = np.random.default_rng(5)
    samples = np.stack([rng.choice(itf.shape[0], n_sampled, replace=False) for _ in range(uf.shape[0])])
    ref = loss_ops.wmrb_step_reference(uf, itf, interactions, wu, wi, bu, bi, samples, balanced=balanced)
    model, stepper, loss, pred = run_kernel_step(T, case, n_sampled, balanced, bf16=False, samples=samples)
    order = csr_order(interactions)
    mask = ref['positive_mask'][order]
    assert np.allclose(pred, ref['pred_serial'][order], rtol=2e-5, atol=2e-6)
    full_loss = np.zeros(len(order), F32)
    full_loss[ref['positive_mask']] = ref['loss']
    assert np.allclose(loss, full_loss[order], rtol=2e-5, atol=2e-6)
    assert np.all(loss[~mask] == 0.0)
    g = {k: v.cpu().numpy() for k, v in stepper.last['grads'].items()}
    for name, key in (('linear_weights_user_0', 'd_w_user'), ('linear_weights_item', 'd_w_item')):
        scale = max(1.0, float(np.abs(ref[key]).max()))
        assert np.allclose(g[name], ref[key], rtol=2e-4, atol=2e-5 * scale), name
    if biased:
        for name, key in (('feature_biases_user', 'd_b_user'), ('feature_biases_item', 'd_b_item')):
            scale = max(1.0, float(np.abs(ref[key]).max()))
            assert np.allclose(g[name], ref[key], rtol=2e-4, atol=2e-5 * scale), name


def test_wmrb_step_bf16_representations(T):
    """BASELINE config #4 allows bf16: the representations are rounded once (nearest even) and gathered as bf16; the
    oracle evaluated on the same rounded representations must agree as tightly as in fp32."""
    case = make_case(seed=3, n_users=260, n_items=300, d=128)
    interactions, uf, itf, wu, wi, bu, bi = case
    rng = np.random.default_rng(6)
    samples = np.stack([rng.choice(itf.shape[0], 20, replace=False) for _ in range(uf.shape[0])])
    ref = loss_ops.wmrb_step_reference(uf, itf, interactions, wu, wi, bu, bi, samples,
                                       round_repr=loss_ops.round_to_bfloat16)
    model, stepper, loss, pred = run_kernel_step(T, case, 20, False, bf16=True, samples=samples)
    order = csr_order(interactions)
    # K1 and the oracle's scipy product round the fp32 representations differently in the last bit; where such a value
    # sits on a bf16 rounding boundary the two bf16 representations differ by one bf16 ulp (2^-8 relative) in that element.
    # So: nearly every prediction agrees to fp32 accuracy, all of them to a few bf16 ulps of one element.
    err = np.abs(pred - ref['pred_serial'][order])
    assert np.mean(err <= 2e-5 * np.abs(pred) + 2e-6) > 0.9
    assert err.max() < 0.02
    g = stepper.last['grads']['linear_weights_item'].cpu().numpy()
    scale = max(1.0, float(np.abs(ref['d_w_item']).max()))
    assert np.abs(g - ref['d_w_item']).max() < 5e-3 * scale
    assert np.mean(np.abs(g - ref['d_w_item']) <= 2e-4 * np.abs(ref['d_w_item']) + 2e-5 * scale) > 0.9
    exact = loss_ops.wmrb_step_reference(uf, itf, interactions, wu, wi, bu, bi, samples)
    diff = np.abs(pred - exact['pred_serial'][order])
    assert diff.max() < 0.25 and diff.mean() > 1e-5            # bf16 is close to fp32, and it is not fp32


def test_adam_step_and_the_whole_update_match_the_oracle(T):
    import torch
    case = make_case(seed=9, n_users=120, n_items=90, d=16)
    interactions, uf, itf, wu, wi, bu, bi = case
    rng = np.random.default_rng(2)
    samples = np.stack([rng.choice(itf.shape[0], 8, replace=False) for _ in range(uf.shape[0])])
    lr, l2 = 0.1, 0.3
    model, stepper, _, _ = run_kernel_step(T, case, 8, False, bf16=False, samples=samples, lr=lr, l2=l2)
    ref = loss_ops.wmrb_step_reference(uf, itf, interactions, wu, wi, bu, bi, samples)
    new = model.get_weights()
    for name, w0, key in (('linear_weights_user_0', wu, 'd_w_user'), ('linear_weights_item', wi, 'd_w_item'),
                          ('feature_biases_user', bu[:, None], 'd_b_user'), ('feature_biases_item', bi[:, None], 'd_b_item')):
        grad = stepper.last['grads'][name].cpu().numpy().reshape(w0.shape)      # the kernel's own gradient: isolates Adam
        exp, _, _ = loss_ops.adam_reference(w0, grad, np.zeros_like(w0), np.zeros_like(w0), 1, lr, l2=l2)
        assert np.allclose(new[name], exp, rtol=1e-6, atol=1e-7), name
        exp_ref, _, _ = loss_ops.adam_reference(w0, ref[key].reshape(w0.shape), np.zeros_like(w0), np.zeros_like(w0), 1,
                                                lr, l2=l2)
        # first Adam step = lr * sign(g) wherever |g| >> eps: insensitive to gradient rounding except at g ~ 0
        assert np.mean(np.abs(new[name] - exp_ref) < 1e-4) > 0.98, name
    # a second step uses the moments of the first
    w1 = {k: v.copy() for k, v in new.items()}
    from tensorrec_b200.input_utils import SparseInput
    st = torch.from_numpy(samples.astype(np.int32)).cuda()
    stepper.step(SparseInput(interactions), SparseInput(uf), SparseInput(itf), 8, lr, l2, samples=st)
    g2 = stepper.last['grads']['linear_weights_item'].cpu().numpy()
    g1 = loss_ops.wmrb_step_reference(uf, itf, interactions, wu, wi, bu, bi, samples)['d_w_item']
    _, m1, v1 = loss_ops.adam_reference(wi, g1, np.zeros_like(wi), np.zeros_like(wi), 1, lr, l2=l2)
    exp2, _, _ = loss_ops.adam_reference(w1['linear_weights_item'], g2, m1, v1, 2, lr, l2=l2)
    assert np.allclose(model.get_weights()['linear_weights_item'], exp2, rtol=1e-4, atol=1e-5)


def test_fit_with_wmrb_runs_on_the_kernels_and_learns(T):
    from tensorrec_b200 import util
    interactions, uf, itf = util.generate_dummy_data(num_users=200, num_items=300, interaction_density=.05, seed=4)
    for lg in (T.loss_graphs.WMRBLossGraph(), T.loss_graphs.BalancedWMRBLossGraph()):
        model = T.TensorRec(n_components=8, loss_graph=lg)
        model.fit(interactions, uf, itf, epochs=1, n_sampled_items=20, learning_rate=0.05)
        assert model._wmrb_step is not None and model._wmrb_step.t == 1, 'the kernel training path was not taken'
        first = float(model._wmrb_step.last['loss'].sum())
        model.fit_partial(interactions, uf, itf, epochs=30, n_sampled_items=20, learning_rate=0.05)
        assert model._wmrb_step.t == 31
        assert float(model._wmrb_step.last['loss'].sum()) < 0.8 * first          # the summed WMRB loss goes down
        ranks = model.predict_rank(uf, itf)
        assert ranks.shape == (200, 300) and ranks.min() == 1
        # positives end up ranked better than chance
        pos = sp.coo_matrix(interactions)
        keep = pos.data > 0
        assert ranks[pos.row[keep], pos.col[keep]].mean() < 0.4 * 300
    # user batching and a failing first step keep working through this path
    model = T.TensorRec(n_components=8, loss_graph=T.loss_graphs.WMRBLossGraph())
    model.fit(interactions, uf, itf, epochs=2, n_sampled_items=10, user_batch_size=64)
    assert model._wmrb_step.t == 2 * 4
    bad = T.TensorRec(n_components=8, loss_graph=T.loss_graphs.WMRBLossGraph())
    with pytest.raises(ValueError):
        bad.fit(interactions, uf, itf, epochs=1, n_sampled_items=301)
    assert bad.tf_prediction is None
