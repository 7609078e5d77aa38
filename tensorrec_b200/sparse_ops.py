"""Differentiable sparse x dense product for the training step: tf.sparse_tensor_dense_matmul and its
gradient w.r.t. the dense operand, both evaluated by the CSR gather-reduce kernel K1 (trk_csr_gather_reduce_f32).

    forward   out = A . W            K1 on the CSR of A          (tensorrec/representation_graphs.py:40)
    backward  dW  = A^T . d_out      K1 on the CSR of A^T        (the op's registered gradient in TensorFlow)

The transposed CSR keeps the entries of a feature column in ascending row order, so the backward accumulates in a
fixed order: training is run-to-run deterministic (a scatter-add with atomics would not be).  Sparse tensors that do
not come from input_utils.SparseInput (or CPU tensors: the CPU test-suite trains tiny models) use torch.sparse.mm."""
import torch


class _CsrMatmul(torch.autograd.Function):
    @staticmethod
    def forward(ctx, dense, source, device):
        from . import kernels
        ctx.source, ctx.device = source, device
        out, _, _ = kernels.gather_reduce(source.device_csr(device), dense.detach().contiguous(), want_f32=True)
        return out

    @staticmethod
    def backward(ctx, grad_out):
        from . import kernels
        grad_dense, _, _ = kernels.gather_reduce(ctx.source.device_csr_t(ctx.device), grad_out.contiguous(),
                                                 want_f32=True)
        return grad_dense, None, None


def sparse_dense_matmul(tf_features, dense):
    """tf.sparse_tensor_dense_matmul(tf_features, dense), differentiable w.r.t. `dense`."""
    source = getattr(tf_features, '_trk_source', None)
    if (source is not None and dense.is_cuda and dense.dtype == torch.float32 and dense.dim() == 2
            and dense.shape[1] >= 1 and tf_features.shape[0] > 0):
        return _CsrMatmul.apply(dense, source, dense.device)
    return torch.sparse.mm(tf_features, dense)
