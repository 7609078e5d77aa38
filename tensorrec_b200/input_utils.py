"""Input coercion: scipy sparse matrices -> the TensorRec data format (tensorrec/input_utils.py).

The reference turns every input into a 5-tuple (row int64[nnz], col int64[nnz], values float32[nnz], d0, d1) wrapped
in a tf.data.Dataset (input_utils.py:15-40) and rebuilds a tf.SparseTensor from it (tensorrec.py:285-293).  Here the
same 5-tuple is a `TensorRecDataset`; the kernels consume it as device CSR (kernels.DeviceCSR), the training step
(torch autograd, outside the predict hot path) as a torch sparse tensor."""
import collections

import numpy as np
import scipy.sparse as sp
import torch

TensorRecDataset = collections.namedtuple('TensorRecDataset', ['row_index', 'col_index', 'values', 'd0', 'd1'])


def create_tensorrec_dataset_from_sparse_matrix(sparse_matrix):
    """input_utils.py:22-40: COO order as sp.coo_matrix() yields it, values cast to float32."""
    if not isinstance(sparse_matrix, sp.coo_matrix):
        sparse_matrix = sp.coo_matrix(sparse_matrix)
    return TensorRecDataset(np.asarray(sparse_matrix.row, dtype=np.int64),
                            np.asarray(sparse_matrix.col, dtype=np.int64),
                            np.asarray(sparse_matrix.data, dtype=np.float32),
                            int(sparse_matrix.shape[0]), int(sparse_matrix.shape[1]))


def get_dimensions_from_tensorrec_dataset(dataset):
    """input_utils.py:56-69."""
    return dataset.d0, dataset.d1


def sparse_matrix_from_tensorrec_dataset(dataset):
    return sp.coo_matrix((dataset.values, (dataset.row_index, dataset.col_index)), shape=(dataset.d0, dataset.d1))


def _tfrecord_unavailable(*_args, **_kwargs):
    raise NotImplementedError('TFRecord files are a TensorFlow wire format (tf.train.Example + CRC32C framing, '
                              'tensorrec/input_utils.py:72-127); this build has no TensorFlow and does not read or '
                              'write them (SURVEY.md 8f, rank 4)')


write_tfrecord_from_sparse_matrix = _tfrecord_unavailable
write_tfrecord_from_tensorrec_dataset = _tfrecord_unavailable
create_tensorrec_dataset_from_tfrecord = _tfrecord_unavailable


class SparseInput(object):
    """One feature / interaction matrix, convertible to the device formats on demand (each built at most once)."""

    def __init__(self, matrix):
        if isinstance(matrix, TensorRecDataset):
            matrix = sparse_matrix_from_tensorrec_dataset(matrix)
        if not sp.issparse(matrix):
            raise ValueError('Input must be a scipy sparse matrix, an iterable of scipy sprase matrices, or a '
                             'TensorFlow Dataset')
        self.matrix = matrix
        self.shape = (int(matrix.shape[0]), int(matrix.shape[1]))
        self._dataset = None
        self._csr = {}
        self._torch = {}

    @property
    def dataset(self):
        if self._dataset is None:
            self._dataset = create_tensorrec_dataset_from_sparse_matrix(self.matrix)
        return self._dataset

    def device_csr(self, device):
        from .kernels import DeviceCSR
        key = str(device)
        if key not in self._csr:
            self._csr[key] = DeviceCSR.from_scipy(self.matrix, device=device)
        return self._csr[key]

    def device_csr_t(self, device):
        """CSR of the transposed matrix on the device (the backward operand of the K1 training step)."""
        from .kernels import DeviceCSR
        key = 'T:' + str(device)
        if key not in self._csr:
            self._csr[key] = DeviceCSR.from_scipy_transposed(self.matrix, device=device)
        return self._csr[key]

    def torch_sparse(self, device):
        """Uncoalesced COO tensor in reference entry order (duplicates are summed by torch.sparse.mm)."""
        key = str(device)
        if key not in self._torch:
            ds = self.dataset
            idx = torch.from_numpy(np.stack([ds.row_index, ds.col_index]))
            self._torch[key] = torch.sparse_coo_tensor(idx, torch.from_numpy(ds.values), size=self.shape,
                                                       is_coalesced=False, check_invariants=False).to(device)
            # lets sparse_ops.sparse_dense_matmul find the CSR forms of this matrix (K1 forward / backward)
            self._torch[key]._trk_source = self
        return self._torch[key]
