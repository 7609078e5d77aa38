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


_TFRECORD_KEYS = ('row_index', 'col_index', 'values', 'd0', 'd1')


def write_tfrecord_from_tensorrec_dataset(tfrecord_path, dataset):
    """input_utils.py:72-103: one tf.train.Example with the five TensorRec features, framed as one TFRecord.
    Written without TensorFlow (tfrecord.py); a list of datasets becomes one record (= one batch) each."""
    from . import tfrecord
    datasets = [dataset] if isinstance(dataset, TensorRecDataset) else list(dataset)
    payloads = []
    for ds in datasets:
        payloads.append(tfrecord.encode_example(collections.OrderedDict([
            ('row_index', ('int64', ds.row_index)), ('col_index', ('int64', ds.col_index)),
            ('values', ('float', ds.values)), ('d0', ('int64', [ds.d0])), ('d1', ('int64', [ds.d1]))])))
    return tfrecord.write_records(tfrecord_path, payloads)


def write_tfrecord_from_sparse_matrix(tfrecord_path, sparse_matrix):
    """input_utils.py:43-53."""
    return write_tfrecord_from_tensorrec_dataset(
        tfrecord_path=tfrecord_path, dataset=create_tensorrec_dataset_from_sparse_matrix(sparse_matrix=sparse_matrix))


def create_tensorrec_dataset_from_tfrecord(tfrecord_path):
    """input_utils.py:106-127: the records of a TFRecord file as TensorRecDatasets, one per record (the reference's
    tf.data pipeline yields one batch per record).  Checksums are verified; a record without the five features of
    the layout is an error."""
    from . import tfrecord
    datasets = []
    for payload in tfrecord.read_records(tfrecord_path):
        features = tfrecord.decode_example(payload)
        missing = [key for key in _TFRECORD_KEYS if features.get(key) is None]
        if missing:
            raise ValueError('{}: record without the TensorRec features {}'.format(tfrecord_path, missing))
        row, col, values = features['row_index'], features['col_index'], features['values']
        if not (len(row) == len(col) == len(values)) or len(features['d0']) != 1 or len(features['d1']) != 1:
            raise ValueError('{}: inconsistent feature lengths in a TensorRec record'.format(tfrecord_path))
        datasets.append(TensorRecDataset(np.asarray(row, dtype=np.int64), np.asarray(col, dtype=np.int64),
                                         np.asarray(values, dtype=np.float32), int(features['d0'][0]),
                                         int(features['d1'][0])))
    return datasets


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

    @property
    def n_positive(self):
        """Stored interactions with a value > 0 (the entries WMRB's positive_interaction_mask keeps)."""
        if getattr(self, '_n_positive', None) is None:
            self._n_positive = int(np.count_nonzero(np.asarray(self.matrix.data) > 0))
        return self._n_positive

    def positive_item_sums(self, device):
        """BalancedWMRBLossGraph's per-item sum of the positive interaction values, on the device."""
        key = 'possum:' + str(device)
        if key not in self._csr:
            from .train_kernels import positive_item_sums
            self._csr[key] = torch.from_numpy(positive_item_sums(self.matrix, self.shape[1])).to(device)
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
