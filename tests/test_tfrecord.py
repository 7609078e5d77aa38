"""TFRecord input files in the reference's layout (tensorrec/input_utils.py:72-127) without TensorFlow: checksums,
the protobuf encoding of tf.train.Example, round trips through the TensorRec API."""
import os
import struct

import numpy as np
import pytest
import scipy.sparse as sp

from tensorrec_b200 import TensorRec, tfrecord
from tensorrec_b200.input_utils import (
    TensorRecDataset, create_tensorrec_dataset_from_sparse_matrix, create_tensorrec_dataset_from_tfrecord,
    write_tfrecord_from_sparse_matrix, write_tfrecord_from_tensorrec_dataset)
from tensorrec_b200.util import generate_dummy_data


def test_crc32c_known_values_and_chunked_path():
    assert tfrecord.crc32c(b'') == 0
    assert tfrecord.crc32c(b'a') == 0xC1D04330
    assert tfrecord.crc32c(b'123456789') == 0xE3069283                 # the CRC-32C check value
    assert tfrecord.crc32c(bytes(32)) == 0x8A9136AA                    # RFC 3720 B.4: 32 bytes of zeros
    assert tfrecord.crc32c(bytes([0xFF] * 32)) == 0x62A8AB43           # RFC 3720 B.4: 32 bytes of ones
    assert tfrecord.crc32c(bytes(range(32))) == 0x46DD794E             # RFC 3720 B.4: incrementing bytes
    rng = np.random.default_rng(0)
    for n in (3, 4, 5, 65535, 65536, 70001, 300000):                   # both sides of the chunk-parallel path
        data = rng.integers(0, 256, n, dtype=np.uint8)
        slow = tfrecord._crc_raw_small(data, 0xFFFFFFFF) ^ 0xFFFFFFFF
        assert tfrecord.crc32c(data.tobytes()) == slow
    crc = tfrecord.crc32c(b'123456789')
    assert tfrecord.masked_crc32c(b'123456789') == (((crc >> 15) | (crc << 17)) + 0xA282EAD8) & 0xFFFFFFFF


def test_varints_and_example_bytes_by_hand():
    values = np.array([0, 1, 127, 128, 300, 2 ** 31 - 1, 2 ** 40, -1, -5], dtype=np.int64)
    assert np.array_equal(tfrecord.decode_varints(tfrecord.encode_varints(values)), values)
    assert tfrecord.encode_varints([300]) == b'\xac\x02' and tfrecord.encode_varints([-1]) == b'\xff' * 9 + b'\x01'
    # Example{features{feature{key: "d0", value{int64_list{value: [3]}}}}} encoded by hand from the protobuf spec:
    # 1A 03 0A 01 03 = Feature.int64_list(3){Int64List.value(1) packed [3]}; 0A 02 'd0' 12 05 <feature> = map entry
    expect = bytes.fromhex('0a0d' '0a0b' '0a026430' '1205' '1a030a0103')
    assert tfrecord.encode_example({'d0': ('int64', [3])}) == expect
    assert list(tfrecord.decode_example(expect)['d0']) == [3]
    # the unpacked encodings (one tag per element) must be read too: Int64List{08 03 08 04}, FloatList{0D 00 00 80 3F}
    unpacked = bytes.fromhex('0a1b' '0a0b' '0a0161' '1206' '1a04' '08030804' '0a0c' '0a0162' '1207' '1205' '0d0000803f')
    got = tfrecord.decode_example(unpacked)
    assert list(got['a']) == [3, 4] and list(got['b']) == [1.0]


def test_round_trip_single_and_multi_record(tmp_path):
    rng = np.random.default_rng(5)
    m1 = sp.random(40, 70, density=.1, format='coo', random_state=rng, dtype=np.float32)
    m2 = sp.coo_matrix((3, 70), dtype=np.float32)                                      # empty matrix
    path = str(tmp_path / 'features.tfrecord')
    assert write_tfrecord_from_sparse_matrix(path, m1) == path
    (ds,) = create_tensorrec_dataset_from_tfrecord(path)
    ref = create_tensorrec_dataset_from_sparse_matrix(m1)
    assert isinstance(ds, TensorRecDataset) and (ds.d0, ds.d1) == (40, 70)
    assert ds.row_index.dtype == np.int64 and ds.values.dtype == np.float32
    for a, b in zip(ds[:3], ref[:3]):
        assert np.array_equal(a, b)
    both = str(tmp_path / 'two.tfrecord')
    write_tfrecord_from_tensorrec_dataset(both, [ref, create_tensorrec_dataset_from_sparse_matrix(m2)])
    first, second = create_tensorrec_dataset_from_tfrecord(both)
    assert np.array_equal(first.values, ref.values) and second.values.size == 0 and (second.d0, second.d1) == (3, 70)
    # framing: length | masked crc | payload | masked crc
    raw = open(path, 'rb').read()
    (length,) = struct.unpack('<Q', raw[:8])
    assert len(raw) == 8 + 4 + length + 4
    assert struct.unpack('<I', raw[8:12])[0] == tfrecord.masked_crc32c(raw[:8])
    assert struct.unpack('<I', raw[-4:])[0] == tfrecord.masked_crc32c(raw[12:12 + length])


def test_corruption_is_detected(tmp_path):
    path = str(tmp_path / 'm.tfrecord')
    write_tfrecord_from_sparse_matrix(path, sp.identity(5, dtype=np.float32, format='coo'))
    raw = bytearray(open(path, 'rb').read())
    raw[20] ^= 0x01
    open(path, 'wb').write(bytes(raw))
    with pytest.raises(ValueError, match='crc'):
        create_tensorrec_dataset_from_tfrecord(path)
    open(path, 'wb').write(bytes(raw[:15]))
    with pytest.raises(ValueError, match='truncated'):
        create_tensorrec_dataset_from_tfrecord(path)


def test_fit_from_tfrecords_like_the_reference_test(tmp_path):
    """test/test_tensorrec.py:169-172: fit from three TFRecord paths (and mixed with in-memory inputs)."""
    interactions, user_features, item_features = generate_dummy_data(
        num_users=15, num_items=30, interaction_density=.5, num_user_features=40, num_item_features=30,
        n_features_per_user=10, n_features_per_item=10, pos_int_ratio=.5)
    paths = {}
    for name, matrix in (('interactions', interactions), ('user_features', user_features),
                         ('item_features', item_features)):
        paths[name] = write_tfrecord_from_sparse_matrix(str(tmp_path / (name + '.tfrecord')), matrix)
    model = TensorRec(n_components=5)
    model.fit(paths['interactions'], paths['user_features'], paths['item_features'], epochs=3)
    assert model.tf_prediction is not None
    model.fit_partial(interactions, paths['user_features'], item_features, epochs=1)
    assert os.path.exists(paths['interactions'])
