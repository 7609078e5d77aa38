"""TFRecord files in the reference's layout (tensorrec/input_utils.py:72-127), without TensorFlow.

One record = one sparse matrix = one tf.train.Example with five features: 'row_index' and 'col_index' (int64 lists),
'values' (float list), 'd0' and 'd1' (single int64).  A file may hold several records; the reference's Dataset yields
one batch per record.

Formats implemented here (both public and stable):
  * TFRecord framing: uint64 length | uint32 masked crc32c(length) | payload | uint32 masked crc32c(payload), little
    endian, masked crc = rotr15(crc) + 0xa282ead8 (mod 2^32), crc = CRC-32C (Castagnoli);
  * protobuf wire format of tf.train.Example: Example{1: Features}, Features{1: map<string, Feature>} (map entries are
    messages {1: key, 2: value}, in any order), Feature{1: BytesList | 2: FloatList | 3: Int64List},
    FloatList{1: packed float32}, Int64List{1: packed varint}.  Readers must also accept the unpacked encodings.
All bulk work is vectorised numpy (a 4M-entry index list is a few hundred milliseconds, not a Python loop)."""
import struct

import numpy as np

# ---------------------------------------------------------------------------------------------------- CRC-32C
_CRC_POLY = 0x82F63B78   # reflected Castagnoli polynomial


def _make_crc_tables():
    tables = np.zeros((1, 256), dtype=np.uint32)
    for i in range(256):
        c = i
        for _ in range(8):
            c = (c >> 1) ^ _CRC_POLY if c & 1 else c >> 1
        tables[0, i] = c
    return tables


_CRC_TABLES = _make_crc_tables()


def _crc_raw_small(buf, crc):
    """Bytewise table update of a raw (no init / xor-out handling) reflected CRC state."""
    t0 = _CRC_TABLES[0]
    for byte in buf.tolist():
        crc = int(t0[(crc ^ byte) & 0xFF]) ^ (crc >> 8)
    return crc


def _gf2_apply(matrix, vec):
    """matrix: 32 uint32 columns (image of every state bit); returns matrix . vec over GF(2)."""
    out, bit = 0, 0
    while vec:
        if vec & 1:
            out ^= matrix[bit]
        vec >>= 1
        bit += 1
    return out


def _zero_bytes_operator(n_bytes):
    """The linear map 'advance the raw CRC state over n_bytes zero bytes' as 32 columns (square-and-multiply)."""
    t0 = _CRC_TABLES[0]
    one = [int(t0[(1 << b) & 0xFF]) ^ ((1 << b) >> 8) for b in range(32)]      # one zero byte
    result = [1 << b for b in range(32)]                                        # identity
    power = one
    while n_bytes:
        if n_bytes & 1:
            result = [_gf2_apply(power, col) for col in result]
        power = [_gf2_apply(power, col) for col in power]
        n_bytes >>= 1
    return result


def crc32c(data):
    """CRC-32C (Castagnoli) of a bytes-like object.

    Large inputs are cut into equal chunks whose raw CRC states advance together, one numpy step per byte position
    (the recurrence is sequential along a chunk but independent across chunks); the chunk states are then folded with
    the GF(2) operator that advances a state over one chunk of zero bytes."""
    buf = np.frombuffer(bytes(data), dtype=np.uint8)
    n = buf.size
    if n < 4:
        return _crc_raw_small(buf, 0xFFFFFFFF) ^ 0xFFFFFFFF
    # init 0xFFFFFFFF == raw CRC (init 0) of the message with its first four bytes complemented; a raw CRC with zero
    # state ignores leading zero bytes, so the message can be front-padded to a whole number of chunks
    if n < (1 << 16):
        msg = buf.copy()
        msg[:4] ^= 0xFF
        return _crc_raw_small(msg, 0) ^ 0xFFFFFFFF
    n_chunks = 4096
    chunk = -(-n // n_chunks)
    padded = np.zeros(n_chunks * chunk, dtype=np.uint8)
    padded[-n:] = buf
    first = n_chunks * chunk - n
    padded[first:first + 4] ^= 0xFF
    cols = padded.reshape(n_chunks, chunk)
    t0 = _CRC_TABLES[0]
    state = np.zeros(n_chunks, dtype=np.uint32)
    for j in range(chunk):
        state = t0[(state ^ cols[:, j]) & np.uint32(0xFF)] ^ (state >> np.uint32(8))
    advance = _zero_bytes_operator(chunk)
    crc = 0
    for s_c in state.tolist():
        crc = _gf2_apply(advance, crc) ^ s_c
    return crc ^ 0xFFFFFFFF


def masked_crc32c(data):
    crc = crc32c(data)
    return (((crc >> 15) | (crc << 17)) + 0xA282EAD8) & 0xFFFFFFFF


# ---------------------------------------------------------------------------------------------------- varints
def _varint(value):
    value &= 0xFFFFFFFFFFFFFFFF
    out = bytearray()
    while True:
        byte = value & 0x7F
        value >>= 7
        if value:
            out.append(byte | 0x80)
        else:
            out.append(byte)
            return bytes(out)


def encode_varints(values):
    """int64 array -> the concatenated base-128 varints (two's complement for negatives, as protobuf int64)."""
    v = np.asarray(values, dtype=np.int64).astype(np.uint64)
    if v.size == 0:
        return b''
    n_bytes = np.ones(v.shape, dtype=np.int64)
    rest = v >> np.uint64(7)
    while np.any(rest):
        n_bytes += (rest != 0)
        rest = rest >> np.uint64(7)
    ends = np.cumsum(n_bytes)
    starts = ends - n_bytes
    out = np.zeros(int(ends[-1]), dtype=np.uint8)
    for k in range(int(n_bytes.max())):
        sel = n_bytes > k
        chunk = ((v[sel] >> np.uint64(7 * k)) & np.uint64(0x7F)).astype(np.uint8)
        more = (n_bytes[sel] > k + 1).astype(np.uint8) << 7
        out[starts[sel] + k] = chunk | more
    return out.tobytes()


def decode_varints(data):
    """bytes holding back-to-back varints -> int64 array."""
    buf = np.frombuffer(bytes(data), dtype=np.uint8)
    if buf.size == 0:
        return np.zeros(0, dtype=np.int64)
    last = buf < 0x80
    if not last[-1]:
        raise ValueError('truncated varint')
    ends = np.flatnonzero(last)
    starts = np.concatenate([[0], ends[:-1] + 1])
    if np.any(ends - starts >= 10):
        raise ValueError('varint longer than 10 bytes')
    pos_in = np.arange(buf.size) - np.repeat(starts, ends - starts + 1)
    parts = (buf & 0x7F).astype(np.uint64) << (np.uint64(7) * pos_in.astype(np.uint64))
    return np.add.reduceat(parts, starts).astype(np.uint64).astype(np.int64)


def _read_varint(buf, pos):
    result, shift = 0, 0
    while True:
        if pos >= len(buf):
            raise ValueError('truncated varint')
        byte = buf[pos]
        pos += 1
        result |= (byte & 0x7F) << shift
        if byte < 0x80:
            return result, pos
        shift += 7
        if shift > 63:
            raise ValueError('varint longer than 10 bytes')


def _fields(buf):
    """Yields (field number, wire type, value) of one serialized message; value is an int (varint / fixed) or a
    memoryview (length-delimited)."""
    view = memoryview(buf)
    pos = 0
    while pos < len(view):
        tag, pos = _read_varint(view, pos)
        number, wire = tag >> 3, tag & 7
        if wire == 0:
            value, pos = _read_varint(view, pos)
        elif wire == 1:
            value, pos = struct.unpack_from('<Q', view, pos)[0], pos + 8
        elif wire == 2:
            size, pos = _read_varint(view, pos)
            if pos + size > len(view):
                raise ValueError('truncated length-delimited field')
            value, pos = view[pos:pos + size], pos + size
        elif wire == 5:
            value, pos = struct.unpack_from('<I', view, pos)[0], pos + 4
        else:
            raise ValueError('unsupported protobuf wire type %d' % wire)
        yield number, wire, value


def _len_delimited(number, payload):
    return _varint((number << 3) | 2) + _varint(len(payload)) + payload


# ---------------------------------------------------------------------------------------------------- Example
def encode_feature(kind, values):
    """tf.train.Feature holding an Int64List (kind 'int64') or a FloatList (kind 'float'), packed."""
    if kind == 'int64':
        return _len_delimited(3, _len_delimited(1, encode_varints(values)))
    if kind == 'float':
        return _len_delimited(2, _len_delimited(1, np.asarray(values, dtype='<f4').tobytes()))
    raise ValueError('unknown feature kind %r' % (kind,))


def encode_example(features):
    """{name: (kind, values)} -> serialized tf.train.Example (map entries in the given order)."""
    entries = b''.join(_len_delimited(1, _len_delimited(1, name.encode('utf-8')) + _len_delimited(2, encode_feature(*kv)))
                       for name, kv in features.items())
    return _len_delimited(1, entries)


def _decode_feature(buf):
    for number, wire, value in _fields(buf):
        if number == 3 and wire == 2:          # Int64List
            parts = []
            for n2, w2, v2 in _fields(value):
                if n2 == 1 and w2 == 2:
                    parts.append(decode_varints(v2))            # packed
                elif n2 == 1 and w2 == 0:
                    parts.append(np.array([v2], dtype=np.uint64).astype(np.int64))   # unpacked
            return np.concatenate(parts) if parts else np.zeros(0, dtype=np.int64)
        if number == 2 and wire == 2:          # FloatList
            parts = []
            for n2, w2, v2 in _fields(value):
                if n2 == 1 and w2 == 2:
                    parts.append(np.frombuffer(bytes(v2), dtype='<f4'))
                elif n2 == 1 and w2 == 5:
                    parts.append(np.array([v2], dtype='<u4').view('<f4'))
            return np.concatenate(parts).astype(np.float32) if parts else np.zeros(0, dtype=np.float32)
        if number == 1 and wire == 2:          # BytesList: not part of the TensorRec layout
            return [bytes(v2) for n2, w2, v2 in _fields(value) if n2 == 1 and w2 == 2]
    return None


def decode_example(buf):
    """serialized tf.train.Example -> {name: numpy array}."""
    out = {}
    for number, wire, features in _fields(buf):
        if number != 1 or wire != 2:
            continue
        for n2, w2, entry in _fields(features):
            if n2 != 1 or w2 != 2:
                continue
            name, feature = None, None
            for n3, w3, v3 in _fields(entry):
                if n3 == 1 and w3 == 2:
                    name = bytes(v3).decode('utf-8')
                elif n3 == 2 and w3 == 2:
                    feature = _decode_feature(v3)
            if name is not None:
                out[name] = feature
    return out


# ---------------------------------------------------------------------------------------------------- framing
def write_records(path, payloads):
    with open(path, 'wb') as f:
        for payload in payloads:
            header = struct.pack('<Q', len(payload))
            f.write(header)
            f.write(struct.pack('<I', masked_crc32c(header)))
            f.write(payload)
            f.write(struct.pack('<I', masked_crc32c(payload)))
    return path


def read_records(path):
    records = []
    with open(path, 'rb') as f:
        while True:
            header = f.read(8)
            if not header:
                return records
            if len(header) < 8:
                raise ValueError('%s: truncated TFRecord header' % path)
            (length,) = struct.unpack('<Q', header)
            (crc,) = struct.unpack('<I', f.read(4))
            if crc != masked_crc32c(header):
                raise ValueError('%s: corrupted TFRecord length (crc mismatch)' % path)
            payload = f.read(length)
            footer = f.read(4)
            if len(payload) < length or len(footer) < 4:
                raise ValueError('%s: truncated TFRecord payload' % path)
            if struct.unpack('<I', footer)[0] != masked_crc32c(payload):
                raise ValueError('%s: corrupted TFRecord payload (crc mismatch)' % path)
            records.append(payload)
