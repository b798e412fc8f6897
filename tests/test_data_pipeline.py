"""The input side of the hot path: TFRecord framing, tf.train.Example decoding and
the reference's ``input_fn`` contract (fm/fm.py:100-112, din/din.py:52-80) through the product's
multi-threaded host decoder (csrc/records.cu behind the C ABI), checked against

  * the committed fixtures: tests/golden/criteo_shard256.tfrecord (the first 256 frames of the
    reference's only data file, byte for byte) and criteo_shard256.npz (what the oracle's
    independent Python reader made of them);
  * the oracle's reader (oracle/tfrecord.py) on all 10 000 records of
    /root/reference/xdeepfm/part-r-00000 where that file exists (the authoring container);
  * the CRC-32C check value of the iSCSI polynomial (RFC 3720: crc32c("123456789") = 0xE3069283).

Host logic only: runs without a GPU."""
import os
import struct

import numpy as np
import pytest
import torch

from oracle import criteo, tfrecord

GOLD = os.path.join(os.path.dirname(__file__), "golden")
FIX = os.path.join(GOLD, "criteo_shard256.tfrecord")
SHARD = "/root/reference/xdeepfm/part-r-00000"


@pytest.fixture(scope="module")
def data(built_lib):
    from recsys_b200 import data
    return data


def test_masked_crc32c_known_answers(data):
    crc = 0xE3069283                                   # RFC 3720 check value
    want = ((((crc >> 15) | (crc << 17)) & 0xFFFFFFFF) + 0xA282EAD8) & 0xFFFFFFFF
    assert data.masked_crc32c(b"123456789") == want
    assert data.masked_crc32c(b"") == 0xA282EAD8       # crc32c("") = 0
    for s in (b"a", b"abcdefgh", b"x" * 31, bytes(range(256)) * 5):
        assert data.masked_crc32c(s) == tfrecord.masked_crc32c(s)


def test_fixture_frames_carry_valid_crcs_and_corruption_is_detected(data, tmp_path):
    rf = data.RecordFile(FIX)
    assert len(rf) == 256
    raw = bytearray(open(FIX, "rb").read())
    payloads = list(data.iter_tfrecords(FIX))
    assert len(payloads) == 256 and payloads == tfrecord.read_records(FIX, verify_crc=True)
    for pos, what in ((int(rf.off[3]) + 5, "payload"), (int(rf.off[7]) - 12, "length")):
        bad = bytearray(raw)
        bad[pos] ^= 0x40
        p = tmp_path / ("bad_%s.tfrecord" % what)
        p.write_bytes(bytes(bad))
        with pytest.raises(ValueError, match="crc|truncated"):
            data.RecordFile(str(p))
        with pytest.raises(ValueError, match="crc|truncated"):
            list(data.iter_tfrecords(str(p)))
        if what == "payload":
            assert len(data.RecordFile(str(p), verify_crc=False)) == 256
    p = tmp_path / "cut.tfrecord"
    p.write_bytes(bytes(raw[:-7]))
    with pytest.raises(ValueError, match="truncated"):
        data.RecordFile(str(p))


def _check_batches(batches, feats_want, labels_want):
    row = 0
    for f, lab in batches:
        B = lab.shape[0]
        sl = slice(row, row + B)
        assert lab.dtype == torch.float32 and tuple(lab.shape) == (B, 1)
        assert np.array_equal(lab.numpy(), labels_want[sl])
        for k in criteo.CONT:
            assert tuple(f[k].shape) == (B, 1)
            assert np.array_equal(f[k].numpy(), feats_want[k][sl]), k
        for k in criteo.CAT:
            got = f[k]
            assert got.shape == (B, 1)
            assert [bytes(v) for v in got.reshape(-1)] == [bytes(v) for v in feats_want[k][sl].reshape(-1)], k
        row += B
    return row


@pytest.mark.parametrize("batch_size,threads", [(256, 0), (100, 1), (7, 3)])
def test_criteo_input_fn_matches_the_frozen_fixture(data, batch_size, threads):
    z = np.load(os.path.join(GOLD, "criteo_shard256.npz"))
    feats_want = {k: z[k] for k in criteo.CONT}
    feats_want.update({k: np.array([bytes(v) for v in z[k]], dtype=object).reshape(-1, 1)
                       for k in criteo.CAT})
    it = data.criteo_input_fn([FIX], batch_size, num_epochs=1, n_threads=threads)
    n = _check_batches(it, feats_want, z["labels"])
    assert n == 256
    # the native decoder and the plain-Python one agree record by record
    ex = [data.parse_example(p) for p in data.iter_tfrecords(FIX)]
    assert all(np.float32(e["_c0"][0]) == z["labels"][i, 0] for i, e in enumerate(ex))
    assert all((e.get("_c20") or [b"NULL"])[0] == bytes(feats_want["_c20"][i, 0])
               for i, e in enumerate(ex))


def test_epochs_shuffle_and_cross_file_batches(data):
    two = list(data.criteo_input_fn([FIX, FIX], 200, num_epochs=2))
    assert [int(l.shape[0]) for _, l in two] == [200, 200, 112, 200, 200, 112]
    a = [l.clone() for _, l in data.criteo_input_fn([FIX], 32, num_epochs=1)]
    b = [l.clone() for _, l in data.criteo_input_fn([FIX], 32, num_epochs=1, need_shuffle=True,
                                                    shuffle_buffer=4, seed=1)]
    assert len(a) == len(b) == 8
    # whole batches are permuted (shuffle follows batch, fm/fm.py:108-110), nothing is lost
    assert sorted(float(x.sum()) for x in a) == sorted(float(x.sum()) for x in b)
    assert any(not torch.equal(x, y) for x, y in zip(a, b))
    endless = data.criteo_input_fn([FIX], 256)            # num_epochs = -1 repeats for ever
    assert sum(1 for _, _ in zip(range(5), endless)) == 5


@pytest.mark.skipif(not os.path.exists(SHARD), reason="reference data file not present here")
def test_whole_reference_shard_against_the_oracle_reader(data):
    payloads = tfrecord.read_records(SHARD, verify_crc=True)
    assert len(payloads) == 10000
    feats_want, labels_want = tfrecord.criteo_batch(payloads)
    n = _check_batches(data.criteo_input_fn([SHARD], 1000, num_epochs=1), feats_want, labels_want)
    assert n == 10000
    assert abs(float(labels_want.mean()) - 0.2182) < 1e-4


# ----------------------------------------------------------------------------- DIN
def _vi(x):
    x &= (1 << 64) - 1
    out = bytearray()
    while True:
        b = x & 0x7F
        x >>= 7
        out.append(b | (0x80 if x else 0))
        if not x:
            return bytes(out)


def _ld(tag, body):
    return bytes([tag]) + _vi(len(body)) + body


def _int64_feature(key, vals, packed=True):
    if packed:
        lst = _ld(0x0A, b"".join(_vi(v) for v in vals)) if vals else b""
    else:
        lst = b"".join(b"\x08" + _vi(v) for v in vals)
    return _ld(0x0A, _ld(0x0A, key.encode()) + _ld(0x12, _ld(0x1A, lst)))


def _example(feats):
    return _ld(0x0A, b"".join(feats))


def _frame(payload):
    ln = struct.pack("<Q", len(payload))
    return ln + struct.pack("<I", tfrecord.masked_crc32c(ln)) + payload + \
        struct.pack("<I", tfrecord.masked_crc32c(payload))


def _din_file(path, recs, packed=True):
    with open(path, "wb") as f:
        for r in recs:
            f.write(_frame(_example([_int64_feature(k, v, packed) for k, v in r.items()])))


def test_din_input_fn_round_trip(data, tmp_path):
    from recsys_b200.din import din
    rng = np.random.default_rng(3)
    P = 9
    recs = []
    for i in range(50):
        n = int(rng.integers(1, P + 1))
        recs.append({"label": [int(rng.integers(0, 2))], "i_id": [int(rng.integers(1, 63002))],
                     "i_cate": [int(rng.integers(1, 802))],
                     "u_iid_seq": [int(v) for v in rng.integers(1, 63002, size=n)] + [0] * (P - n),
                     "u_icat_seq": [int(v) for v in rng.integers(1, 802, size=n)] + [0] * (P - n)})
    recs[4]["i_id"] = [-3]                             # negative int64: a 10-byte varint
    for packed in (True, False):
        path = str(tmp_path / ("din_%d.tfrecord" % packed))
        _din_file(path, recs, packed)
        got = list(din.input_fn([path], 16, num_epochs=1))
        assert [int(l.shape[0]) for _, l in got] == [16, 16, 16, 2]
        row = 0
        for f, lab in got:
            for j in range(lab.shape[0]):
                r = recs[row + j]
                assert int(lab[j]) == r["label"][0]
                assert int(f["i_id"][j]) == r["i_id"][0] and int(f["i_cate"][j]) == r["i_cate"][0]
                assert f["u_iid_seq"][j].tolist() == r["u_iid_seq"]
                assert f["u_icat_seq"][j].tolist() == r["u_icat_seq"]
            assert f["u_iid_seq"].dtype == torch.int64 and tuple(f["u_iid_seq"].shape[1:]) == (P,)
            row += lab.shape[0]
        # the oracle's reader sees the same records
        o = [tfrecord.parse_example(p) for p in tfrecord.read_records(path, verify_crc=True)]
        assert [int(e["i_id"][0]) for e in o] == [r["i_id"][0] for r in recs]
    # .batch() (not padded_batch): a ragged batch is an error, as in TF
    recs[20]["u_iid_seq"] = recs[20]["u_iid_seq"][:-2]
    path = str(tmp_path / "ragged.tfrecord")
    _din_file(path, recs)
    with pytest.raises(ValueError, match="history lengths differ"):
        list(din.input_fn([path], 16, num_epochs=1))
    # a record without one of the FixedLenFeatures (no default) is an error
    del recs[1]["i_cate"]
    recs[20]["u_iid_seq"] = recs[20]["u_iid_seq"] + [0, 0]
    path = str(tmp_path / "missing.tfrecord")
    _din_file(path, recs)
    with pytest.raises(ValueError, match="lacks"):
        list(din.input_fn([path], 16, num_epochs=1))


def test_criteo_missing_float_feature_is_an_error(data, tmp_path):
    payloads = tfrecord.read_records(FIX, limit=3)
    # drop the _c7 entry of the second record by rewriting its key (same length, unknown name)
    broken = payloads[1].replace(b"\x0a\x03_c7\x12", b"\x0a\x03_x7\x12")
    assert broken != payloads[1]
    p = tmp_path / "nofloat.tfrecord"
    p.write_bytes(b"".join(_frame(x) for x in (payloads[0], broken, payloads[2])))
    with pytest.raises(ValueError, match="lacks one of the float features"):
        list(data.criteo_input_fn([str(p)], 3, num_epochs=1))
