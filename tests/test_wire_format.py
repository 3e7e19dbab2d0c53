"""/compact_data wire format: host-only C-ABI functions against the oracle restatement of
PointOdometry.cc:732-762 and PointMapping.cc:171-238.  No GPU needed (pure host byte shuffling)."""
import numpy as np
import pytest

from lio_mapping_b200 import wire, _lib


def _clouds(rng, nc, ns, nf):
    mk = lambda n: np.concatenate([rng.uniform(-50, 50, (n, 3)), rng.uniform(0, 64, (n, 1))], 1).astype(np.float32)
    return mk(nc), mk(ns), mk(nf)


@pytest.mark.parametrize("nc,ns,nf", [(5, 7, 11), (0, 3, 0), (1, 0, 0), (1000, 4000, 30000)])
def test_encode_matches_reference_layout(oracle, nc, ns, nf):
    rng = np.random.default_rng(nc + ns + nf)
    q = rng.normal(size=4); q /= np.linalg.norm(q)
    tf7 = np.concatenate([q, rng.normal(size=3)]).astype(np.float32)
    c, s, f = _clouds(rng, nc, ns, nf)
    e = wire.compact_encode(tf7, c, s, f)
    eo = oracle.compact_encode(tf7, c, s, f)
    assert e.shape == eo.shape == (3 + nc + ns + nf, 4)
    assert np.array_equal(e.view(np.uint32), eo.view(np.uint32))       # bit-exact, including the stale intensity of point 2
    assert np.array_equal(e[0, :3], tf7[4:]) and e[0, 3] == 0
    assert np.array_equal(e[1], tf7[:4])
    assert e[2, 0] == nc and e[2, 1] == ns and e[2, 2] == nf and e[2, 3] == tf7[3]


@pytest.mark.parametrize("nc,ns,nf", [(5, 7, 11), (0, 1, 0), (300, 2000, 9000)])
def test_round_trip_and_decoder_parity(oracle, nc, ns, nf):
    rng = np.random.default_rng(17 + nc)
    q = rng.normal(size=4); q /= np.linalg.norm(q)
    tf7 = np.concatenate([q, rng.normal(size=3)]).astype(np.float32)
    c, s, f = _clouds(rng, nc, ns, nf)
    e = oracle.compact_encode(tf7, c, s, f)            # reference-side encoder -> our decoder
    t2, c2, s2, f2 = wire.compact_decode(e)
    to, co, so, fo = oracle.compact_decode(e)
    for a, b in ((t2, to), (c2, co), (s2, so), (f2, fo)):
        assert np.array_equal(a, b)
    assert np.array_equal(t2, tf7) and np.array_equal(c2, c) and np.array_equal(s2, s) and np.array_equal(f2, f)
    # our encoder -> reference-side decoder
    to, co, so, fo = oracle.compact_decode(wire.compact_encode(tf7, c, s, f))
    assert np.array_equal(to, tf7) and np.array_equal(co, c) and np.array_equal(so, s) and np.array_equal(fo, f)


def test_decoder_error_paths_match_reference(oracle):
    rng = np.random.default_rng(1)
    tf7 = np.array([0, 0, 0, 1, 1, 2, 3], np.float32)
    c, s, f = _clouds(rng, 4, 5, 6)
    e = wire.compact_encode(tf7, c, s, f)
    # fewer than 4 points: "compact_points not enough" (a header-only message with three empty clouds is rejected too)
    for bad in (e[:3], e[:0], wire.compact_encode(tf7, c[:0], s[:0], f[:0])):
        assert oracle.compact_decode(bad) is None
        with pytest.raises(_lib.LioError):
            wire.compact_decode(bad)
    # size mismatch: "compact data error"
    for bad in (e[:-1], np.concatenate([e, e[-1:]])):
        assert oracle.compact_decode(bad) is None
        with pytest.raises(_lib.LioError):
            wire.compact_decode(bad)
    neg = e.copy(); neg[2, 0] = -1.0; neg = np.concatenate([neg, neg[-1:]])[: 3 - 1 + 5 + 6 + 3]
    with pytest.raises(_lib.LioError):
        wire.compact_decode(neg)
    # sizes are carried as floats: 2^24 and above are refused by the encoder instead of silently rounding
    big = np.zeros((1, 4), np.float32)
    rc = _lib.lib().lio_compact_encode(tf7, big, 1 << 24, big, 0, big, 0, np.zeros((8, 4), np.float32), 8, __import__("ctypes").byref(__import__("ctypes").c_int()))
    assert rc == -3


def test_pcl32_record_layout():
    rng = np.random.default_rng(2)
    c, _, _ = _clouds(rng, 100, 0, 0)
    rec = wire.to_pcl32(c)
    assert rec.shape == (100, 32)
    as_f = rec.view(np.float32).reshape(100, 8)
    assert np.array_equal(as_f[:, 0:3], c[:, 0:3])             # x, y, z at byte offsets 0, 4, 8
    assert np.all(as_f[:, 3] == 1.0)                           # PCL_ADD_POINT4D padding lane data[3] = 1.0f
    assert np.array_equal(as_f[:, 4], c[:, 3])                 # intensity at byte offset 16
    assert np.array_equal(wire.from_pcl32(rec), c)
