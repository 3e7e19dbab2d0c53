"""/compact_data wire format (PointOdometry.cc:732-762 encoder, PointMapping::CompactDataHandler PointMapping.cc:171-238).

Host-only helpers over the C-ABI (`lio_compact_*`): no device needed.  Clouds are (n, 4) float32 arrays of
(x, y, z, intensity); `tf7` = (qx, qy, qz, qw, px, py, pz) = `transform_sum_`.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib


def _c4(a):
    return np.ascontiguousarray(a, np.float32).reshape(-1, 4)


def compact_encode(tf7, corner, surf, full) -> np.ndarray:
    """Returns the (3 + nc + ns + nf, 4) float32 compact cloud."""
    c, s, f = _c4(corner), _c4(surf), _c4(full)
    n = 3 + c.shape[0] + s.shape[0] + f.shape[0]
    out = np.zeros((n, 4), np.float32)
    m = C.c_int()
    _lib.check(_lib.lib().lio_compact_encode(np.ascontiguousarray(tf7, np.float32), c, c.shape[0], s, s.shape[0], f, f.shape[0],
                                             out, n, C.byref(m)), "lio_compact_encode")
    return out[:m.value]


def compact_decode(compact):
    """Returns (tf7, corner, surf, full); raises LioError where the reference's handler logs an error and returns."""
    d = _c4(compact)
    sz = np.zeros(3, np.int32)
    _lib.check(_lib.lib().lio_compact_sizes(d, d.shape[0], sz), "lio_compact_sizes")
    tf7 = np.zeros(7, np.float32)
    outs = [np.zeros((max(int(k), 1), 4), np.float32) for k in sz]
    _lib.check(_lib.lib().lio_compact_decode(d, d.shape[0], tf7, *outs), "lio_compact_decode")
    return tf7, outs[0][:sz[0]], outs[1][:sz[1]], outs[2][:sz[2]]


def to_pcl32(cloud) -> np.ndarray:
    """(n, 4) float32 -> (n, 32) uint8 pcl::PointXYZI records (the `data` of the PointCloud2)."""
    c = _c4(cloud)
    out = np.zeros((c.shape[0], 32), np.uint8)
    _lib.check(_lib.lib().lio_xyzi_to_pcl32(c, c.shape[0], out.reshape(-1)), "lio_xyzi_to_pcl32")
    return out


def from_pcl32(data) -> np.ndarray:
    d = np.ascontiguousarray(data, np.uint8).reshape(-1, 32)
    out = np.zeros((d.shape[0], 4), np.float32)
    _lib.check(_lib.lib().lio_pcl32_to_xyzi(d.reshape(-1), d.shape[0], out), "lio_pcl32_to_xyzi")
    return out
