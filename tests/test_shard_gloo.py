"""Multi-GPU path, host-side logic on CPU (world_size 2, gloo): frames dealt to ranks by lio_est_frame_owner,
per-frame 7x7 S blocks summed by one allreduce == the unsharded reduction."""
import os

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from lio_mapping_b200 import _lib


def s_blocks(features, R, t):
    """numpy statement of the kernel's per-frame reduction: S = sum rho'(r^2) [g;r][g;r]^T, cost = sum log(1+r^2)."""
    p, c = features
    a = c[:, :3] @ R                     # a = R^T w
    g = np.concatenate([a, np.cross(p[:, :3], a)], 1)
    r = (a * (p[:, :3] + t)).sum(1) + c[:, 3]
    u = np.concatenate([g, r[:, None]], 1)
    w = 1.0 / (1.0 + r * r)
    S = (u * w[:, None]).T @ u
    return S, np.log1p(r * r).sum()


def _worker(rank, world, port, feats, Rs, ts, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    L = _lib.lib()
    O = len(feats)
    buf = torch.zeros(O, 32, dtype=torch.float64)
    for i in range(1, O + 1):
        if L.lio_est_frame_owner(i, world) != rank:
            continue
        S, c = s_blocks(feats[i - 1], Rs[i - 1], ts[i - 1])
        buf[i - 1, :28] = torch.from_numpy(S[np.triu_indices(7)])
        buf[i - 1, 28] = c
    dist.all_reduce(buf)                 # the one exchange step of the path
    if rank == 0:
        out.put(buf.numpy().copy())
    dist.destroy_process_group()


def test_frame_owner_partition():
    L = _lib.lib()
    for world in [1, 2, 4, 8]:
        owners = [L.lio_est_frame_owner(i, world) for i in range(1, 11)]
        assert set(owners) == set(range(min(world, 10)))
        counts = np.bincount(owners, minlength=world)
        assert counts.max() - counts[counts > 0].min() <= 1      # balanced round-robin
    assert L.lio_est_frame_owner(0, 2) == -1


def test_sharded_allreduce_equals_unsharded():
    rng = np.random.default_rng(0)
    O = 5
    feats, Rs, ts = [], [], []
    for _ in range(O):
        n = int(rng.integers(200, 400))
        p = rng.uniform(-20, 20, (n, 4)); w = rng.normal(size=(n, 3)); w /= np.linalg.norm(w, axis=1, keepdims=True)
        c = np.concatenate([w * rng.uniform(0.2, 1, (n, 1)), rng.normal(0, 0.1, (n, 1))], 1)
        q = rng.normal(size=4); q /= np.linalg.norm(q)
        x, y, z, ww = q
        R = np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - z * ww), 2 * (x * z + y * ww)],
                      [2 * (x * y + z * ww), 1 - 2 * (x * x + z * z), 2 * (y * z - x * ww)],
                      [2 * (x * z - y * ww), 2 * (y * z + x * ww), 1 - 2 * (x * x + y * y)]])
        feats.append((p, c)); Rs.append(R); ts.append(rng.normal(size=3))
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, feats, Rs, ts, out)) for r in range(2)]
    for p in procs:
        p.start()
    got = out.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for i in range(O):
        S, c = s_blocks(feats[i], Rs[i], ts[i])
        assert np.allclose(got[i, :28], S[np.triu_indices(7)], rtol=1e-13, atol=1e-12)
        assert abs(got[i, 28] - c) <= 1e-12 * max(1.0, c)
