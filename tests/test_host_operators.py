"""Host-side logic of liblio_b200.so that needs no GPU: C-ABI surface, factor operators vs the oracle."""
import ctypes as C
import os
import re

import numpy as np

from lio_mapping_b200 import _lib, estimator, synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_cabi_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "lio_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    names = set(re.findall(r"\b(lio_[a-z0-9_]+)\s*\(", hdr))
    names -= {"lio_allreduce_fn"}
    assert len(names) > 40
    L = C.CDLL(_lib.LIB_PATH)
    missing = [n for n in sorted(names) if not hasattr(L, n)]
    assert not missing, missing


def test_no_cpu_fallback_without_device():
    L = _lib.lib()
    if L.lio_device_count() > 0:
        return
    cfg = _lib.PPConfig()
    L.lio_pp_default_config(C.byref(cfg))
    h = C.c_void_p()
    assert L.lio_pp_create(C.byref(cfg), 1000, 0, None, C.byref(h)) == -4   # LIO_ERR_NO_DEVICE
    ec = _lib.EstConfig()
    L.lio_est_default_config(C.byref(ec))
    assert L.lio_est_create(C.byref(ec), 0, None, C.byref(h)) == -4
    out = np.zeros((4, 4), np.float32)
    n = C.c_int()
    assert L.lio_voxel_grid_host(out, 4, 0.4, out, 4, C.byref(n), 0) == -4


def rand_pose(rng, scale=5.0):
    q = rng.normal(size=4); q /= np.linalg.norm(q)
    return np.concatenate([rng.uniform(-scale, scale, 3), q])


def test_ppp_operator_matches_oracle(oracle):
    rng = np.random.default_rng(0)
    for _ in range(20):
        x0, xi, xe = rand_pose(rng), rand_pose(rng), rand_pose(rng, 0.5)
        p = rng.uniform(-20, 20, 3); coeff = rng.normal(size=4)
        r, J = estimator.ppp_evaluate(p, coeff, x0, xi, xe)
        ro, Jo = oracle.ppp_evaluate(p, coeff, x0, xi, xe)
        assert abs(r - ro) <= 1e-12 * max(1.0, abs(ro))
        for a, b in zip(J, Jo):
            assert np.abs(a - b).max() <= 1e-12 * max(1.0, np.abs(b).max())


def test_preintegration_and_imu_factor_match_oracle(oracle):
    rng = np.random.default_rng(1)
    traj = synth.Trajectory(ax=3.0, ay=2.0, az=0.1, period=20.0)
    tt, acc, gyr = synth.make_imu(traj, 1.0, 1.1)
    _, _, _, g0, a0 = traj.state(np.array(1.0))
    ba, bg = np.array([0.01, -0.02, 0.03]), np.array([0.001, 0.002, -0.001])
    po = oracle.Pim(a0, g0, ba, bg, acc_n=0.2, gyr_n=0.02)
    pg = estimator.Pim(a0, g0, ba, bg, acc_n=0.2, gyr_n=0.02)
    last = 1.0
    for j in range(len(tt)):
        po.push_back(tt[j] - last, acc[j], gyr[j]); pg.push_back(tt[j] - last, acc[j], gyr[j]); last = tt[j]
    so, sg = po.get(), pg.get()
    for k in ["delta_p", "delta_q", "delta_v", "sum_dt", "jacobian", "covariance"]:
        assert np.allclose(sg[k], so[k], rtol=1e-12, atol=1e-15), k
    pi, pj = rand_pose(rng, 2.0), rand_pose(rng, 2.0)
    sbi, sbj = rng.normal(0, 0.1, 9), rng.normal(0, 0.1, 9)
    ro, Jo = po.imu_factor(pi, sbi, pj, sbj)
    rg, Jg = pg.imu_factor(pi, sbi, pj, sbj)
    assert np.allclose(rg, ro, rtol=1e-9, atol=1e-9 * np.abs(ro).max())
    for a, b in zip(Jg, Jo):
        assert np.allclose(a, b, rtol=1e-9, atol=1e-9 * np.abs(b).max())


def test_header_is_plain_c99_and_c_client_links(tmp_path):
    """include/lio_b200.h must be consumable by a C compiler (extern "C" boundary, plain pointers and sizes) and a C
    client must link against the shared library; without a device the client stops after the host-only entry points."""
    import os
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = tmp_path / "minimal"
    cmd = ["gcc", "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-I" + os.path.join(root, "include"),
           os.path.join(root, "examples", "minimal.c"), "-L" + os.path.join(root, "lio_mapping_b200"), "-llio_b200",
           "-Wl,-rpath," + os.path.join(root, "lio_mapping_b200"), "-lm", "-o", str(exe)]
    out = subprocess.run(cmd, capture_output=True, text=True)
    assert out.returncode == 0, out.stderr
    run = subprocess.run([str(exe)], capture_output=True, text=True, timeout=120)
    assert run.returncode == 0, (run.stdout, run.stderr)
    assert "liblio_b200 version" in run.stdout


def _build_shim(tmp_path):
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = tmp_path / "estimator_shim"
    cmd = ["g++", "-std=c++14", "-Wall", "-Wextra", "-Werror", "-I" + os.path.join(root, "include"),
           os.path.join(root, "examples", "estimator_shim.cc"), "-L" + os.path.join(root, "lio_mapping_b200"), "-llio_b200",
           "-Wl,-rpath," + os.path.join(root, "lio_mapping_b200"), "-o", str(exe)]
    out = subprocess.run(cmd, capture_output=True, text=True)
    assert out.returncode == 0, out.stderr
    return exe


def test_cxx_estimator_shim_compiles_and_links(tmp_path):
    """The C++ shim that keeps the reference's Estimator member names / control flow (examples/estimator_shim.cc) compiles
    with -Wall -Wextra -Werror against the C header and links against the library; without a device it reports that the
    estimator cannot be created (there is no CPU fallback) and exits cleanly."""
    import subprocess
    exe = _build_shim(tmp_path)
    if _lib.lib().lio_device_count() > 0:
        return
    run = subprocess.run([str(exe)], capture_output=True, text=True, timeout=120)
    assert run.returncode == 0, (run.stdout, run.stderr)
    assert "no CPU fallback" in run.stdout
