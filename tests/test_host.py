"""CPU: the C-ABI library loads and exports everything include/podb200.h declares, the
product fails loudly without a GPU, and the host-side logic of the shims."""
import ctypes
import os

import numpy as np
import pytest

from pod_uqnn_b200 import _lib, fields, workloads as wl
from oracle import pod_oracle as po, ensemble as ens


def test_library_exports_every_header_symbol():
    lib = _lib.lib()
    hs = _lib.header_symbols()
    assert len(hs) >= 40
    for name in hs:
        assert hasattr(lib, name), f"{name} declared in include/podb200.h but not exported"
    assert set(hs) == set(_lib.SIGNATURES), set(hs) ^ set(_lib.SIGNATURES)


def test_no_cpu_fallback_without_gpu():
    h = ctypes.c_void_p()
    rc = _lib.lib().pb_create(0, ctypes.byref(h))
    if rc == 0:                                   # running on a GPU box
        _lib.lib().pb_destroy(h)
        pytest.skip("a CUDA device is present")
    assert rc == _lib.PB_ERR_CUDA
    assert b"no CPU fallback" in _lib.lib().pb_last_error(None)
    with pytest.raises(_lib.PodB200Error):
        _lib.Context(0)
    from pod_uqnn_b200.pod import perform_pod
    with pytest.raises(_lib.PodB200Error):
        perform_pod(np.ones((8, 4)), 1e-4, verbose=False)


def test_null_context_is_rejected():
    L = _lib.lib()
    assert L.pb_gram(None, None, 1, 1, 1, None) == _lib.PB_ERR_ARG
    assert L.pb_pod(None, None, 1, 1, 1, 0., 0, 0, None) == _lib.PB_ERR_ARG
    assert L.pb_snapshots(None, 1, 1, 1, None, 1, 2, None, 0, None, 0, 1, None, 1) == _lib.PB_ERR_ARG
    assert L.pb_ensemble_predict(None, 1, 1, None, None, 0, None, None, 1., None, 1, None, None) == _lib.PB_ERR_ARG


def test_field_registry():
    assert fields.resolve("ackley") is fields.u_ackley
    assert fields.resolve(fields.u_burgers).has_t

    def u(X, t, mu):
        return X
    with pytest.raises(NotImplementedError):
        fields.resolve(u)
    u.pb_field = "shekel"
    assert fields.resolve(u) is fields.u_shekel
    with pytest.raises(NotImplementedError):
        fields.u_shekel(None, 0, None)            # no host implementation


def test_workloads_match_reference_mesh():
    m = wl.create_linear_mesh(-5., 5., 7, -5., 5., 4)
    assert np.array_equal(m, po.create_linear_mesh(-5., 5., 7, -5., 5., 4))
    assert np.array_equal(wl.create_linear_mesh(0, 10, 9), po.create_linear_mesh(0, 10, 9))
    w = wl.ackley(8, 6, 5)
    assert w["x_mesh"].shape == (48, 3) and w["mu"].shape == (5, 3) and (np.abs(w["mu"]) <= 1).all()
    U, s, phi = wl.dct_lowrank_host(64, 16, r=8)
    assert np.allclose(np.linalg.svd(U, compute_uv=False)[:8], s, rtol=0, atol=1e-13)
    assert wl.dct_rank(s, 1e-2) == po.truncation_rank(s ** 2, 1e-2, 8)


def test_time_grid_matches_numba_linspace():
    from pod_uqnn_b200.acceleration import time_grid
    assert np.array_equal(time_grid(1., 5., 100), po.linspace_numba_parallel(1., 5., 100))
    assert time_grid(1., 5., 100)[0] == 1. and time_grid(1., 5., 100)[-1] == 5.


def test_pack_members_layout():
    from pod_uqnn_b200.varneuralnetwork import pack_members, VarNeuralNetwork
    layers = [2, 4, 3]
    ms = [ens.init_weights(layers, s, .1) for s in (1, 2)]
    flat = pack_members(ms)
    per = 2 * 4 + 4 + 4 * 6 + 6
    assert flat.shape == (2 * per,)
    assert np.array_equal(flat[:8], ms[0][0].ravel()) and np.array_equal(flat[per + 8: per + 12], ms[1][1])
    net = VarNeuralNetwork(layers, 0.01, 0.001, seed=3)
    assert [w.shape for w in net.weights] == [(2, 4), (4,), (4, 6), (6,)]
    with pytest.raises(ValueError):
        net.set_weights(ms[0][:2])


def test_logger_and_metrics_mirror_reference(capsys):
    """Logger prints at multiples of `frequency` only (logger.py:45-64); re_s is the mean column-wise relative error
    (metrics.py:21-30); normalize_host follows varneuralnetwork.py:75-86."""
    from pod_uqnn_b200.logger import Logger
    from pod_uqnn_b200.metrics import re_s
    from pod_uqnn_b200.varneuralnetwork import normalize_host
    lg = Logger(10, 5)
    lg.set_val_err_fn(lambda: {"RE_val": 0.5, "MPIW_val": 2.})
    lg.log_train_start()
    for e in range(10):
        lg.log_train_epoch(e, 1.5)
    lg.log_train_end(10, 1.25)
    out = capsys.readouterr().out
    assert out.count("#:") == 3 and "RE_val: 5.0000e-01" in out and "Training finished (epoch 10)" in out
    header, table = lg.get_logs()
    assert header == "epoch\tL\tRE_val\tMPIW_val" and table.shape == (3, 4) and list(table[:, 0]) == [0, 5, 10]
    assert Logger(10, 5, silent=True).get_logs() is None
    U = np.array([[1., 2.], [2., 0.]]); Up = np.array([[1., 1.], [2., 0.]])
    assert abs(re_s(U, Up) - 0.5 * (0. + 1. / 2.)) < 1e-15
    assert abs(re_s(U, Up, div_max=True) - 0.25) < 1e-15
    X = np.array([[1., 10.], [3., 30.]])
    assert np.array_equal(normalize_host(X, "meanstd", (X.mean(0), X.std(0))), np.array([[-1., -1.], [1., 1.]]))
    assert np.array_equal(normalize_host(X, "center", (X.min(0), X.max(0))), np.array([[-1., -10.], [1., 10.]]))
    assert normalize_host(X, "meanstd", None) is not None and np.array_equal(normalize_host(X, "meanstd", None), X)


def test_podnnmodel_host_layout_helpers(tmp_path):
    from pod_uqnn_b200.podnnmodel import PodnnModel
    x_mesh = wl.create_linear_mesh(0, 1, 5)
    m = PodnnModel(str(tmp_path), 2, x_mesh, 4)
    rng = np.random.default_rng(0)
    Us = rng.standard_normal((2, 5, 4, 3))
    assert np.array_equal(m.destruct(Us), po.destruct(Us, 4))
    assert np.array_equal(m.restruct(m.destruct(Us)), Us)
    m0 = PodnnModel(str(tmp_path), 2, x_mesh, 0)
    Us0 = rng.standard_normal((2, 5, 7))
    assert np.array_equal(m0.destruct(Us0), po.destruct(Us0, 0))
    assert np.array_equal(m0.restruct(m0.destruct(Us0)), Us0)
    assert os.path.exists(os.path.join(str(tmp_path), "setup_data.pkl"))
    n_v, xm, n_t = PodnnModel.load_setup_data(str(tmp_path))
    assert n_v == 2 and n_t == 0 and np.array_equal(xm, x_mesh)
    with pytest.raises(ValueError):                       # podnnmodel.py:294-295: no ensemble defined yet
        m.train_model(0, None, None, None, None, 1)
    with pytest.raises(ValueError):
        m.predict_v(np.zeros((3, 2)))


def test_null_context_is_rejected_tsqr():
    L = _lib.lib()
    assert L.pb_tsqr_r(None, None, 1, 1, 1, None) == _lib.PB_ERR_ARG
    assert L.pb_pod_from_r(None, None, 1, 0., 0, None) == _lib.PB_ERR_ARG
    assert _lib.POD_MODES["tsqr"] == 3 and _lib.T_QR == 10


def test_u_level_output_tiles():
    """PodnnModel._u_tiles: column tiles cover [0, N) exactly once, each with device buffers of the tile's shape."""
    from pod_uqnn_b200.podnnmodel import PodnnModel

    class FakeArr:
        def __init__(self, *shape):
            self.shape = shape

    class FakeCtx:
        def alloc(self, *shape):
            return FakeArr(*shape)

    m = PodnnModel.__new__(PodnnModel)
    m._ctx, m.n_h = FakeCtx(), 100
    for N, cap in ((75, 8 * 100 * 32), (64, 8 * 100 * 32), (5, 8 * 100 * 1000), (7, 1)):
        m.U_TILE_BYTES = cap
        tiles = list(m._u_tiles(N))
        assert tiles[0][0] == 0 and tiles[-1][1] == N
        assert all(a[1] == b[0] for a, b in zip(tiles, tiles[1:]))
        for j0, j1, Um, Us, t in tiles:
            assert Um.shape == Us.shape == (100, j1 - j0) and 0 < j1 - j0 <= max(1, cap // 800)
        assert [t[4] for t in tiles] == list(range(len(tiles)))


@pytest.mark.parametrize("cv", list(range(1, 17)))
def test_gram_tile_shapes_cover_needed_atoms_once(cv):
    """The special tile kinds of the stream-K Gram kernel: every wanted 8 x 8 atom is computed by exactly one warp,
    no atom twice, every frame inside the 16-group panel, and the weights are the busiest sub-partition's."""
    out = (ctypes.c_int * (3 * 8 * 3))()
    tr, wt = (ctypes.c_int * 3)(), (ctypes.c_int * 3)()
    assert _lib.lib().pb_gram_tile_shapes(cv, out, tr, wt) == 0
    sh = np.array(out[:]).reshape(3, 8, 3)
    for k, kind in enumerate(("edge", "diag", "diag_edge")):
        got = np.zeros((16, 16), int)
        load = np.zeros(4, int)
        for w in range(8):
            r0, c0, nt = (int(x) for x in sh[k, w])
            assert 0 <= nt <= 8 and 0 <= r0 and r0 + nt <= 16 and c0 in (0, 4, 8, 12)
            for t in range(nt):
                for u in range(4):
                    r, c = (c0 + u, r0 + t) if tr[k] else (r0 + t, c0 + u)     # atom of the tile (panel ti, panel tj)
                    got[r, c] += 1
                    load[w % 4] += 1
        want = np.ones((16, 16), int)
        if kind != "edge":
            want = np.triu(want)
        if kind != "diag":
            want[:, cv:] = 0
        assert got.max() <= 1 and np.all(got >= want), (cv, kind)
        # weight: tenths of an atom of the busiest sub-partition + the measured per-stage overhead
        assert 10 * load.max() <= wt[k] <= 10 * load.max() + 60 and wt[k] < 640 + 60
        if kind == "diag":
            assert load.max() == 40                  # 10 blocks of 16 atoms over 4 sub-partitions
        if kind == "edge" and cv == 13:
            assert load.max() == 52 and tr[k] == 1   # n_s = 1000: 13 atom rows x 16 atom columns, transposed
    assert _lib.lib().pb_gram_tile_shapes(0, out, tr, wt) == _lib.PB_ERR_ARG


def test_bench_host_logic_and_no_gpu_failure():
    """bench.py: the algorithmic flop count of SURVEY 8d, the workload description per N, and -- without a GPU -- a loud
    failure of the product arm (no CPU fallback may stand in for the CUDA path)."""
    import importlib.util, subprocess, sys, os
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    spec = importlib.util.spec_from_file_location("bench_mod", os.path.join(root, "bench.py"))
    b = importlib.util.module_from_spec(spec)
    argv = sys.argv
    try:
        sys.argv = ["bench.py"]
        spec.loader.exec_module(b)
    finally:
        sys.argv = argv
    assert b.algo_flops(160000, 1000, 14) == 160000 * 1000 * 1001 + 4 * 160000 * 1000 * 14
    cfg = b.workload_config(8)
    assert cfg["n_h_per_gpu"] == 160000 and "n_h = 1280000" in cfg["workload"] and cfg["parallelism"] == "row-sharded x8"
    assert b.METRIC == "pod_fp64_tflops" and b.MODE == "gram2"
    import torch
    if not torch.cuda.is_available():
        r = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--steps", "1", "--warmup", "1"],
                           capture_output=True, text=True, timeout=300)
        assert r.returncode != 0 and "no CPU fallback" in r.stderr and not r.stdout.strip().startswith("{")


def test_generate_hifi_inputs_and_summary():
    """podnnmodel.py:60-87 (row layout of the large prediction inputs) and varneuralnetwork.py:167-169; host only."""
    import numpy as np
    from pod_uqnn_b200.podnnmodel import PodnnModel
    from pod_uqnn_b200.varneuralnetwork import VarNeuralNetwork
    x_mesh = np.column_stack((np.arange(5), np.linspace(0., 1., 5)))
    m = PodnnModel.__new__(PodnnModel)
    m.has_t, m.n_t = True, 4
    np.random.seed(3)
    X_v = m.generate_hifi_inputs(6, [0.1, 1.], [0.2, 2.], t_min=1., t_max=2.5)
    assert X_v.shape == (24, 3)
    assert np.array_equal(X_v[:4, 0], np.linspace(1., 2.5, 4)) and np.array_equal(X_v[4:8, 0], X_v[:4, 0])
    assert np.all(X_v[:4, 1:] == X_v[0, 1:]) and not np.array_equal(X_v[0, 1:], X_v[4, 1:])
    assert np.all((X_v[:, 1] >= 0.1) & (X_v[:, 1] <= 0.2) & (X_v[:, 2] >= 1.) & (X_v[:, 2] <= 2.))
    m.has_t = False
    assert m.generate_hifi_inputs(6, [0.1, 1.], [0.2, 2.]).shape == (6, 2)
    net = VarNeuralNetwork([3, 8, 8, 5], 1e-3, 1e-3, norm="meanstd")
    text = net.summary()
    assert "Total params: " + str(3 * 8 + 8 + 8 * 8 + 8 + 8 * 10 + 10) in text
    net.set_normalize_bounds(np.array([[0., 2.], [2., 6.]]))
    assert np.allclose(net.normalize(np.array([[1., 4.]])), 0.)
