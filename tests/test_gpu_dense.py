"""Dense BLAS-1 kernels vs the oracle (elementwise ops bit-identical; reductions within
the reference's own device-vs-reference tolerance 10*eps*sqrt-free bound)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def mk(gko, exec_, a, stride=None):
    return gko.matrix.Dense.from_numpy(exec_, a, stride=stride)


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
@pytest.mark.parametrize("shape,stride", [((1000, 1), None), ((257, 3), 5), ((64, 33), None), ((0, 2), None)])
@pytest.mark.parametrize("percol", [False, True])
def test_elementwise_bit_identical(gko, exec_, ora, dtype, shape, stride, percol):
    rng = np.random.default_rng(0)
    n, k = shape
    x, y = rng.standard_normal(shape).astype(dtype), rng.standard_normal(shape).astype(dtype)
    alpha = rng.standard_normal((1, k if percol else 1)).astype(dtype)
    V = "f64" if dtype == np.float64 else "f32"
    for op in ("scale", "inv_scale", "add_scaled", "sub_scaled"):
        want = y.copy()
        dy = mk(gko, exec_, y, stride)
        if op in ("scale", "inv_scale"):
            getattr(ora.lib(), f"oracle_dense_{op}_{V}")(ora.i64(n), ora.i64(k), ora.P(alpha), ora.i64(alpha.shape[1]),
                                                         ora.P(want), ora.i64(k))
            getattr(dy, op)(mk(gko, exec_, alpha))
        else:
            getattr(ora.lib(), f"oracle_dense_{op}_{V}")(ora.i64(n), ora.i64(k), ora.P(alpha), ora.i64(alpha.shape[1]),
                                                         ora.P(x), ora.i64(k), ora.P(want), ora.i64(k))
            getattr(dy, op)(mk(gko, exec_, alpha), mk(gko, exec_, x, stride))
        assert np.array_equal(dy.to_numpy(), want), op


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
@pytest.mark.parametrize("shape,stride", [((100000, 1), None), ((4097, 3), 4), ((513, 40), None), ((1, 1), None),
                                          ((0, 3), None)])
def test_reductions(gko, exec_, ora, dtype, shape, stride):
    rng = np.random.default_rng(1)
    n, k = shape
    x, y = rng.standard_normal(shape).astype(dtype), rng.standard_normal(shape).astype(dtype)
    V = "f64" if dtype == np.float64 else "f32"
    dx, dy = mk(gko, exec_, x, stride), mk(gko, exec_, y, stride)
    res = gko.matrix.Dense.create(exec_, (1, k), dx.t.dtype)
    want = np.zeros(k, dtype=dtype)
    eps = np.finfo(dtype).eps
    # dot: error relative to sum |x_i y_i|
    dx.compute_dot(dy, res)
    getattr(ora.lib(), f"oracle_dense_compute_dot_{V}")(ora.i64(n), ora.i64(k), ora.P(x), ora.i64(k), ora.P(y),
                                                        ora.i64(k), ora.P(want))
    scale = np.abs(x.astype(np.float64) * y).sum(axis=0) + 1e-300
    assert np.all(np.abs(res.to_numpy()[0] - want) <= 4 * eps * np.sqrt(max(n, 1)) * scale)
    for name in ("compute_norm2", "compute_norm1"):
        getattr(dx, name)(res)
        getattr(ora.lib(), f"oracle_dense_{name}_{V}")(ora.i64(n), ora.i64(k), ora.P(x), ora.i64(k), ora.P(want))
        assert np.allclose(res.to_numpy()[0], want, rtol=4 * eps * np.sqrt(max(n, 1)), atol=0), name
    # reductions are single-pass and run-to-run bit-reproducible
    dx.compute_norm2(res)
    first = res.to_numpy().copy()
    for _ in range(3):
        dx.compute_norm2(res)
        assert np.array_equal(res.to_numpy(), first)


def test_row_gather_fill_copy(gko, exec_):
    rng = np.random.default_rng(2)
    src = rng.standard_normal((50, 3))
    import torch
    for idt in (torch.int32, torch.int64):
        rows = torch.tensor([4, 4, 0, 49, 17], dtype=idt, device=exec_.device)
        out = gko.matrix.Dense.create(exec_, (5, 3))
        mk(gko, exec_, src, 4).row_gather(rows, out)
        assert np.array_equal(out.to_numpy(), src[[4, 4, 0, 49, 17]])
    d = gko.matrix.Dense.create(exec_, (7, 2), stride=3)
    d.fill(2.5)
    assert np.all(d.to_numpy() == 2.5)
    e = gko.matrix.Dense.create(exec_, (7, 2))
    e.copy_from(d)
    assert np.array_equal(e.to_numpy(), d.to_numpy())
