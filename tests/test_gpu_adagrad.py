"""GPU parity of the row-wise Adagrad extension (etb_adagrad_update).

The reference's update! exists for Flux.Descent only, so `oracle.adagrad_update` defines the semantics
(include/embtab_b200.h) including the kernel's fixed summation order; the kernels are compared bit for bit."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def E():
    import embtab
    return embtab


@pytest.fixture(scope="module")
def O():
    import oracle
    return oracle


def _bits(a):
    a = np.ascontiguousarray(a)
    return a.view({2: np.uint16, 4: np.uint32, 8: np.uint64}[a.dtype.itemsize])


def _rand(rng, shape, dt):
    return np.asfortranarray(rng.standard_normal(shape).astype(np.float32).astype(dt))


def _dtypes(E):
    return {"f32": np.dtype(np.float32), "f64": np.dtype(np.float64), "f16": np.dtype(np.float16), "bf16": E.bfloat16}


@pytest.mark.parametrize("order", ["strict", "split"])
@pytest.mark.parametrize("name", ["f32", "f64", "f16", "bf16"])
@pytest.mark.parametrize("dim", [128, 64, 16, 4, 80, 6, 256, 512])
def test_adagrad_steps_match_oracle(E, O, name, dim, order):
    dt = _dtypes(E)[name]
    if dim * dt.itemsize > 2048:
        pytest.skip("row longer than one pass")
    E.set_update_order(order)
    try:
        rng = np.random.default_rng(300 + dim)
        base = _rand(rng, (dim, 300), dt)
        for static in (True, False):
            table = E.SimpleEmbedding(base.copy(order="F"), E.Static(dim) if static else E.Dynamic())
            opt = E.Adagrad(0.05, 1e-6)
            want = base.copy(order="F")
            state = np.zeros(300, np.float64 if name == "f64" else np.float32)
            for step, shape in enumerate(((200,), (5, 180), (3, 400))):      # three steps: the state accumulates
                I = rng.integers(1, 301, shape)
                delta = _rand(rng, (dim, shape[-1]), dt)
                E.update_(opt, table, E.SparseEmbeddingUpdate(table.lookup_type, delta, I))
                O.adagrad_update(want, state, delta, I, 0.05, 1e-6)
                assert np.array_equal(_bits(table.to_numpy()), _bits(want)), (name, dim, static, step)
                assert np.array_equal(_bits(opt.state(table).numpy()), _bits(state)), (name, dim, static, step)
    finally:
        E.set_update_order("strict")


def test_adagrad_ensemble_and_hot_rows(E, O):
    # an ensemble (one launch), medium buckets (5..128 members) and a hot row (> 128 members): strict order is
    # bit-identical; the chunked order differs only by Float32 association of g
    rng = np.random.default_rng(17)
    bases = [_rand(rng, (128, 200), np.float32) for _ in range(3)]
    I = rng.integers(1, 201, (8, 600, 3))
    I[:, :, 1][rng.random((8, 600)) < 0.3] = 7
    Id = E.as_device_indices(I)
    deltas = [_rand(rng, (128, 600), np.float32) for _ in range(3)]
    for order in ("strict", "split"):
        E.set_update_order(order)
        tables = [E.SimpleEmbedding(b.copy(order="F"), E.Static(128)) for b in bases]
        opt = E.Adagrad(0.1, 1e-8)
        grads = [E.SparseEmbeddingUpdate(E.Static(128), d, i) for d, i in zip(deltas, E.colwrap(Id))]
        E.update_(opt, tables, grads, [E.Indexer()])
        for k in range(3):
            want, state = bases[k].copy(order="F"), np.zeros(200, np.float32)
            O.adagrad_update(want, state, deltas[k], I[:, :, k], 0.1, 1e-8)
            got, gstate = tables[k].to_numpy(), opt.state(tables[k]).numpy()
            if order == "strict":
                assert np.array_equal(_bits(got), _bits(want)) and np.array_equal(_bits(gstate), _bits(state))
            else:
                assert np.allclose(got, want, rtol=1e-5, atol=1e-6) and np.allclose(gstate, state, rtol=1e-5)
    E.set_update_order("strict")


def test_adagrad_rejects_long_rows(E):
    table = E.SimpleEmbedding(np.zeros((1024, 10), np.float32))
    g = E.SparseEmbeddingUpdate(E.Dynamic(), np.zeros((1024, 4), np.float32), np.array([1, 2, 3, 4]))
    with pytest.raises(E.EmbTabError):
        E.update_(E.Adagrad(0.1), table, g)
