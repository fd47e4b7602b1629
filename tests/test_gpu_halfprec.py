"""GPU parity of the half-precision extension (ETB_F16 / ETB_BF16 tables, Float32 arithmetic).

The reference has no such tables, so there is no reference behaviour to match: the oracle's `lookup_lowp` /
`update_lowp` define the semantics (Float32 accumulation in the reference's order, one rounding to the storage
type) and the kernels are compared with them bit for bit."""
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


def _dtypes(E):
    return [np.dtype(np.float16), E.bfloat16]


def _bits(a):
    return np.ascontiguousarray(a).view(np.uint16)


def _rand(rng, shape, dtype):
    return np.asfortranarray(rng.standard_normal(shape).astype(np.float32).astype(dtype))


@pytest.mark.parametrize("which", [0, 1])
@pytest.mark.parametrize("dim", [128, 64, 16, 8, 6, 250, 520, 1024])
def test_pooled_and_gather(E, O, which, dim):
    dt = _dtypes(E)[which]
    rng = np.random.default_rng(100 + dim)
    base = _rand(rng, (dim, 700), dt)
    for static in (True, False):
        table = E.SimpleEmbedding(base, E.Static(dim) if static else E.Dynamic())
        for bag in (1, 2, 4, 7, 32, 40):
            I = rng.integers(1, 701, (bag, 333))
            got = E.lookup(table, I).numpy()
            assert got.dtype == dt
            assert np.array_equal(_bits(got), _bits(O.lookup_lowp(base, I))), (dim, bag)
        Iv = rng.integers(1, 701, 500)
        assert np.array_equal(_bits(E.lookup(table, Iv).numpy()), _bits(base[:, Iv - 1]))


@pytest.mark.parametrize("which", [0, 1])
def test_float32_accumulation_is_visible(E, O, which):
    # 1 + 2^-9 is exactly representable in neither half type's sum chain: accumulating in the storage type would
    # lose the small terms one by one, Float32 accumulation keeps them until the single final rounding
    dt = _dtypes(E)[which]
    base = np.zeros((8, 40), np.float32)
    base[:, 0] = 1.0
    base[:, 1:] = 2.0 ** -9
    base = np.asfortranarray(base.astype(dt))
    I = np.arange(1, 41).reshape(40, 1)
    got = E.lookup(E.SimpleEmbedding(base), I).numpy()
    want = np.float32(1.0 + 39 * 2.0 ** -9).astype(dt)
    assert np.all(got.astype(np.float32) == np.float32(want)) and np.float32(want) > 1.0
    assert np.array_equal(_bits(got), _bits(O.lookup_lowp(base, I)))


@pytest.mark.parametrize("which", [0, 1])
def test_maplookup_preallocation_and_split_tables(E, O, which):
    dt = _dtypes(E)[which]
    rng = np.random.default_rng(7)
    bases = [_rand(rng, (64, 300), dt) for _ in range(3)]
    tables = [E.SimpleEmbedding(bases[0], E.Static(64)), E.SplitEmbedding(bases[1], 37), E.SimpleEmbedding(bases[2], E.Static(64))]
    I = rng.integers(1, 301, (5, 200, 3))
    out = E.maplookup(E.PreallocationStrategy(32), tables, I).numpy()
    assert out.shape == (32 + 3 * 64, 200) and out.dtype == dt
    for k in range(3):
        assert np.array_equal(_bits(out[32 + 64 * k:32 + 64 * (k + 1)]), _bits(O.lookup_lowp(bases[k], I[:, :, k])))


@pytest.mark.parametrize("order", ["strict", "split"])
@pytest.mark.parametrize("which", [0, 1])
@pytest.mark.parametrize("dim", [128, 64, 16, 10, 250, 520, 1024])
def test_update(E, O, which, dim, order):
    dt = _dtypes(E)[which]
    E.set_update_order(order)
    try:
        rng = np.random.default_rng(200 + dim)
        base = _rand(rng, (dim, 400), dt)
        for static in (True, False):
            for shape in ((300,), (6, 250)):
                table = E.SimpleEmbedding(base.copy(order="F"), E.Static(dim) if static else E.Dynamic())
                I = rng.integers(1, 401, shape)
                batch = shape[-1]
                delta = _rand(rng, (dim, batch), dt)
                E.update_(E.Descent(0.37), table, E.SparseEmbeddingUpdate(table.lookup_type, delta, I))
                want = O.update_lowp(base.copy(order="F"), delta, I, 0.37)
                assert np.array_equal(_bits(table.to_numpy()), _bits(want)), (dim, static, shape)
    finally:
        E.set_update_order("strict")


@pytest.mark.parametrize("which", [0, 1])
def test_update_hot_rows(E, O, which):
    # a row with thousands of duplicates: strict order is bit-identical to the oracle; the chunked order keeps
    # Float32 partial sums and differs only by Float32 association (well inside one unit of the storage type)
    dt = _dtypes(E)[which]
    rng = np.random.default_rng(9)
    base = _rand(rng, (128, 50), dt)
    I = rng.integers(1, 51, (8, 2000))
    I[rng.random(I.shape) < 0.6] = 3
    delta = _rand(rng, (128, 2000), dt)
    want = O.update_lowp(base.copy(order="F"), delta, I, 0.01)
    for order in ("strict", "split"):
        E.set_update_order(order)
        table = E.SimpleEmbedding(base.copy(order="F"), E.Static(128))
        E.update_(E.Descent(0.01), table, E.SparseEmbeddingUpdate(E.Static(128), delta, I))
        got = table.to_numpy()
        if order == "strict":
            assert np.array_equal(_bits(got), _bits(want))
        else:
            g, w = got.astype(np.float32), want.astype(np.float32)
            eps = 2.0 ** -10 if dt == np.float16 else 2.0 ** -7          # one unit in the last place, relative
            assert np.all(np.abs(g - w) <= eps * np.maximum(np.abs(w), 1e-3))
    E.set_update_order("strict")


@pytest.mark.parametrize("which", [0, 1])
def test_ensemble_update_with_sliced_cotangent(E, O, which):
    dt = _dtypes(E)[which]
    rng = np.random.default_rng(11)
    bases = [_rand(rng, (64, 500), dt) for _ in range(4)]
    tables = [E.SimpleEmbedding(b.copy(order="F"), E.Static(64)) for b in bases]
    I = rng.integers(1, 501, (4, 300, 4))
    Id = E.as_device_indices(I)
    delta = _rand(rng, (16 + 4 * 64, 300), dt)
    dd = E.DeviceArray.from_numpy(delta)
    slicer = E.Slicer(17, 1, dd)
    grads = [E.SparseEmbeddingUpdate(E.Static(64), slicer(64), i) for i in E.colwrap(Id)]
    E.update_(E.Descent(0.1), tables, grads, [E.Indexer()])
    for k in range(4):
        want = O.update_lowp(bases[k].copy(order="F"), delta[16 + 64 * k:16 + 64 * (k + 1)], I[:, :, k], 0.1)
        assert np.array_equal(_bits(tables[k].to_numpy()), _bits(want))


def test_odd_dim_is_rejected(E):
    table = E.SimpleEmbedding(np.zeros((7, 10), np.float16))
    with pytest.raises(E.EmbTabError):
        E.lookup(table, np.ones((2, 3), np.int64))
