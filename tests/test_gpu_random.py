"""Randomised GPU-vs-oracle parity over shapes the fixed tests do not enumerate: random dims (incl.
non-multiples of 4 -> the 8- and 4-byte vector paths), dtypes, bag/batch sizes (incl. 1 and ragged tails),
Simple/Split storage, all strategies, int32/int64 indices, strided destinations and cotangents, skewed
index distributions.  Fixed seeds; every comparison is bit-exact (strict order) or within the north
star's 1e-5 (split order, long buckets only)."""
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
    oracle.build()
    return oracle


def rand_base(rng, dim, nrows, dtype):
    if np.issubdtype(dtype, np.integer):
        info = np.iinfo(dtype)
        return rng.integers(info.min // 4, info.max // 4, (dim, nrows)).astype(dtype)
    return (rng.standard_normal((dim, nrows)) * rng.choice([1e-3, 1.0, 1e3])).astype(dtype)


def rand_indices(rng, nrows, shape, skew):
    if skew:
        w = 1.0 / np.arange(1, nrows + 1) ** 1.2
        cdf = np.cumsum(w) / w.sum()
        return (np.searchsorted(cdf, rng.random(shape)) + 1).astype(np.int64)
    return rng.integers(1, nrows + 1, shape)


@pytest.mark.parametrize("seed", range(24))
def test_random_lookup_and_maplookup(E, O, seed):
    rng = np.random.default_rng(1000 + seed)
    nt = int(rng.integers(1, 6))
    dtype = rng.choice([np.float32, np.float32, np.float64, np.int32, np.int64])
    same_dim = rng.random() < 0.5
    dims = [int(rng.choice([1, 2, 3, 4, 6, 8, 12, 16, 20, 32, 48, 64, 100, 128, 192, 256, 520, 1030]))] * nt if same_dim \
        else [int(rng.choice([4, 8, 16, 20, 64, 128, 256])) for _ in range(nt)]
    nrows = [int(rng.integers(1, 700)) for _ in range(nt)]
    batch = int(rng.choice([1, 2, 31, 32, 33, 100, 257, 1000]))
    bag = int(rng.choice([0, 1, 2, 3, 4, 5, 8, 31, 32, 33, 70]))
    skew = rng.random() < 0.4
    idx32 = rng.random() < 0.3
    base = [rand_base(rng, d, n, dtype) for d, n in zip(dims, nrows)]
    tables, refs = [], []
    for b in base:
        kind = rng.choice(["dyn", "static", "split"])
        if kind == "split":
            shard = int(rng.integers(1, b.shape[1] + 1))
            tables.append(E.SplitEmbedding(b, shard)); refs.append(O.Table(b, cols_per_shard=shard))
        else:
            tables.append(E.SimpleEmbedding(b, E.Static(b.shape[0]) if kind == "static" else None))
            refs.append(O.Table(b, static=kind == "static"))
    I = [rand_indices(rng, n, (bag, batch) if bag else (batch,), skew) for n in nrows]
    Id = [E.as_device_indices(i.astype(np.int32) if idx32 else i) for i in I]
    want = [O.lookup(r, i) for r, i in zip(refs, I)]
    for t, i, w in zip(tables, Id, want):
        assert np.array_equal(E.lookup(t, i).numpy(), w)
    out = E.maplookup(E.SimpleParallelStrategy(), tables, Id)
    assert all(np.array_equal(o.numpy(), w) for o, w in zip(out, want))
    prepend = int(rng.choice([0, 1, 5, 16, 128]))
    dst = E.DeviceArray.from_numpy(np.full((prepend + sum(dims), batch), 3, dtype))
    E.maplookup_(E.PreallocationStrategy(prepend), dst, tables, Id)
    got = dst.numpy()
    assert np.all(got[:prepend] == 3) and np.array_equal(got[prepend:], np.concatenate(want, axis=0))


@pytest.mark.parametrize("seed", range(24))
@pytest.mark.parametrize("order", ["strict", "split"])
def test_random_update(E, O, seed, order):
    E.set_update_order(order)
    try:
        rng = np.random.default_rng(5000 + seed)
        nt = int(rng.integers(1, 5))
        dtype = rng.choice([np.float32, np.float32, np.float64])
        dims = [int(rng.choice([1, 3, 4, 8, 16, 20, 32, 64, 80, 128, 256, 520]))] * nt if rng.random() < 0.6 \
            else [int(rng.choice([4, 16, 64, 128])) for _ in range(nt)]
        nrows = [int(rng.integers(1, 600)) for _ in range(nt)]
        batch = int(rng.choice([1, 7, 64, 300, 1500]))
        bag = int(rng.choice([0, 1, 2, 5, 32]))
        skew = rng.random() < 0.5
        eta = float(rng.choice([0.01, 0.5, 10.0]))
        base = [rand_base(rng, d, n, dtype) for d, n in zip(dims, nrows)]
        statics = [bool(rng.random() < 0.5) for _ in base]
        tables = [E.SimpleEmbedding(b.copy(), E.Static(b.shape[0]) if st else None) for b, st in zip(base, statics)]
        refs = [O.Table(b.copy(order="F"), static=st) for b, st in zip(base, statics)]
        I = [rand_indices(rng, n, (bag, batch) if bag else (batch,), skew) for n in nrows]
        prepend = int(rng.choice([0, 4, 16]))
        Delta = np.asfortranarray(rng.standard_normal((prepend + sum(dims), batch)).astype(dtype))
        Dd = E.DeviceArray.from_numpy(Delta)
        grads, deltas, off = [], [], prepend
        for t, d, i in zip(tables, dims, I):
            grads.append(E.SparseEmbeddingUpdate(t.lookup_type, Dd.rows(off, off + d), i))   # strided views
            deltas.append(Delta[off:off + d, :])
            off += d
        if nt == 1 and rng.random() < 0.5:
            E.update_(E.Descent(eta), tables[0], grads[0])
        else:
            E.update_(E.Descent(eta), tables, grads, [E.Indexer()])
        for t, r, d, i, n in zip(tables, refs, deltas, I, nrows):
            O.update(r, d, i, eta)
            got = t.to_numpy()
            if order == "strict":
                assert np.array_equal(got, r.data)
            else:
                counts = np.bincount(np.asarray(i).ravel(), minlength=n + 1)[1:]
                short = counts <= 128
                assert np.array_equal(got[:, short], r.data[:, short])
                if (~short).any():
                    assert np.linalg.norm(got - r.data) <= 1e-5 * np.linalg.norm(r.data)
    finally:
        E.set_update_order("strict")
