"""Host-tier tables behind an HBM row cache (CachedEmbedding): wherever a row lives -- page-locked host
memory or its HBM slot -- lookups and update! give the bits of the all-HBM SimpleEmbedding and of the oracle.  The
Update-phase hook (IndexingContext Update -> etb_cache_admit) fills the cache with the rows a batch touched often."""
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


def zipf(rng, nrows, n, alpha=1.05):
    """Zipf(alpha) ranks; rank -> row through ONE fixed permutation, so that the hot rows stay hot from batch to batch"""
    w = 1.0 / np.arange(1, nrows + 1, dtype=np.float64) ** alpha
    cdf = np.cumsum(w)
    cdf /= cdf[-1]
    return np.random.default_rng(nrows).permutation(nrows)[np.searchsorted(cdf, rng.random(n))] + 1


@pytest.mark.parametrize("dim,static", [(128, True), (64, False), (20, False)])
def test_cached_table_equals_hbm_table_over_several_steps(E, O, dim, static):
    rng = np.random.default_rng(dim)
    nrows, bag, batch, cache_rows = 5000, 8, 512, 300
    base = rng.standard_normal((dim, nrows)).astype(np.float32)
    S = E.Static(dim) if static else E.Dynamic()
    cached = E.CachedEmbedding(base, cache_rows, S, min_count=2)
    plain = E.SimpleEmbedding(base.copy(), S if static else None)
    ref = O.Table(base.copy(order="F"), static=static)
    opt = E.Descent(0.05)
    hit = []
    for step in range(4):
        I = zipf(rng, nrows, bag * batch).reshape((bag, batch), order="F")
        hit.append(cached.hit_rate(I))
        out_c, back_c = E.pullback(E.lookup, cached, I)
        assert isinstance(cached.context, E.Forward)               # the lookup asked for the Forward descriptor
        out_p, back_p = E.pullback(E.lookup, plain, I)
        want = O.lookup(ref, I)
        assert np.array_equal(out_c.numpy(), want) and np.array_equal(out_p.numpy(), want)
        delta = rng.standard_normal((dim, batch)).astype(np.float32)
        E.update_(opt, cached, back_c(delta)[1])
        assert isinstance(cached.context, E.Update)                # ... and update! for the Update one
        E.update_(opt, plain, back_p(delta)[1])
        O.update(ref, delta, I, 0.05)
        g = E.lookup(cached, np.arange(1, nrows + 1)).numpy()     # every row, through the cache rule
        assert np.array_equal(g, ref.data), f"step {step}: cached table differs from the oracle"
    assert np.array_equal(cached.to_numpy(), ref.data)            # after the flush the HOST table holds everything
    assert np.array_equal(plain.to_numpy(), ref.data)
    assert 0 < cached.cached_rows() <= cache_rows
    assert hit[0] == 0.0 and hit[-1] > 0.3                        # hot Zipf rows were admitted by the Update-phase hook


def test_cached_tables_in_an_ensemble_with_hbm_tables(E, O):
    rng = np.random.default_rng(9)
    dims, nrows, bag, batch, prepend = [64, 64, 64], 2000, 4, 256, 16
    base = [rng.standard_normal((d, nrows)).astype(np.float32) for d in dims]
    tables = [E.CachedEmbedding(base[0], 100, E.Static(64)), E.SimpleEmbedding(base[1].copy(), E.Static(64)),
              E.CachedEmbedding(base[2], 0, E.Static(64))]       # the last one: no cache at all, every row on the host
    refs = [O.Table(b.copy(order="F"), static=True) for b in base]
    for step in range(3):
        I = np.stack([zipf(rng, nrows, bag * batch).reshape((bag, batch), order="F") for _ in dims], axis=2)
        out, back = E.pullback(E.maplookup, E.PreallocationStrategy(prepend), tables, I)
        want = O.maplookup("preallocation", refs, I, prependrows=prepend,
                           out=np.zeros((prepend + sum(dims), batch), np.float32, order="F"))
        assert np.array_equal(out.numpy()[prepend:], want[prepend:])
        delta = np.asfortranarray(rng.standard_normal(out.shape).astype(np.float32))
        E.update_(E.Descent(0.1), tables, back(delta)[2], [E.Indexer()])
        O.update_ensemble(refs, [delta[prepend + 64 * k: prepend + 64 * (k + 1)] for k in range(3)], [I[:, :, k] for k in range(3)], 0.1)
    for t, r in zip(tables, refs):
        assert np.array_equal(t.to_numpy(), r.data)
    assert tables[0].cached_rows() > 0 and tables[2].cached_rows() == 0
