"""GPU parity: pullback -> SparseEmbeddingUpdate -> index! -> update!(Descent) -- the reference's
test/update.jl, test/misc.jl (Indexer known answer) and the gradient half of test/map.jl.

Tolerances: the strictly sequential kernel keeps the reference's association and epilogue, so
results are compared bit-for-bit with the oracle; the north star's 1e-5 relative tolerance is
asserted explicitly where a different association is allowed."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu
RTOL = 1e-5  # BASELINE.json north_star: Float32 pooled sums and SGD updates


@pytest.fixture(scope="module")
def E():
    import embtab
    return embtab


@pytest.fixture(autouse=True, params=["strict", "split"])
def order(request, E):
    """Every test runs in both reduction orders.  They differ only for buckets with more than 128
    members; every reference-shaped test stays below that, so `==` holds in both."""
    E.set_update_order(request.param)
    yield request.param
    E.set_update_order("strict")   # the default


@pytest.fixture(scope="module")
def O():
    import oracle
    oracle.build()
    return oracle


def test_indexer_known_answer(E, O, golden):
    # reference test/misc.jl:74-110, compared per bucket (bucket ORDER is first-seen in the
    # reference, ascending row here; members keep occurrence order)
    g = golden["index"]
    table = E.SimpleEmbedding(np.zeros((16, g["maxindex"]), np.float32))
    delta = E.DeviceArray.zeros((16, len(g["A"])))
    grad = E.SparseEmbeddingUpdate(E.Dynamic(), delta, np.array(g["A"]))
    ix = E.Indexer()
    for _ in range(2):
        E.index_(ix, table, grad)
        got = {row: m for (slot, row), m in ix.buckets().items()}
        want = O.buckets([tuple(c) for c in g["cumulative"]], np.array(g["map"]))
        assert got == want
    # matrix traversal order (test/misc.jl:9-10)
    y = np.array([[1, 2], [3, 4]])
    grad = E.SparseEmbeddingUpdate(E.Dynamic(), E.DeviceArray.zeros((16, 2)), y)
    E.index_(ix, table, grad)
    assert {r: m for (_, r), m in ix.buckets().items()} == {1: [1], 3: [1], 2: [2], 4: [2]}


def test_readme_update(E, golden):
    g = golden["readme_update"]
    A = E.SimpleEmbedding(np.zeros(g["table_shape"], np.float32))
    y, back = E.pullback(E.lookup, A, g["inds"])
    assert np.all(y.numpy() == 0)
    gradient = back(np.array(g["adjoint_rows"], np.float32))[1]
    assert isinstance(gradient, E.SparseEmbeddingUpdate)
    assert E.update_(E.Descent(g["eta"]), A, gradient) is None
    expect = np.array(g["expect_rows"], np.float32)
    got = A.to_numpy()
    assert np.all(np.abs(got - expect) <= np.spacing(np.abs(expect)))  # see test_oracle_golden


def _update_inner(E, O, shape_of, rows, static, numtests=3):
    # reference test/update.jl:4-84
    rng = np.random.default_rng(rows + int(static))
    ncols = 100
    base = rng.standard_normal((rows, ncols)).astype(np.float32)
    table = E.SimpleEmbedding(base.copy(), E.Static(rows) if static else None)
    opt = E.Descent(10.0)
    for _ in range(numtests):
        indices = rng.integers(1, ncols + 1, shape_of(ncols))
        out, back = E.pullback(E.lookup, table, indices)
        assert np.array_equal(out.numpy(), O.lookup(O.Table(base, static=static), indices))
        diff_out = rng.standard_normal(out.shape).astype(np.float32)
        res = back(diff_out)
        assert len(res) == 3 and res[0] is None and res[2] is None
        grad = res[1]
        assert isinstance(grad, E.SparseEmbeddingUpdate)
        assert np.array_equal(grad.indices.numpy(), indices)
        # uncompress == dense gradient (test/update.jl:44-45) -- same association: bit-equal
        assert np.array_equal(E.uncompress(grad, ncols).numpy(), O.uncompress(diff_out, indices, ncols))
        # update!(Descent(10), zeros(table), grad) (test/update.jl:55-61 and :73-82)
        for nontemporal in (True, False):
            zt = table.zeros()
            assert isinstance(zt, E.AbstractEmbeddingTable)
            E.update_(opt, zt, grad, E.Indexer(), nontemporal)
            zo = O.Table(np.zeros_like(base, order="F"), static=static)
            O.update(zo, diff_out, indices, 10.0)
            assert np.array_equal(zt.to_numpy(), zo.data)
            dense = -10.0 * O.uncompress(diff_out, indices, ncols).astype(np.float64)
            assert np.allclose(zt.to_numpy(), dense, rtol=RTOL, atol=1e-5)


@pytest.mark.parametrize("rows", [64, 80, 256])
@pytest.mark.parametrize("static", [True, False])
def test_update_nonreducing(E, O, rows, static):
    _update_inner(E, O, lambda n: n, rows, static)


@pytest.mark.parametrize("rows", [64, 80, 256])
@pytest.mark.parametrize("static", [True, False])
def test_update_reducing(E, O, rows, static):
    _update_inner(E, O, lambda n: (10, n), rows, static)


def test_update_partitions(E, O):
    # reference test/update.jl:90-120: full update == 4 IndexerView partial updates
    rng = np.random.default_rng(5)
    base = rng.standard_normal((16, 100)).astype(np.float32)
    delta = rng.standard_normal((16, 512)).astype(np.float32)
    inds = rng.integers(1, 101, 512)
    A = E.SimpleEmbedding(base.copy(), E.Static(16))
    grad = E.SparseEmbeddingUpdate(E.Static(16), delta, inds)
    indexer = E.Indexer()
    E.index_(indexer, A, grad)
    E.update_table_(A, grad, indexer, np.float32(1.0))
    B = E.SimpleEmbedding(base.copy(), E.Static(16))
    E.index_(indexer, B, grad)
    for s in range(1, 5):
        E.update_table_(B, grad, E.IndexerView(indexer, 4, s), np.float32(1.0))
    assert A == B
    ref = O.Table(base.copy(order="F"), static=True)
    O.update(ref, delta, inds, 1.0)
    assert np.array_equal(A.to_numpy(), ref.data)


@pytest.mark.parametrize("dtype,dim", [(np.float64, 24), (np.float32, 5), (np.float32, 1504), (np.float64, 130),
                                       # rows of 3 * 2^k / 5 * 2^k vectors: the exact-fit layouts (4..32 lanes x 3 or 5 vectors)
                                       (np.float32, 12), (np.float32, 48), (np.float32, 96), (np.float32, 192), (np.float32, 384),
                                       (np.float32, 20), (np.float32, 80), (np.float32, 160), (np.float32, 320), (np.float32, 640),
                                       (np.float64, 40), (np.float16, 80)])
def test_update_other_shapes(E, O, dtype, dim):
    if dtype == np.float16:
        pytest.skip("half-precision shapes are covered by test_gpu_halfprec.py")
    rng = np.random.default_rng(dim)
    base = rng.standard_normal((dim, 60)).astype(dtype)
    inds = rng.integers(1, 61, (7, 90))
    delta = rng.standard_normal((dim, 90)).astype(dtype)
    t = E.SimpleEmbedding(base.copy())
    E.update_(E.Descent(0.25), t, E.SparseEmbeddingUpdate(E.Dynamic(), delta, inds))
    ref = O.Table(base.copy(order="F"))
    O.update(ref, delta, inds, 0.25)
    assert np.array_equal(t.to_numpy(), ref.data)


def test_update_heavy_duplicates_and_empty(E, O, order):
    # one row takes most of the occurrences (Zipf-like hot row): long sequential bucket
    rng = np.random.default_rng(6)
    base = rng.standard_normal((64, 1000)).astype(np.float32)
    inds = rng.integers(1, 1001, 5000)
    inds[rng.random(5000) < 0.6] = 17
    delta = rng.standard_normal((64, 5000)).astype(np.float32)
    t = E.SimpleEmbedding(base.copy(), E.Static(64))
    E.update_(E.Descent(0.01), t, E.SparseEmbeddingUpdate(E.Static(64), delta, inds))
    ref = O.Table(base.copy(order="F"), static=True)
    O.update(ref, delta, inds, 0.01)
    got = t.to_numpy()
    if order == "strict":
        assert np.array_equal(got, ref.data)
    else:  # row 17 (~3000 members) is summed as 128-member chunks: same value within 1e-5
        assert np.array_equal(np.delete(got, 16, axis=1), np.delete(ref.data, 16, axis=1))
        assert not np.array_equal(got[:, 16], base[:, 16])
        assert np.linalg.norm(got[:, 16] - ref.data[:, 16]) <= RTOL * np.linalg.norm(ref.data[:, 16])
        assert np.allclose(got[:, 16], ref.data[:, 16], rtol=1e-4, atol=1e-6)
        t3 = E.SimpleEmbedding(base.copy(), E.Static(64))      # deterministic: same bits every run
        E.update_(E.Descent(0.01), t3, E.SparseEmbeddingUpdate(E.Static(64), delta, inds))
        assert np.array_equal(t3.to_numpy(), got)
    # empty update: nothing changes
    t2 = E.SimpleEmbedding(base.copy(), E.Static(64))
    E.update_(E.Descent(0.01), t2, E.SparseEmbeddingUpdate(E.Static(64), np.zeros((64, 0), np.float32), np.zeros(0, np.int64)))
    assert t2 == base


def test_update_split_table(E, O):
    # the reference never updates a SplitEmbedding (no zeros/pointer); C4 needs it
    rng = np.random.default_rng(8)
    base = rng.standard_normal((128, 500)).astype(np.float32)
    inds = rng.integers(1, 501, (32, 64))
    delta = rng.standard_normal((128, 64)).astype(np.float32)
    t = E.SplitEmbedding(base.copy(), 130)
    E.update_(E.Descent(0.1), t, E.SparseEmbeddingUpdate(t.lookup_type, delta, inds))
    ref = O.Table(base.copy(order="F"), cols_per_shard=130)
    O.update(ref, delta, inds, 0.1)
    assert np.array_equal(t.to_numpy(), ref.dense())


def test_map_gradients(E, O):
    # reference test/map.jl:118-177
    rng = np.random.default_rng(10)
    dims = [(5, 5), (5, 10), (5, 15)]
    batch = 5
    I = [rng.integers(1, d[1] + 1, batch) for d in dims]
    tables = [E.SimpleEmbedding(rng.random(d).astype(np.float32)) for d in dims]
    y = rng.random((sum(d[0] for d in dims), batch)).astype(np.float32)

    out, back = E.pullback(E.maplookup, tables, I)
    cat = np.concatenate([o.numpy() for o in out], axis=0)
    dl = (2.0 / cat.size) * (cat - y)  # gradient of Flux.mse
    grads = back([dl[5 * k:5 * k + 5] for k in range(3)])[2]
    for i, g in enumerate(grads):
        assert isinstance(g, E.SparseEmbeddingUpdate) and np.array_equal(g.indices.numpy(), I[i])

    out2, back2 = E.pullback(E.maplookup, E.PreallocationStrategy(), tables, I)
    assert np.array_equal(out2.numpy(), cat)
    grads2 = back2(dl)[2]
    for i, g in enumerate(grads2):
        assert np.array_equal(g.indices.numpy(), I[i])
        assert np.array_equal(g.delta.numpy(), grads[i].delta.numpy())   # per-table row slice of the cotangent

    out3, back3 = E.pullback(E.maplookup, E.PreallocationStrategy(20), tables, I)
    assert np.array_equal(out3.numpy()[20:], cat)
    full = np.concatenate([np.zeros((20, batch), np.float32), dl], axis=0)
    grads3 = back3(full)[2]
    for i, g in enumerate(grads3):
        assert np.array_equal(g.delta.numpy(), grads[i].delta.numpy())


@pytest.mark.parametrize("static", [True, False])
def test_ensemble_update_matches_oracle(E, O, static):
    # ensemble update! (reference src/sparseupdate.jl:199-238) through the Preallocation pullback
    rng = np.random.default_rng(12)
    nt, dim, nrows, bag, batch, prepend = 7, 32, 300, 6, 128, 16
    base = [rng.standard_normal((dim, nrows)).astype(np.float32) for _ in range(nt)]
    S = E.Static(dim) if static else None
    tables = [E.SimpleEmbedding(b.copy(), S) for b in base]
    I = rng.integers(1, nrows + 1, (bag, batch, nt))
    strategy = E.PreallocationStrategy(prepend)
    out, back = E.pullback(E.maplookup, strategy, tables, I)
    Delta = rng.standard_normal(out.shape).astype(np.float32)
    grads = back(Delta)[2]
    called = []
    E.update_(E.Descent(0.05), tables, grads, [E.Indexer() for _ in tables], telemetry_cb=lambda: called.append(1))
    assert called == [1]
    refs = [O.Table(b.copy(order="F"), static=static) for b in base]
    DeltaF = np.asfortranarray(Delta)
    deltas = [DeltaF[prepend + dim * k: prepend + dim * (k + 1), :] for k in range(nt)]
    O.update_ensemble(refs, deltas, [I[:, :, k] for k in range(nt)], 0.05, nthreads=2)
    for t, r in zip(tables, refs):
        assert np.array_equal(t.to_numpy(), r.data)


def test_update_rejects_integer_tables(E):
    # the reference throws too (InexactError at convert(eltype(table), eta))
    t = E.SimpleEmbedding(np.zeros((4, 4), np.int64))
    g = E.SparseEmbeddingUpdate(E.Dynamic(), np.zeros((4, 2), np.int64), [1, 2])
    with pytest.raises(E.EmbTabError):
        E.update_(E.Descent(0.1), t, g)


@pytest.mark.parametrize("dim,dtype", [(128, np.float32), (16, np.float32), (40, np.float64), (520, np.float32),
                                       # the feature-sliced strict kernel: partial last slice, 4-byte pieces, multi-pass rows
                                       (80, np.float32), (5, np.float32), (33, np.float32), (1504, np.float32), (130, np.float64)])
def test_split_long_zipf(E, O, dim, dtype, order):
    # Zipf(1.05)-like duplicates over several tables: many long buckets of different lengths,
    # through the ensemble path; several group widths / vector counts
    rng = np.random.default_rng(dim)
    nt, nrows, bag, batch = 3, 5000, 8, 4096
    w = 1.0 / np.arange(1, nrows + 1) ** 1.05
    cdf = np.cumsum(w) / w.sum()
    base = [rng.standard_normal((dim, nrows)).astype(dtype) for _ in range(nt)]
    I = [(np.searchsorted(cdf, rng.random((bag, batch))) + 1).astype(np.int64) for _ in range(nt)]
    deltas = [rng.standard_normal((dim, batch)).astype(dtype) for _ in range(nt)]
    tables = [E.SimpleEmbedding(b.copy()) for b in base]
    grads = [E.SparseEmbeddingUpdate(E.Dynamic(), d, i) for d, i in zip(deltas, I)]
    E.update_(E.Descent(0.01), tables, grads, [E.Indexer()])
    for t, b, d, i in zip(tables, base, deltas, I):
        ref = O.Table(b.copy(order="F"))
        O.update(ref, d, i, 0.01)
        got = t.to_numpy()
        if order == "strict":
            assert np.array_equal(got, ref.data)
        else:
            counts = np.bincount(i.ravel(), minlength=nrows + 1)[1:]
            short = counts <= 128
            assert short.sum() > 0 and (~short).sum() > 0
            assert np.array_equal(got[:, short], ref.data[:, short])
            assert np.linalg.norm(got - ref.data) <= RTOL * np.linalg.norm(ref.data)


def test_prefetch_index_overlapped(E, O):
    # index! started early on a side stream (GPU-only extension) gives the same update, single and ensemble
    rng = np.random.default_rng(21)
    base = [rng.standard_normal((64, 400)).astype(np.float32) for _ in range(3)]
    I = rng.integers(1, 401, (6, 300, 3))
    deltas = [rng.standard_normal((64, 300)).astype(np.float32) for _ in range(3)]
    tables = [E.SimpleEmbedding(b.copy(), E.Static(64)) for b in base]
    ix = E.Indexer()
    Id = E.as_device_indices(I)
    E.prefetch_index(ix, tables, Id)
    out = E.maplookup(E.PreallocationStrategy(), tables, Id)          # overlaps the side stream
    grads = [E.SparseEmbeddingUpdate(E.Static(64), d, i) for d, i in zip(deltas, E.colwrap(Id))]
    assert ix._prefetched is not None
    E.update_(E.Descent(0.3), tables, grads, [ix])
    assert ix._prefetched is None                                      # consumed
    for t, b, d, k in zip(tables, base, deltas, range(3)):
        ref = O.Table(b.copy(order="F"), static=True)
        O.update(ref, d, I[:, :, k], 0.3)
        assert np.array_equal(t.to_numpy(), ref.data)
    # a prefetch for OTHER indices is ignored (falls back to a fresh index!)
    t1 = E.SimpleEmbedding(base[0].copy(), E.Static(64))
    E.prefetch_index(ix, t1, E.as_device_indices(I[:, :, 1]))
    g = E.SparseEmbeddingUpdate(E.Static(64), deltas[0], I[:, :, 0])
    E.update_(E.Descent(0.3), t1, g, ix)
    ref = O.Table(base[0].copy(order="F"), static=True)
    O.update(ref, deltas[0], I[:, :, 0], 0.3)
    assert np.array_equal(t1.to_numpy(), ref.data)


def test_mixed_static_dynamic_ensemble(E, O):
    # the FMA-vs-two-roundings epilogue is a per-table choice (reference @generated dispatch): an ensemble
    # may mix Static (fma) and Dynamic (two roundings) tables in one call
    rng = np.random.default_rng(31)
    base = [rng.standard_normal((64, 200)).astype(np.float32) for _ in range(4)]
    statics = [True, False, True, False]
    tables = [E.SimpleEmbedding(b.copy(), E.Static(64) if st else None) for b, st in zip(base, statics)]
    I = rng.integers(1, 201, (5, 150, 4))
    deltas = [rng.standard_normal((64, 150)).astype(np.float32) for _ in range(4)]
    grads = [E.SparseEmbeddingUpdate(t.lookup_type, d, I[:, :, k]) for k, (t, d) in enumerate(zip(tables, deltas))]
    E.update_(E.Descent(0.7), tables, grads, [E.Indexer()])
    differ = 0
    for k, (t, b, d, st) in enumerate(zip(tables, base, deltas, statics)):
        ref = O.Table(b.copy(order="F"), static=st)
        O.update(ref, d, I[:, :, k], 0.7)
        assert np.array_equal(t.to_numpy(), ref.data)
        other = O.Table(b.copy(order="F"), static=not st)
        O.update(other, d, I[:, :, k], 0.7)
        differ += int(not np.array_equal(other.data, ref.data))
    assert differ > 0   # the two epilogues really round differently, so the test can tell them apart


def test_cuda_graph_replay_of_a_step(E, O):
    # the whole step (fused lookup + index! + update!) captured once and replayed: same result as eager
    rng = np.random.default_rng(41)
    base = [rng.standard_normal((64, 300)).astype(np.float32) for _ in range(3)]
    I = E.as_device_indices(rng.integers(1, 301, (4, 128, 3)))
    Ih = I.numpy()
    delta = E.DeviceArray.from_numpy(rng.standard_normal((3 * 64, 128)).astype(np.float32))
    tables = [E.SimpleEmbedding(b.copy(), E.Static(64)) for b in base]
    out = E.DeviceArray.zeros((3 * 64, 128))
    ix, opt = E.Indexer(), E.Descent(0.25)
    grads = [E.SparseEmbeddingUpdate(E.Static(64), delta.rows(64 * k, 64 * k + 64), i) for k, i in enumerate(E.colwrap(I))]

    def step():
        E.maplookup_(E.PreallocationStrategy(), out, tables, I)
        E.update_(opt, tables, grads, [ix])

    replay = E.capture(step, warmup=1)      # 1 warm-up + 1 captured (not executed) = 1 update applied so far
    replay()
    replay()                                # 3 updates in total
    refs = [O.Table(b.copy(order="F"), static=True) for b in base]
    dh = delta.numpy()
    for _ in range(3):
        for k, r in enumerate(refs):
            O.update(r, np.asfortranarray(dh[64 * k:64 * k + 64]), Ih[:, :, k], 0.25)
    for t, r in zip(tables, refs):
        assert np.array_equal(t.to_numpy(), r.data)
    # the last replay's lookup saw the tables after 2 updates
    refs2 = [O.Table(b.copy(order="F"), static=True) for b in base]
    for _ in range(2):
        for k, r in enumerate(refs2):
            O.update(r, np.asfortranarray(dh[64 * k:64 * k + 64]), Ih[:, :, k], 0.25)
    want = np.concatenate([O.lookup(r, Ih[:, :, k]) for k, r in enumerate(refs2)], axis=0)
    assert np.array_equal(out.numpy(), want)


def test_many_table_ensemble_update(E, O):
    # more tables than one update launch holds (kUMaxItems = 96): slots are chunked across launches
    rng = np.random.default_rng(51)
    nt = 130
    base = [rng.standard_normal((16, 40)).astype(np.float32) for _ in range(nt)]
    tables = [E.SimpleEmbedding(b.copy(), E.Static(16)) for b in base]
    I = rng.integers(1, 41, (3, 50, nt))
    deltas = [rng.standard_normal((16, 50)).astype(np.float32) for _ in range(nt)]
    grads = [E.SparseEmbeddingUpdate(E.Static(16), d, I[:, :, k]) for k, d in enumerate(deltas)]
    E.update_(E.Descent(0.2), tables, grads, [E.Indexer()])
    for k, (t, b, d) in enumerate(zip(tables, base, deltas)):
        ref = O.Table(b.copy(order="F"), static=True)
        O.update(ref, d, I[:, :, k], 0.2)
        assert np.array_equal(t.to_numpy(), ref.data), k


def test_wide_keys_path(E, O):
    # tables that DECLARE more than 2^32 (table, row) combinations use 64-bit sort keys; the indices only
    # ever touch the first rows, so the declared size needs no memory
    from embtab import _lib

    class Huge(E.SimpleEmbedding):
        def descriptor(self):
            d = super().descriptor()
            d.nrows = 1 << 33
            return d

    rng = np.random.default_rng(52)
    base = [rng.standard_normal((32, 500)).astype(np.float32) for _ in range(3)]
    tables = [Huge(b.copy(), E.Static(32)) for b in base]
    I = rng.integers(1, 501, (4, 200, 3))
    deltas = [rng.standard_normal((32, 200)).astype(np.float32) for _ in range(3)]
    grads = [E.SparseEmbeddingUpdate(E.Static(32), d, I[:, :, k]) for k, d in enumerate(deltas)]
    ix = E.Indexer()
    E.update_(E.Descent(0.2), tables, grads, [ix])
    assert ix.view.key_bytes == 8
    for k, (t, b, d) in enumerate(zip(tables, base, deltas)):
        ref = O.Table(b.copy(order="F"), static=True)
        O.update(ref, d, I[:, :, k], 0.2)
        assert np.array_equal(t.to_numpy(), ref.data)


@pytest.mark.parametrize("n,nrows,wide", [(1_000_003, 5000, False), (300_000, 70_000_000, False), (200_000, 9000, True), (77, 50, False)])
def test_index_sort_is_a_stable_sort(E, n, nrows, wide):
    # the hand-written radix sort behind index!: sorted by row, and STABLE -- members of a bucket keep the
    # occurrence order (what remap! records, reference src/utils.jl:242-272); checked against numpy's stable sort
    from embtab.sparseupdate import _IndicesOnly, _peek

    class Declared(E.SimpleEmbedding):          # only the declared row count matters to index!
        def descriptor(self):
            d = super().descriptor()
            d.nrows = (1 << 34) if wide else nrows
            return d

    rng = np.random.default_rng(n)
    table = Declared(np.zeros((4, 8), np.float32))
    I = rng.integers(1, nrows + 1, n)
    ix = E.Indexer()
    E.index_(ix, table, _IndicesOnly(table, I))   # index! needs the indices only
    v = ix.view
    keys = _peek(v.keys, n, np.uint32 if v.key_bytes == 4 else np.uint64).astype(np.int64)
    mp = _peek(v.map, n, np.int32)
    order = np.argsort(I, kind="stable")
    assert v.key_bytes == (8 if wide else 4)
    assert np.array_equal(keys, I[order] - 1)
    assert np.array_equal(mp, order.astype(np.int32))
    nnz = int(_peek(v.nnz, 1, np.int64)[0])
    assert nnz == np.unique(I).size


def test_strided_dynamic_table(E, O):
    # a Dynamic SimpleEmbedding over a row-slice VIEW of a bigger matrix (leading dimension > featuresize):
    # the reference's Dynamic columnpointer honours strides(A)[2] (src/simple.jl:52, src/EmbeddingTables.jl:83-85)
    rng = np.random.default_rng(61)
    big_h = rng.standard_normal((100, 300)).astype(np.float32)
    big = E.DeviceArray.from_numpy(big_h)
    for dim in (64, 20):                       # 16-byte and 8-byte aligned slices
        view = big.rows(8, 8 + dim)
        table = E.SimpleEmbedding(view)        # Dynamic
        base = big_h[8:8 + dim].copy()
        I = rng.integers(1, 301, (5, 200))
        assert np.array_equal(E.lookup(table, I).numpy(), O.lookup(O.Table(base), I))
        delta = rng.standard_normal((dim, 200)).astype(np.float32)
        before = big.numpy().copy()
        E.update_(E.Descent(0.1), table, E.SparseEmbeddingUpdate(E.Dynamic(), delta, I))
        ref = O.Table(base.copy(order="F"))
        O.update(ref, delta, I, 0.1)
        after = big.numpy()
        assert np.array_equal(after[8:8 + dim], ref.data)
        assert np.array_equal(after[:8], before[:8]) and np.array_equal(after[8 + dim:], before[8 + dim:])  # neighbours untouched
        big_h = after


def test_update_applies_the_cotangent_passed_to_update(E, O):
    # update!(table, update, indexer, alpha) reads update.delta (reference src/sparseupdate.jl:131-154):
    # an indexer filled for one update may be reused with ANOTHER cotangent over the same indices
    rng = np.random.default_rng(71)
    base = rng.standard_normal((32, 200)).astype(np.float32)
    I = rng.integers(1, 201, (4, 150))
    Id = E.as_device_indices(I)
    d1 = rng.standard_normal((32, 150)).astype(np.float32)
    d2 = rng.standard_normal((32, 150)).astype(np.float32)
    table = E.SimpleEmbedding(base.copy(), E.Static(32))
    ix = E.Indexer()
    E.index_(ix, table, E.SparseEmbeddingUpdate(E.Static(32), d1, Id))
    E.update_table_(table, E.SparseEmbeddingUpdate(E.Static(32), d2, Id), ix, 0.25)
    ref = O.Table(base.copy(order="F"), static=True)
    O.update(ref, d2, I, 0.25)
    assert np.array_equal(table.to_numpy(), ref.data)


def test_host_mapped_result_and_cotangent(E, O):
    # page-locked host buffers are device-addressable: the forward may store its result straight into one
    # and update! may read its cotangent straight from one (no staging copy); results are unchanged
    import torch
    rng = np.random.default_rng(72)
    base = [rng.standard_normal((64, 500)).astype(np.float32) for _ in range(2)]
    I = rng.integers(1, 501, (8, 256, 2))
    Id = E.as_device_indices(I)
    tables = [E.SimpleEmbedding(b.copy(), E.Static(64)) for b in base]
    out_h = E.pinned_empty((16 + 128, 256), np.float32)
    out_h[...] = -1.0
    out_m = E.DeviceArray.mapped(out_h)
    E.maplookup_(E.PreallocationStrategy(16), out_m, tables, Id)
    torch.cuda.synchronize()
    for k in range(2):
        assert np.array_equal(out_h[16 + 64 * k:16 + 64 * (k + 1)], O.lookup(O.Table(base[k], static=True), I[:, :, k]))
    assert np.all(out_h[:16] == -1.0)                          # the prepended rows are the caller's
    delta_h = E.pinned_empty((16 + 128, 256), np.float32)
    delta_h[...] = rng.standard_normal(delta_h.shape).astype(np.float32)
    delta_m = E.DeviceArray.mapped(delta_h)
    slicer = E.Slicer(17, 1, delta_m)
    grads = [E.SparseEmbeddingUpdate(E.Static(64), slicer(64), i) for i in E.colwrap(Id)]
    E.update_(E.Descent(0.5), tables, grads, [E.Indexer()])
    for k in range(2):
        ref = O.Table(base[k].copy(order="F"), static=True)
        O.update(ref, np.asfortranarray(delta_h[16 + 64 * k:16 + 64 * (k + 1)]), I[:, :, k], 0.5)
        assert np.array_equal(tables[k].to_numpy(), ref.data)


def test_heterogeneous_split_ensemble(E, O, order):
    # chunked tables of very different row counts, mixed with plain tables and two feature sizes (two kernel classes), bucket
    # count not a multiple of 32, and hot rows (long buckets on chunked tables): a lane that does not own a bucket
    # must never resolve a row through another table's chunk-pointer array
    rng = np.random.default_rng(77)
    spec = [("split", 64, 1000, 130), ("split", 64, 50, 7), ("simple", 16, 300, 0), ("split", 16, 2000, 512),
            ("split", 64, 9, 2), ("simple", 64, 4000, 0)]
    batch, bag = 333, 3
    base = [rng.standard_normal((dim, nrows)).astype(np.float32) for _, dim, nrows, _ in spec]
    tables = [E.SplitEmbedding(b.copy(), chunk) if kind == "split" else E.SimpleEmbedding(b.copy())
              for b, (kind, _, _, chunk) in zip(base, spec)]
    I = [rng.integers(1, nrows + 1, (bag, batch)) for _, _, nrows, _ in spec]
    I[0][rng.random(I[0].shape) < 0.5] = 999          # a hot row in the last chunk of the first table
    I[4][:] = rng.integers(8, 10, I[4].shape)          # two rows take everything
    deltas = [rng.standard_normal((dim, batch)).astype(np.float32) for _, dim, _, _ in spec]
    grads = [E.SparseEmbeddingUpdate(t.lookup_type, d, i) for t, d, i in zip(tables, deltas, I)]
    E.update_(E.Descent(0.05), tables, grads, [E.Indexer()])
    for t, b, d, i, (kind, _, _, chunk) in zip(tables, base, deltas, I, spec):
        # SplitEmbedding(A, cols_per_shard) is Static{featuresize} like the reference's (src/split.jl:9-27): FMA epilogue
        ref = O.Table(b.copy(order="F"), static=True, cols_per_shard=chunk) if kind == "split" else O.Table(b.copy(order="F"))
        O.update(ref, d, i, 0.05)
        got, want = t.to_numpy(), (ref.dense() if kind == "split" else ref.data)
        if order == "strict":
            assert np.array_equal(got, want)
        else:
            assert np.linalg.norm(got - want) <= RTOL * np.linalg.norm(want)
