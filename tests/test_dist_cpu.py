"""World-size-2 test of the table-wise sharding plan and the all-to-all plumbing on CPU (gloo).
The per-rank compute is done by the CPU oracle and the strided pack/unpack by numpy here (the GPU
kernels for those are covered by the -m gpu tests); what this pins is ShardPlan's ownership /
offset / split arithmetic and exchange() in both directions: sharded result == single-process
result, bit for bit."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DIMS = [16, 8, 32, 16, 4]
NROWS, BAG, PREPEND = 50, 3, 3


def make_problem(batch_global):
    rng = np.random.default_rng(42)
    base = [np.asfortranarray(rng.standard_normal((d, NROWS)).astype(np.float32)) for d in DIMS]
    I = [rng.integers(1, NROWS + 1, (BAG, batch_global)) for _ in DIMS]
    total = PREPEND + sum(DIMS)
    delta = np.asfortranarray(rng.standard_normal((total, batch_global)).astype(np.float32))
    return base, I, delta


def worker(rank, world, port, batch_global, q):
    for p in (ROOT, os.path.join(ROOT, "embeddingtables.jl_b200"), os.path.join(ROOT, "oracle")):
        sys.path.insert(0, p)
    import oracle as O
    from embtab.dist import ShardPlan, exchange
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        base, I, delta = make_problem(batch_global)
        plan = ShardPlan(DIMS, world, rank, PREPEND, batch_global)
        mine = list(plan.my_tables)
        tabs = [O.Table(base[t].copy(order="F"), static=True) for t in mine]
        # ---- forward: my tables x global batch, written in SEND layout
        send = np.zeros(plan.my_rows * batch_global, np.float32)
        off = 0
        for t, tab in zip(mine, tabs):
            full = O.lookup(tab, I[t])                                  # dim x B_global
            for p in range(world):
                blk = send[plan.send_block_offset(p): plan.send_block_offset(p) + plan.my_rows * plan.cols[p]]
                blk.reshape((plan.my_rows, plan.cols[p]), order="F")[off:off + DIMS[t], :] = full[:, plan.clo[p]:plan.chi[p]]
            off += DIMS[t]
        recv = torch.empty((plan.total_rows - PREPEND) * plan.my_cols)
        exchange(recv, torch.from_numpy(send), plan.fwd_recv_splits(), plan.fwd_send_splits())
        out = np.full((plan.total_rows, plan.my_cols), -1.0, np.float32, order="F")
        r = recv.numpy()
        for qk in range(world):                                         # unpack (== etb_a2a_unpack)
            blk = r[plan.recv_block_offset(qk): plan.recv_block_offset(qk) + plan.rows[qk] * plan.my_cols]
            out[plan.row_off[qk]: plan.row_off[qk] + plan.rows[qk], :] = blk.reshape((plan.rows[qk], plan.my_cols), order="F")
        # ---- backward: pack my cotangent by owner (== etb_a2a_pack), reverse all-to-all
        dl = delta[:, plan.clo[rank]:plan.chi[rank]]
        send2 = np.concatenate([np.asfortranarray(dl[plan.row_off[qk]: plan.row_off[qk] + plan.rows[qk], :]).reshape(-1, order="F")
                                for qk in range(world)])
        recv2 = torch.empty(plan.my_rows * batch_global)
        exchange(recv2, torch.from_numpy(np.ascontiguousarray(send2)), plan.bwd_recv_splits(), plan.bwd_send_splits())
        dglob = recv2.numpy().reshape((plan.my_rows, batch_global), order="F")   # no unpack needed
        off = 0
        for t, tab in zip(mine, tabs):
            O.update(tab, np.asfortranarray(dglob[off:off + DIMS[t], :]), I[t], 0.1)
            off += DIMS[t]
        q.put((rank, out, {t: tab.data.copy() for t, tab in zip(mine, tabs)}, (plan.clo[rank], plan.chi[rank])))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("batch_global", [10, 11])
def test_sharded_equals_single_process(batch_global):
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import oracle as O
    O.build()
    world, port = 2, 29600 + batch_global
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=worker, args=(r, world, port, batch_global, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    base, I, delta = make_problem(batch_global)
    refs = [O.Table(b.copy(order="F"), static=True) for b in base]
    want = O.maplookup("preallocation", refs, I, prependrows=PREPEND,
                       out=np.full((PREPEND + sum(DIMS), batch_global), -1.0, np.float32, order="F"))
    off = PREPEND
    deltas = []
    for d in DIMS:
        deltas.append(delta[off:off + d, :])
        off += d
    O.update_ensemble(refs, deltas, I, 0.1)
    seen = set()
    for rank, out, tabs, (c0, c1) in results:
        assert np.array_equal(out, want[:, c0:c1])          # incl. the untouched prepend rows
        for t, data in tabs.items():
            assert np.array_equal(data, refs[t].data)
            seen.add(t)
    assert seen == set(range(len(DIMS)))


def test_plan_arithmetic():
    sys.path.insert(0, os.path.join(ROOT, "embeddingtables.jl_b200"))
    from embtab.dist import ShardPlan
    for world in (1, 2, 3, 4, 8):
        dims = [128] * 26
        plans = [ShardPlan(dims, world, r, 128, 16384) for r in range(world)]
        assert sum(len(p.my_tables) for p in plans) == 26
        assert sum(p.my_cols for p in plans) == 16384
        assert plans[0].total_rows == 128 + 26 * 128
        for p in plans:
            assert sum(p.fwd_send_splits()) == p.my_rows * 16384
            assert sum(p.fwd_recv_splits()) == (p.total_rows - 128) * p.my_cols
            assert p.row_off[0] == 128 and p.row_off[-1] + p.rows[-1] == p.total_rows
        # what rank a sends to b is what b expects from a
        for a in plans:
            for b in plans:
                assert a.fwd_send_splits()[b.rank] == b.fwd_recv_splits()[a.rank]
