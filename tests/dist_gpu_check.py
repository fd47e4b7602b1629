"""Run under torchrun (or with RANK/WORLD_SIZE=0/1): sharded forward/backward/update vs the CPU ORACLE and vs the
single-GPU path recomputed locally on every rank, bit for bit -- NCCL all-to-all, fused NVLink peer stores with
the NCCL barrier, and fused with the peer-memory flag barrier and the table-group pipelined backward."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "embeddingtables.jl_b200"), os.path.join(ROOT, "oracle")):
    sys.path.insert(0, p)
import embtab as E
import oracle as O          # the checker (tests/ may use it)
from embtab.dist import ShardedEnsemble, ShardPlan

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
rng = np.random.default_rng(7)                      # same data on every rank
dims = [128, 128, 64, 128, 16, 128, 32, 128, 128, 64]   # >= 8 tables so that every rank of an 8-GPU box owns one
nrows, bag, batch, prepend = 777, 5, 203, 16
base = [rng.standard_normal((d, nrows)).astype(np.float32) for d in dims]
I = [rng.integers(1, nrows + 1, (bag, batch)) for _ in dims]
total = prepend + sum(dims)
delta = rng.standard_normal((total, batch)).astype(np.float32)

def make_table(t):              # C4 uses SplitEmbedding tables: every other table is chunked (ragged last chunk)
    return E.SplitEmbedding(base[t].copy(), 200) if t % 2 else E.SimpleEmbedding(base[t].copy())


# single-GPU reference on this rank (same table types: the update epilogue is chosen per table type)
ref_tables = [make_table(t) for t in range(len(base))]
ref_out, back = E.pullback(E.maplookup, E.PreallocationStrategy(prepend), ref_tables, I)
E.update_(E.Descent(0.1), ref_tables, back(delta)[2], [E.Indexer()])

# ... and the oracle: the single-GPU path equals it, so the sharded path (compared with both below) does too
orc = [O.Table(base[t].copy(order="F"), static=bool(t % 2), cols_per_shard=200 if t % 2 else None) for t in range(len(base))]
orc_out = O.maplookup("preallocation", orc, I, prependrows=prepend, out=np.zeros((total, batch), np.float32, order="F"))
off, deltas = prepend, []
for d in dims:
    deltas.append(np.asfortranarray(delta[off:off + d]))
    off += d
O.update_ensemble(orc, deltas, I, 0.1)
assert np.array_equal(ref_out.numpy()[prepend:], orc_out[prepend:]), "single-GPU forward differs from the oracle"
for t in range(len(base)):
    assert np.array_equal(ref_tables[t].to_numpy(), orc[t].dense()), f"single-GPU update of table {t} differs from the oracle"

plan = ShardPlan(dims, world, rank, prepend, batch)
mine = list(plan.my_tables)
# NCCL all-to-all + pack/unpack; NVLink peer stores + NCCL barrier; peer stores + flag barrier + pipelined backward
# ...; the same with the copy engines carrying the blocks
# ...; an ensemble built with table groups but driven whole (grouped=False: one index!, one exchange, one update!)
for fused, groups, peer_barrier, copy_engine, grouped in ((False, 1, False, False, None), (True, 1, False, False, None),
                                                          (True, 3, True, False, None), (True, 1, True, True, None),
                                                          (True, 3, True, False, False)):
    ens = ShardedEnsemble([make_table(t) for t in mine], plan, fused=fused, table_groups=groups, peer_barrier=peer_barrier,
                          copy_engine=copy_engine)
    ens.out.fill(-5.0)
    torch.cuda.synchronize()
    dist.barrier()
    for rep in range(2):        # twice: buffers are reused across steps
        out = ens.forward([I[t] for t in mine], grouped=grouped)
        got = out.numpy()
        want = orc_out[:, plan.clo[rank]:plan.chi[rank]]
        assert np.array_equal(got[prepend:], want[prepend:]), f"sharded forward differs from the oracle (fused={fused}, copy_engine={copy_engine})"
        assert np.all(got[:prepend] == -5.0), "prepend rows were touched"
        d_local = E.DeviceArray.from_numpy(delta[:, plan.clo[rank]:plan.chi[rank]])
        if rep == 0:
            grads = ens.backward(d_local)
        elif groups > 1:        # update once, after the second (buffer-reusing) round trip: group by group (or whole)
            ens.backward_update_(E.Descent(0.1), d_local, grouped=grouped)
        else:
            ens.update_(E.Descent(0.1), ens.backward(d_local))
    for t, tab in zip(mine, ens.tables):
        assert np.array_equal(tab.to_numpy(), orc[t].dense()), f"table {t} differs from the oracle after update (fused={fused}, groups={groups})"
        assert np.array_equal(tab.to_numpy(), ref_tables[t].to_numpy()), f"table {t} differs after update (fused={fused})"
    torch.cuda.synchronize()
    dist.barrier()
    ens.close()
# data-parallel input: every rank starts with ITS samples' indices for all tables
ens = ShardedEnsemble([E.SimpleEmbedding(base[t].copy()) for t in mine], plan)
I_all = np.stack(I, axis=2)                                             # (bag, batch, T)
for wire in (None, np.int32):
    mine_I = ens.distribute_indices(E.DeviceArray.from_numpy(I_all[:, plan.clo[rank]:plan.chi[rank], :]), wire_dtype=wire)
    for k, t in enumerate(mine):
        assert np.array_equal(mine_I[k].numpy(), I[t]), f"index distribution differs (wire={wire})"
    got = ens.forward(mine_I).numpy()
    assert np.array_equal(got[prepend:], ref_out.numpy()[prepend:, plan.clo[rank]:plan.chi[rank]])
torch.cuda.synchronize()
dist.barrier()
if rank == 0:
    print("dist check ok", world, "(sharded == single GPU == oracle, bit for bit: NCCL, fused + NCCL barrier, fused + peer-flag barrier + grouped backward, copy engines, grouped ensemble driven whole)")
dist.destroy_process_group()
