"""N > 1 host logic on CPU: world_size-2 gloo processes run the key-hash exchange (tidb_b200/parallel.py, same
bookkeeping as the NVLink path) and shard-local joins with the oracle; the union of the shard results must equal the
single-process join bit-exactly, and every row must land on the rank the C partition function names."""
import os
import sys

import numpy as np
import pytest
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import oracle_lib as O
    from nested_loop import columns_sorted
    from tidb_b200 import abi
    from tidb_b200.chunk import Chunk, Column
    from tidb_b200.parallel import (exchange_by_key_host, exchange_segments_host, partition_of_keys_np, recv_bases, region_capacity,
                                    segments_to_dense)
    from tidb_b200.plan import FieldType, JoinPlan
    lib = abi.load_lib()
    INT = FieldType(abi.TYPE_LONGLONG, abi.FLAG_NOT_NULL)
    nb, npr = 5000, 20000
    rng = np.random.default_rng(100 + rank)
    ids = rng.permutation(nb).astype(np.int64) + rank * nb
    bk, bv = ids * np.int64(-7046029254386353131), ids * 7
    pid = rng.integers(0, nb * world, npr).astype(np.int64)
    pk, pv = pid * np.int64(-7046029254386353131), np.arange(npr, dtype=np.int64) + rank * npr

    def all_to_all(pieces):
        gathered = [None] * world
        dist.all_gather_object(gathered, pieces)
        return [gathered[src][rank] for src in range(world)]

    lbk, lbv = exchange_by_key_host(bk, [bk, bv], world, all_to_all)
    lpk, lpv = exchange_by_key_host(pk, [pk, pv], world, all_to_all)
    # every received key belongs here, according to the C function the device kernels share
    assert all(lib.tg_partition_of_key(int(k), world) == rank for k in lbk[:500])
    assert np.all(partition_of_keys_np(lpk, world) == rank)
    # count-matrix bookkeeping used by the NVLink path: bases are the exclusive prefix over source ranks
    cnt = np.bincount(partition_of_keys_np(pk, world), minlength=world)
    mats = [None] * world
    dist.all_gather_object(mats, cnt)
    mat = np.stack(mats)
    base, nrecv = recv_bases(mat, rank)
    assert nrecv == len(lpk) and np.array_equal(base, mat[:rank].sum(axis=0))
    # count-free (segment) exchange, the layout SegmentExchange / tg_partition_exchange_cf use on the device: one fixed-capacity
    # region per sender, fill counts travel separately, nothing is compacted; the valid rows are the same multiset
    cap = region_capacity(npr, world)
    assert cap % 1024 == 0 and cap >= npr / world
    seg_cols, seg_cnt, ovf = exchange_segments_host(pk, [pk, pv], world, rank, cap, all_to_all)
    assert not ovf and int(seg_cnt.sum()) == nrecv and np.array_equal(seg_cnt, mat[:, rank])
    assert all(len(c) == world * cap for c in seg_cols)
    dk, dv = segments_to_dense(seg_cols, seg_cnt, cap)
    assert np.array_equal(np.sort(dv), np.sort(lpv)) and np.array_equal(dk[np.argsort(dv)], lpk[np.argsort(lpv)])
    # a region that is too small drops rows and says so (the device path raises its sticky flag the same way)
    _, small_cnt, small_ovf = exchange_segments_host(pk, [pk, pv], world, rank, 1024, all_to_all)
    assert small_ovf and int(small_cnt.max()) == 1024
    # skew with a spill area (tg_partition_exchange_cf_spill / MailboxExchange.drain_spill): half of the rows carry one hot key,
    # the excess of the overflowing regions stays in the sender's spill list and travels through the counted exchange afterwards;
    # segments + spilled rows together are exactly the rows this rank owns
    hot = bk[0] if rank == 0 else None
    hots = [None] * world
    dist.all_gather_object(hots, hot)
    sk = np.where(np.arange(npr) % 2 == 0, np.int64(hots[0]), pk)
    spilled = []
    cap_s = 12 * 1024            # the hot key's owner gets ~15000 rows from every sender
    s_cols, s_cnt, s_ovf = exchange_segments_host(sk, [sk, pv], world, rank, cap_s, all_to_all, spill=spilled)
    assert not s_ovf
    spk = np.concatenate([e[0] for e in spilled]) if spilled else np.zeros(0, dtype=np.int64)
    spv = np.concatenate([e[1] for e in spilled]) if spilled else np.zeros(0, dtype=np.int64)
    rk, rv = exchange_by_key_host(spk, [spk, spv], world, all_to_all)
    k1, v1 = segments_to_dense(s_cols, s_cnt, cap_s)
    mine_k, mine_v = exchange_by_key_host(sk, [sk, pv], world, all_to_all)       # the counted exchange of the same rows
    owner = int(partition_of_keys_np(np.array([hots[0]], dtype=np.int64), world)[0])
    assert len(spk) > 0 and (int(s_cnt.max()) == cap_s) == (owner == rank)     # every sender spills; only the hot key's owner has full regions
    assert np.array_equal(np.sort(np.concatenate([v1, rv])), np.sort(mine_v))
    allk, allv = np.concatenate([k1, rk]), np.concatenate([v1, rv])
    assert np.array_equal(allk[np.argsort(allv)], mine_k[np.argsort(mine_v)])
    plan = JoinPlan(abi.JOIN_INNER, [INT, INT], [INT, INT], [0], [0])
    n, cols = O.OracleJoin(plan, 2).run([Chunk([Column(lbk), Column(lbv)])], Chunk([Column(dk), Column(dv)]).split(1024))
    shard = [c[0] for c in cols]
    allparts = [None] * world
    dist.all_gather_object(allparts, (shard, (bk, bv, pk, pv)))
    if rank == 0:
        got = [np.concatenate([p[0][i] for p in allparts]) for i in range(4)]
        gbk = np.concatenate([p[1][0] for p in allparts]); gbv = np.concatenate([p[1][1] for p in allparts])
        gpk = np.concatenate([p[1][2] for p in allparts]); gpv = np.concatenate([p[1][3] for p in allparts])
        n1, c1 = O.OracleJoin(plan, 2).run([Chunk([Column(gbk), Column(gbv)])], Chunk([Column(gpk), Column(gpv)]).split(1024))
        ok = n1 == len(got[0]) == npr * world and np.array_equal(
            columns_sorted(c1), columns_sorted([(g, np.zeros(len(g), dtype=bool)) for g in got]))
        q.put(bool(ok))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_partitioned_join_gloo():
    from tidb_b200 import build
    build.build()
    import oracle_lib
    oracle_lib.build_oracle()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=240)
        assert p.exitcode == 0, f"rank exited with {p.exitcode}"
    assert q.get(timeout=5) is True


def test_partition_mirror_matches_c_function():
    from tidb_b200 import abi, build
    from tidb_b200.parallel import partition_of_keys_np
    build.build()
    lib = abi.load_lib()
    rng = np.random.default_rng(9)
    keys = np.concatenate([rng.integers(-(1 << 63), (1 << 63) - 1, 3000, dtype=np.int64), np.arange(-50, 50, dtype=np.int64)])
    for nparts in (1, 2, 3, 8, 16):
        exp = np.array([lib.tg_partition_of_key(int(k), nparts) for k in keys])
        assert np.array_equal(partition_of_keys_np(keys, nparts), exp)


def _q3_worker(rank, world, port, q):
    """The distributed Q3-shape PLAN (tidb_b200/q3.py:Q3Distributed, tpch_suite_out.json:99-123) restated with the oracle operators
    and the host exchange: broadcast customer, J1 on the local orders shard, filtered orders and lineitem repartitioned by order
    key, J2 + HashAgg + TopN shard-local (the GROUP BY key contains the partition key), global TopN over world x 10 rows."""
    sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, os.path.join(ROOT, "oracle"))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import torch
    import oracle_lib as O
    import topn as OT
    from test_gpu_q3 import oracle_q3
    from tidb_b200 import abi, q3
    from tidb_b200.chunk import Chunk, Column
    from tidb_b200.parallel import exchange_by_key_host, partition_of_keys_np
    from tidb_b200.plan import FilterItem, JoinPlan
    INT = q3.INT
    d = q3.gen(torch.device("cpu"), 4000, 40000, 160000, rank=rank, world=world)
    h = {k: v.numpy() for k, v in d.__dict__.items()}

    def all_gather(x):
        out = [None] * world
        dist.all_gather_object(out, x)
        return out

    def all_to_all(pieces):
        return [g[rank] for g in all_gather(pieces)]

    # 1. broadcast side + J1 on the local orders shard
    ck = np.concatenate(all_gather(h["c_custkey"])); cs = np.concatenate(all_gather(h["c_seg"]))
    j1 = JoinPlan(abi.JOIN_INNER, [INT] * 4, [INT] * 2, [1], [0], build_is_right=True, lused=[0, 2, 3], rused=[],
                  build_filter=[FilterItem(abi.CMP_EQ, 1, const_i64=q3.SEGMENT)], probe_filter=[FilterItem(abi.CMP_LT, 2, const_i64=q3.DATE)])
    n1, c1 = O.OracleJoin(j1, 2).run(Chunk([Column(ck), Column(cs)]).split(4096),
                                     Chunk([Column(h["o_orderkey"]), Column(h["o_custkey"]), Column(h["o_date"]), Column(h["o_prio"])]).split(4096))
    o_cols = [v.copy() for v, _ in c1] if n1 else [np.zeros(0, dtype=np.int64)] * 3
    # 2. repartition by order key: equal order keys meet on one rank
    ok, od, op = exchange_by_key_host(o_cols[0], o_cols, world, all_to_all)
    lk, lp, ld, ls = exchange_by_key_host(h["l_orderkey"], [h["l_orderkey"], h["l_price"].view(np.int64), h["l_disc"].view(np.int64), h["l_ship"]], world, all_to_all)
    assert np.all(partition_of_keys_np(ok, world) == rank) and np.all(partition_of_keys_np(lk, world) == rank)
    # 3. shard-local J2 + HashAgg: reuse the single-process plan on the shard, with the exchanged J1 output as "orders ⋈ customer"
    shard = {"c_custkey": np.array([1], dtype=np.int64), "c_seg": np.array([q3.SEGMENT], dtype=np.int64),        # one matching customer:
             "o_orderkey": ok, "o_custkey": np.ones(len(ok), dtype=np.int64), "o_date": od, "o_prio": op,          # J1 passes the shard through
             "l_orderkey": lk, "l_price": lp.view(np.float64), "l_disc": ld.view(np.float64), "l_ship": ls}
    _n1, n2, ng, (gk, rev, gd, gp) = oracle_q3(shard)
    assert _n1 == len(ok)
    local_top = OT.topn_rows(list(zip(gk.tolist(), rev.tolist(), gd.tolist(), gp.tolist())), ["int", "real", "int", "int"], [(1, True), (2, False)], 0, 10)
    # 4. global TopN over the world x 10 local rows; group counts add up (groups never span ranks)
    tops = all_gather(local_top); counts = all_gather((n1, n2, ng)); shards = all_gather(h)
    if rank == 0:
        final = OT.topn_rows([r for t in tops for r in t], ["int", "real", "int", "int"], [(1, True), (2, False)], 0, 10)
        whole = {k: np.concatenate([s[k] for s in shards]) for k in h}
        w1, w2, wg, (wk, wrev, wd, wp) = oracle_q3(whole)
        exp = OT.topn_rows(list(zip(wk.tolist(), wrev.tolist(), wd.tolist(), wp.tolist())), ["int", "real", "int", "int"], [(1, True), (2, False)], 0, 10)
        ok_ = (sum(c[0] for c in counts), sum(c[1] for c in counts), sum(c[2] for c in counts)) == (w1, w2, wg) and len(final) == len(exp) == 10
        ok_ = ok_ and all((f[0], f[2], f[3]) == (e[0], e[2], e[3]) and abs(f[1] - e[1]) <= 1e-9 * abs(e[1]) for f, e in zip(final, exp))
        q.put(bool(ok_))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_q3_plan_gloo():
    from tidb_b200 import build
    build.build()
    import oracle_lib
    oracle_lib.build_oracle()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 31500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_q3_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=300)
        assert p.exitcode == 0, f"rank exited with {p.exitcode}"
    assert q.get(timeout=5) is True
