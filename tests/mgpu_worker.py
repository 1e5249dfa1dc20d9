"""Worker of tests/test_gpu_multigpu.py (one process per GPU, launched with torch.distributed.run): the product path of
bench.py --gpus N at test size — build side through the counted exchange, probe side through MailboxExchange +
tg_join_probe_dev_seg for several pipelined steps — every rank dumps what it produced for the parent to compare with
the oracle.  Also exercises a forced region overflow (must be reported on every rank, never silently dropped)."""
import argparse
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def gen(rank, world, nb, npr, step):
    """deterministic shards: unique build keys over the global key set, probe keys 70 % matching"""
    rng = np.random.default_rng(1000 + rank)
    ids = rng.permutation(nb).astype(np.int64) + rank * nb
    bk, bv = ids * np.int64(-7046029254386353131), ids * 7
    rng = np.random.default_rng(5000 + 97 * step + rank)
    pid = rng.integers(0, nb * world, npr).astype(np.int64)
    miss = rng.random(npr) < 0.3
    pk = np.where(miss, rng.integers(1 << 40, 1 << 41, npr).astype(np.int64) * 2 + 1, pid * np.int64(-7046029254386353131))
    pv = np.arange(npr, dtype=np.int64) + (step * world + rank) * npr
    return bk, bv, pk, pv


HOT_ID = 12345   # a build row of rank 0's shard: the key most probe rows of the skewed steps carry


def gen_skew(rank, world, nb, npr, step):
    """the probe shard of gen() with 60 % of the rows rewritten to ONE hot key: its owner's receive regions overflow"""
    _, _, pk, pv = gen(rank, world, nb, npr, step)
    rng = np.random.default_rng(9000 + 31 * step + rank)
    hot = (np.array([HOT_ID], dtype=np.int64) * np.int64(-7046029254386353131))[0]   # wraps like the build keys of gen()
    return np.where(rng.random(npr) < 0.6, hot, pk), pv


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", required=True)
    ap.add_argument("--build-rows", type=int, default=60_000)
    ap.add_argument("--probe-rows", type=int, default=300_000)
    ap.add_argument("--steps", type=int, default=4)
    ap.add_argument("--dma", type=int, default=0, help="0: SM stores, 1: copy engines, 2: hybrid (copy engines + 1 direct peer), 3: SM region copy kernel")
    a = ap.parse_args()
    import torch
    import torch.distributed as dist
    from tidb_b200 import abi
    from tidb_b200.device import DeviceJoin
    from tidb_b200.parallel import KeyExchange, MailboxExchange
    from tidb_b200.plan import FieldType, JoinPlan
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    stream, xstream = torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev)
    INT = FieldType(abi.TYPE_LONGLONG, abi.FLAG_NOT_NULL)
    plan = JoinPlan(abi.JOIN_INNER, [INT, INT], [INT, INT], [0], [0], build_is_right=True, device=local, stream=stream.cuda_stream)
    t = lambda x: torch.from_numpy(x).to(dev)

    def view(p, n):
        class _A:
            pass
        o = _A()
        o.__cuda_array_interface__ = {"shape": (n,), "typestr": "<i8", "data": (p, False), "version": 3}
        return torch.as_tensor(o, device=dev)

    bk, bv, _, _ = gen(rank, world, a.build_rows, a.probe_rows, 0)
    with torch.cuda.stream(stream):
        xb = KeyExchange(rank, world, local, stream, 2, int(a.build_rows * 1.2) + 4096, "p2p")
        dbk, dbv = t(bk), t(bv)
        lbk, lbv = xb.exchange(dbk, [dbk, dbv])
        join = DeviceJoin(plan)
        join.build([lbk, lbv])
    xm = MailboxExchange(rank, world, local, xstream, 2, a.probe_rows, dma=bool(a.dma), direct_peers=1 if a.dma == 2 else 0, sm_copy=a.dma == 3)
    inputs = []
    for s in range(a.steps):
        _, _, pk, pv = gen(rank, world, a.build_rows, a.probe_rows, s)
        inputs.append((t(pk), t(pv)))
    torch.cuda.synchronize(dev)
    dist.barrier()
    # pipelined exactly like bench.py: every SM kernel on ONE stream, the exchange one step ahead of the probe.  Rule of the
    # mailbox protocol: a wait may only depend on signals enqueued EARLIER in program order on every rank (send(k+1)'s ACK
    # wait needs release(k-1), recv(k) needs send(k)) — then no host-side synchronisation can deadlock against a spinning
    # wait kernel.
    outs = []
    with torch.cuda.stream(stream):
        xm.send(inputs[0][0], [inputs[0][0], inputs[0][1]], stream)
        for s in range(a.steps):
            if s + 1 < a.steps:
                xm.send(inputs[s + 1][0], [inputs[s + 1][0], inputs[s + 1][1]], stream)
            cols_in, seg_cnt, cap, set_, ep = xm.recv(stream)
            rows, cols, _ = join.probe_segments(cols_in, seg_cnt, cap, sync=True)
            outs.append([view(p, rows).cpu().numpy().copy() for p in cols])
            xm.release(stream, set_, ep)
    xm.check()
    xm.close()
    res = {f"s{s}c{c}": outs[s][c] for s in range(a.steps) for c in range(4)}
    # skewed keys with a spill area: the hot key's owner cannot hold 60 % of every sender's rows in its regions; the regroup
    # kernel appends what does not fit to the local spill area, the counted exchange moves it afterwards, nothing is lost
    xs = MailboxExchange(rank, world, local, xstream, 2, a.probe_rows, dma=bool(a.dma), direct_peers=1 if a.dma == 2 else 0, sm_copy=a.dma == 3,
                         spill_rows=2 * a.probe_rows)
    skew_steps = 2
    sk = [tuple(t(x) for x in gen_skew(rank, world, a.build_rows, a.probe_rows, s)) for s in range(skew_steps)]
    torch.cuda.synchronize(dev)
    dist.barrier()
    souts = []
    with torch.cuda.stream(stream):
        xs.send(sk[0][0], [sk[0][0], sk[0][1]], stream)
        for s in range(skew_steps):
            if s + 1 < skew_steps:
                xs.send(sk[s + 1][0], [sk[s + 1][0], sk[s + 1][1]], stream)
            cols_in, seg_cnt, cap, set_, ep = xs.recv(stream)
            rows, cols, _ = join.probe_segments(cols_in, seg_cnt, cap, sync=True)
            souts.append([view(p, rows).cpu().numpy().copy() for p in cols])
            xs.release(stream, set_, ep)
    xs.check()                       # the spill area took the excess: no overflow error
    nsp, spcols = xs.drain_spill()
    res["spilled_rows"] = np.array([nsp])
    with torch.cuda.stream(stream):
        xk = KeyExchange(rank, world, local, stream, 2, 2 * a.probe_rows * skew_steps * world + 4096, "p2p")
        lk, lv = xk.exchange(spcols[0], [spcols[0], spcols[1]])
        if lk.numel() > 0:           # only the hot key's owner receives spilled rows
            rows, cols, _ = join.probe([lk, lv], sync=True)
            if rows > 0:
                souts.append([view(p, rows).cpu().numpy().copy() for p in cols])
    torch.cuda.synchronize(dev)
    xk.close()
    xs.close()
    for c in range(4):
        res[f"skew_c{c}"] = np.concatenate([o[c] for o in souts])
    # forced overflow: regions far too small for the rows that arrive -> every rank must see the error
    xo = MailboxExchange(rank, world, local, xstream, 2, 4096, slack=1.0)
    with torch.cuda.stream(stream):
        xo.send(inputs[0][0], [inputs[0][0], inputs[0][1]], stream)
        _c, _n, _cap, set_, ep = xo.recv(stream)
        xo.release(stream, set_, ep)
    try:
        xo.check()
        res["overflow_detected"] = np.array([0])
    except RuntimeError as e:
        res["overflow_detected"] = np.array([1 if "overflow" in str(e) else 2])
    xo.close()
    # a sender that never shows up: the wait must time out and raise the error flag instead of hanging the GPU
    xt = MailboxExchange(rank, world, local, xstream, 2, 4096, timeout_ms=300)
    with torch.cuda.stream(stream):
        xt.recv(stream)          # nobody sent step 0
    try:
        xt.check()
        res["timeout_detected"] = np.array([0])
    except RuntimeError as e:
        res["timeout_detected"] = np.array([1 if "timed out" in str(e) else 2])
    xt.close()
    np.savez(os.path.join(a.out, f"rank{rank}.npz"), **res)
    join.close()
    xb.close()
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
