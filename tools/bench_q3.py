#!/usr/bin/env python
"""Q3-shape at SF-like sizes on one GPU (default SF=10: 1.5M / 15M / 60M rows; SF=100 needs ~30 GB): end-to-end device time
and GB/s over the scanned column bytes.  Verified against the torch rendering of the query."""
import argparse, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from bench import peaks
from tidb_b200 import q3
ap = argparse.ArgumentParser(); ap.add_argument("--sf", type=float, default=10); ap.add_argument("--steps", type=int, default=3)
ap.add_argument("--out", default="gpurun_out/bench_q3.jsonl"); ap.add_argument("--verify", type=int, default=1)
a = ap.parse_args()
dev = torch.device("cuda", 0); stream = torch.cuda.Stream(device=dev)
nc, no, nl = int(150_000 * a.sf), int(1_500_000 * a.sf), int(6_000_000 * a.sf)
with torch.cuda.stream(stream):
    d = q3.gen(dev, nc, no, nl)
    got = q3.run(d, dev, stream)
    if a.verify:
        exp = q3.reference(d)
        order = torch.argsort(got["orderkey"])
        assert torch.equal(got["orderkey"][order], exp["orderkey"]) and torch.equal(got["o_date"][order], exp["o_date"])
        assert torch.allclose(got["revenue"][order], exp["revenue"], rtol=1e-6, atol=0)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(a.steps):
        q3.run(d, dev, stream, keep_groups=False)
    e1.record(stream)
    stream.synchronize()
    t = {}
    q3.run(d, dev, stream, keep_groups=False, timings=t)
stream.synchronize()
ms = e0.elapsed_time(e1) / a.steps
peak, src = peaks()
rec = dict(op="Q3-shape", sf=a.sf, rows=dict(customer=nc, orders=no, lineitem=nl), groups=int(got["orderkey"].numel()), ms=ms, phases_ms={k: round(v, 3) for k, v in t.items() if k != "rows"}, operator_rows=t.get("rows"),
           scanned_gb=d.scanned_bytes() / 1e9, gbs=d.scanned_bytes() / ms / 1e6, frac=d.scanned_bytes() / ms / 1e6 / peak, verified=bool(a.verify))
print(json.dumps(rec)); os.makedirs(os.path.dirname(a.out), exist_ok=True); open(a.out, "a").write(json.dumps(rec) + "\n")
