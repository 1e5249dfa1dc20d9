"""Q3-shape pipeline (BASELINE configs[3]) at reduced scale: 2 joins with fused Selections + projection + aggregation
through the C-ABI operators, against the same query in plain torch.  Keys / dates / priorities bit-exact, SUM within 1e-6."""
import pytest
import torch

from tidb_b200 import q3

pytestmark = pytest.mark.gpu


def test_q3_shape_small():
    dev = torch.device("cuda", 0)
    stream = torch.cuda.Stream(device=dev)
    with torch.cuda.stream(stream):
        d = q3.gen(dev, 15_000, 150_000, 600_000)
        got = q3.run(d, dev, stream)
        exp = q3.reference(d)
        stream.synchronize()
        assert got["orderkey"].numel() == exp["orderkey"].numel() > 0
        order = torch.argsort(got["orderkey"])
        assert torch.equal(got["orderkey"][order], exp["orderkey"])          # sorted unique keys
        assert torch.equal(got["o_date"][order], exp["o_date"]) and torch.equal(got["o_prio"][order], exp["o_prio"])
        assert torch.allclose(got["revenue"][order], exp["revenue"], rtol=1e-6, atol=0)
