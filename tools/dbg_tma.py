import os, sys
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
os.environ.update(TG_PROBE_TMA="1", TG_PROBE_PARTITION="0", TG_PROBE_PART_MIN_MB="0")
import numpy as np
from tidb_b200 import abi
from tidb_b200.chunk import Chunk, Column
from tidb_b200.executor import HashJoinExec, MockDataSource, drain
from tidb_b200.plan import FieldType, JoinPlan
INT_NN = FieldType(abi.TYPE_LONGLONG, abi.FLAG_NOT_NULL)
rng = np.random.default_rng(17)
nb, npr = 300_001, 2_500_003
bk = rng.permutation(nb).astype(np.int64) * 2654435761 - (1 << 40)
bk[5] = -(1 << 63)
build = Chunk([Column(bk), Column(np.arange(nb, dtype=np.int64) * 3)])
pick = rng.integers(0, int(nb * 1.25), npr)
pk = np.where(pick < nb, bk[np.minimum(pick, nb - 1)], pick.astype(np.int64) * 7 + 1)
probe = Chunk([Column(pk), Column(np.arange(npr, dtype=np.int64))])
plan = JoinPlan(abi.JOIN_INNER, [INT_NN, INT_NN], [INT_NN, INT_NN], [0], [0])
e = HashJoinExec(plan, MockDataSource(plan.left_types, [probe]), MockDataSource(plan.right_types, [build]))
chunks = drain(e, 1 << 22)
got = [np.concatenate([c.columns[i].data for c in chunks]) for i in range(4)]
order = np.argsort(bk); sb = bk[order]
pos = np.searchsorted(sb, pk); pos[pos >= nb] = nb - 1
hit = sb[pos] == pk
exp_rows = np.nonzero(hit)[0]
missing = np.setdiff1d(exp_rows, got[1])
print("expected", len(exp_rows), "got", len(got[0]), "missing rows", missing, "keys", pk[missing], "tile", missing // 1024, "in-tile", missing % 1024)
print("sentinel probe rows:", np.nonzero(pk == -(1 << 63))[0])
extra = np.setdiff1d(got[1], exp_rows); print("extra", extra[:10])
u, cnts = np.unique(got[1], return_counts=True)
print("dup row ids in output:", int((cnts > 1).sum()), "max mult", int(cnts.max()))
valid = (got[1] >= 0) & (got[1] < npr)
print("row ids out of range:", int((~valid).sum()))
g1 = np.clip(got[1], 0, npr - 1)
print("key!=pk[rowid]:", int((got[0] != pk[g1]).sum()), " probekey!=buildkey:", int((got[0] != got[2]).sum()))
bp = (order[np.clip(np.searchsorted(sb, got[0]), 0, nb - 1)] * 3)
print("build payload wrong for its key:", int((got[3] != bp).sum()))
bad = np.nonzero(got[0] != pk[g1])[0]
print("bad output positions (first 40):", bad[:40], "their rowids", got[1][bad[:10]], "keys", got[0][bad[:5]])
