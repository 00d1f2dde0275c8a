"""One peer-memory lookup (B = 65,536 Zipf ids, de-duplicated) into a 10M-row table behind a communicator of one rank -- the
kernel chain of BASELINE configs[3] without the NVLink hops -- for ncu launch lists."""
import sys
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import recsys_b200 as rs  # noqa: E402
from recsys_b200.sharded import ShardedEmbedding  # noqa: E402

dev = torch.device("cuda", 0)
n_rows, B = 10_000_000, 65536
table = ShardedEmbedding.fill_synthetic(n_rows, 128, 0, 1, 42, dev)
emb = ShardedEmbedding(n_rows, 128, local_table=table)
emb.enable_peer_exchange(B, dedup=True)
rng = np.random.default_rng(0)
ids = torch.from_numpy(rs.synth.zipf_ids(rng, n_rows, B).astype(np.int64)).to(dev)
for _ in range(3):
    out = emb.lookup(ids)
    torch.cuda.synchronize()
assert np.array_equal(out[:1000].cpu().numpy(), rs.synth.hash_table_rows(ids[:1000].cpu().numpy(), 128, seed=42))
emb.close()
