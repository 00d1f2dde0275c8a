import sys; sys.path.insert(0, ".")
import numpy as np, torch, recsys_b200 as rs
B=65536
rng=np.random.default_rng(0)
for K,N in ((256,128),(64,128)):
    x=torch.randn((B,K),device="cuda"); wk=torch.from_numpy(rs.synth.glorot_uniform(rng,K,N)).cuda(); b=torch.zeros((N,),device="cuda")
    hi,lo=rs.ops.dense_x3_prepare(wk); y=torch.empty((B,N),device="cuda")
    for _ in range(3): rs.ops.dense_x3(x,hi,lo,b,"relu",out=y)
    torch.cuda.synchronize()
