import sys; sys.path.insert(0, ".")
import torch, recsys_b200 as rs
cfg = rs.ModelConfig(); w = rs.synth.make_weights(cfg, 6040, 3706, seed=10)
m = rs.MultiTaskModel(cfg, rs.synth.make_vocab(6040), rs.synth.make_vocab(3706), {}); rs.load_weights(m, w)
B = 65536; f = {"user_id": torch.randint(0, 6041, (B,), device="cuda"), "movie_id": torch.randint(0, 3707, (B,), device="cuda")}
m(f); torch.cuda.synchronize(); m(f); torch.cuda.synchronize()
