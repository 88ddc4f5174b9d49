"""Two eager answers of one config-2 query (for ncu: the 14th k_ks_level_cluster launch is level 6, warm)."""
import os, sys
os.environ["PIRB_GRAPHS"] = "0"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import torch
import bench
from pir_b200 import sharded
import pir_b200 as pb
nq = int(sys.argv[1]) if len(sys.argv) > 1 else 1
params = bench.make_params("cfg2")
srv = sharded.ShardServer(params, device=0); srv.db.fill_random(1)
q, elts, keys = bench.synth_inputs(params, nq, 5)
srv.set_keys(pb.GaloisKeys(elts, keys.reshape(-1)))
dq = sharded.to_device(q, srv.device)
for _ in range(2):
    srv.answer(dq)
torch.cuda.synchronize()
