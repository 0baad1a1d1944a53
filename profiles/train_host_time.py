"""Host time to enqueue one training step (train.TrainStep, config 3: 3072 rays, bf16 kernels), eager vs cuda_graph=True.
python profiles/train_host_time.py  ->  profiles/r02_train_host_time.txt"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import bench
import ideal_nerf_b200 as M

dev = torch.device("cuda", 0)
torch.cuda.set_device(dev)
for graph in (False, True):
    r = bench.train_step_bench(M, dev, 50, "bf16", cuda_graph=graph)
    print(json.dumps({k: r[k] for k in ("cuda_graph", "ms_per_step", "host_ms_per_step", "rays_per_s", "loss")}))
