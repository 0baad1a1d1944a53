"""BASELINE.json config 3 alone (bench.py's train_step_bench): N_rand=3072 training step, for ncu launch lists / captures.
    python profiles/train_step.py [bf16|fp32] [steps]"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
import ideal_nerf_b200 as M

mode = sys.argv[1] if len(sys.argv) > 1 else "bf16"
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 4
print(json.dumps(bench.train_step_bench(M, torch.device("cuda", 0), steps, mode)))
