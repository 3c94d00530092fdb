"""Where the wall clock of a device-resident call goes (run on the GPU box): python wall vs library wall vs CUDA-event sum,
with SCLANE_TRACE=1 printing the library's host-side phase timestamps.  usage: wall_gap.py [genes] [cells] [batch_streams]"""
import os
import sys
import time

os.environ["SCLANE_TRACE"] = "1"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch

import bench
import sclane_b200 as sb

G = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
N = int(sys.argv[2]) if len(sys.argv) > 2 else 200000
BS = int(sys.argv[3]) if len(sys.argv) > 3 else 0
dev = torch.device("cuda", 0)
d_counts, pt = bench.make_counts_gpu(0, G, N, dev)
torch.cuda.synchronize()
o = sb.api._opts(gpu_ids=[0], batch_streams=BS)
for it in range(3):
    t0 = time.perf_counter()
    recs, _ = sb.run_records(G, pt, None, o, device_ptr=d_counts.data_ptr(), device=0)
    t1 = time.perf_counter()
    p = sb.get_profile()
    print(f"step {it}: python wall {1000 * (t1 - t0):.2f} ms, library wall {p['wall_ms']:.2f} ms, events {sum(p['ms'].values()):.2f} ms", flush=True)
