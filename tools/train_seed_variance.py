"""B = 1 train step (fp32, graph-replayed) timed for the samples that ranks 0..7 of an 8-GPU weak-scaling run draw.
The bench takes the MAX over ranks, so the spread between samples bounds what 'scaling efficiency' can show."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from vampire_b200.config import NAMED

sys.argv = [sys.argv[0]]
args = bench.parse_args()
cfg = NAMED[args.config]
dev = torch.device("cuda", 0)
torch.cuda.set_device(dev)
for r in range(8):
    w = bench.Workload(args, cfg, dev, 1, r, True, 1, "fp32", False, allreduce=True, graph=True)
    ms = w.timed(10, 3)
    w.free()
    print(f"sample of rank {r} (seed {1234 + r}): {ms:.4f} ms/step", flush=True)
