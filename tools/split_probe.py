"""Time the render forward (camera branch alone, and both branches) for depth-split settings of the march.

  python tools/split_probe.py [--batches 1 2 4 8] [--dtype bf16]
Prints one line per (batch, planned, segments): CUDA-event ms per call, median of `--reps` calls after warm-up.
"""
import argparse, os, statistics, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from vampire_b200 import cabi, ops, synth
from vampire_b200.config import NAMED
from vampire_b200.matrices import prepare_matrices
from vampire_b200.plan import PlanCache

ap = argparse.ArgumentParser()
ap.add_argument("--config", default="r50_256x704")
ap.add_argument("--batches", type=int, nargs="+", default=[1, 2, 4, 8])
ap.add_argument("--dtype", default="bf16")
ap.add_argument("--reps", type=int, default=20)
ap.add_argument("--segments", type=int, nargs="+", default=[1, 2, 4, 8, 0])
a = ap.parse_args()
cfg = NAMED[a.config]
dt = {"bf16": torch.bfloat16, "fp32": torch.float32, "fp16": torch.float16}[a.dtype]
cid = ops.register_config(cfg)
beta = torch.tensor(0.1, device="cuda")
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")


def timed(fn, reps):
    ts = []
    for i in range(reps + 3):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        if i >= 3:
            ts.append(e0.elapsed_time(e1))
    return statistics.median(ts)


for B in a.batches:
    m = synth.make_mats(cfg, B, "val")
    prep = prepare_matrices(m["sensor2ego_mats"][:, 0], m["intrin_mats"][:, 0], m["ida_mats"][:, 0], m["bda_mat"]).cuda()
    den, sem, feat, rgb = [t.cuda() for t in synth.make_render_inputs(cfg, B, field="surface", dtype=dt)]
    table = PlanCache().render(ops.state(cid), cid, prep, True).table
    for planned in (False, True):
        for branches in (1, 3):
            row = []
            for nseg in a.segments:
                cabi.render_set_march_split(nseg)
                with torch.no_grad():
                    ms = timed(lambda: ops.render_fwd(den, sem, rgb, feat, beta, prep, None, cid, True, branches,
                                                      table if planned else None), a.reps)
                row.append(f"seg{nseg}={ms:.4f}")
            print(f"B={B} {a.dtype} planned={int(planned)} branches={branches}: " + "  ".join(row), flush=True)
    cabi.render_set_march_split(0)
    del den, sem, feat, rgb
