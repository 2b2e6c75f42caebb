"""Run a few device-resident steps of the bench workload (for ncu)."""
import argparse, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from vampire_b200 import ops, synth
from vampire_b200.config import NAMED
from vampire_b200.matrices import prepare_matrices

ap = argparse.ArgumentParser()
ap.add_argument("--config", default="r50_256x704")
ap.add_argument("--batch", type=int, default=8)
ap.add_argument("--dtype", default="bf16")
ap.add_argument("--steps", type=int, default=3)
ap.add_argument("--field", default="surface")
ap.add_argument("--train", action="store_true")
ap.add_argument("--plans", default="off", choices=["off", "on", "both"], help="cached lift plans")
ap.add_argument("--only", default="all", choices=["all", "lift", "render"])
a = ap.parse_args()
cfg = NAMED[a.config]
dt = {"bf16": torch.bfloat16, "fp32": torch.float32, "fp16": torch.float16}[a.dtype]
cid = ops.register_config(cfg)
m = synth.make_mats(cfg, a.batch, "val")
prep = prepare_matrices(m["sensor2ego_mats"][:, 0], m["intrin_mats"][:, 0], m["ida_mats"][:, 0], m["bda_mat"]).cuda()
depth, ctx = [t.cuda() for t in synth.make_lift_inputs(cfg, a.batch, dtype=dt)]
den, sem, feat, rgb = [t.cuda() for t in synth.make_render_inputs(cfg, a.batch, field=a.field, dtype=dt)]
beta = torch.tensor(0.1, device="cuda")
leaves = [depth, ctx, den, sem, feat, rgb, beta]
tables = [None]
rtab = {None: None}
if a.plans != "off":
    from vampire_b200.plan import PlanCache
    pc = PlanCache()
    tab = pc.lift(ops.state(cid), cid, prep, True).table
    tables = [tab] if a.plans == "on" else [None, tab]
    rtab[tab] = pc.render(ops.state(cid), cid, prep, True).table
torch.manual_seed(0)
for _ in range(a.steps):
  for tab in tables:
    if a.train:
        for t in leaves:
            t.requires_grad_(True); t.grad = None
        outs = []
        if a.only in ("all", "lift"):
            outs.append(ops.lift_pool_fwd(depth, ctx, prep, cid, True, False, True, tab)[0])
        if a.only in ("all", "render"):
            outs += list(ops.render_fwd(den, sem, rgb, feat, beta, prep, None, cid, True, 3))
        torch.autograd.backward(outs, [torch.randn_like(o) for o in outs])
    else:
        with torch.no_grad():
            if a.only in ("all", "lift"):
                ops.lift_pool_fwd(depth, ctx, prep, cid, True, False, False, tab)
            if a.only in ("all", "render"):
                ops.render_fwd(den, sem, rgb, feat, beta, prep, None, cid, True, 3, rtab[tab])
torch.cuda.synchronize()
print("done")
