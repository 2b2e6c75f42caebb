"""One pass over every libvb200 kernel family at the MINI geometry, for compute-sanitizer:

    compute-sanitizer --tool memcheck python tools/memcheck_probe.py [--full]
"""
import dataclasses, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from vampire_b200 import cabi, ops, synth
from vampire_b200.config import MINI, R50_256x704
from vampire_b200.matrices import prepare_matrices
from vampire_b200.plan import PlanCache
from vampire_b200.view_transform import LiftRenderB200

torch.manual_seed(0)
FULL = "--full" in sys.argv          # the R50 256x704 geometry, B = 1 (the fast paths the bench runs)
BASE, B = (R50_256x704, 1) if FULL else (MINI, 2)
for cfg in (BASE, dataclasses.replace(BASE, density_mode="naive", cat_seg=True)):
    mod = LiftRenderB200(plans="off", **cfg.backbone_kwargs()).cuda().train()
    mats = synth.make_mats(cfg, B, "stress")
    prep = prepare_matrices(mats["sensor2ego_mats"][:, 0], mats["intrin_mats"][:, 0], mats["ida_mats"][:, 0],
                            mats["bda_mat"]).cuda()
    for dt in (torch.float32, torch.bfloat16):
        depth, ctx = [t.cuda().requires_grad_(True) for t in synth.make_lift_inputs(cfg, B, dtype=dt)]
        den, sem, feat, rgb = [t.cuda().requires_grad_(True) for t in synth.make_render_inputs(cfg, B, dtype=dt)]
        cid = mod.cfg_id
        st = ops.state(cid)
        pc = PlanCache()
        for plan in (None, pc.lift(st, cid, prep, True).table):
            for cl in (False, True):
                vox, _ = ops.lift_pool_fwd(depth, ctx, prep, cid, True, cl, True, plan)
                vox.float().sum().backward()
        rplan = pc.render(st, cid, prep, True).table
        for seg in (1, 2, 8):
            cabi.render_set_march_split(seg)
            for plan in (None, rplan):
                with torch.no_grad():
                    ops.render_fwd(den, sem, rgb, feat, mod._beta(den.device), prep, None, cid, True, 3, plan)
        cabi.render_set_march_split(0)
        with torch.no_grad():
            ops.render_fwd(den, sem, rgb, feat, mod._beta(den.device), prep, None, cid, True, 3, rplan, True)
        outs = mod.render(mats, den, sem, feat, rgb)
        sum(o.float().sum() for o in outs).backward()
        geom = ops.get_geometry(prep, cid, True, True)
        with torch.no_grad():
            mod.volume_rendering_from_multiple_views(geom, den, sem, feat, rgb)
            ops.get_pixel(prep, cid, True)
            ops.lift_indices(prep, cid, True)
            ops.render_indices(prep, cid, True)
        logits = torch.randn(B * cfg.num_cams, cfg.D, cfg.fH, cfg.fW, device="cuda", dtype=dt, requires_grad=True)
        mod.depth_softmax(logits).sum().backward()
        maps = torch.randn(B, cfg.num_cams, 5, cfg.fH, cfg.fW, device="cuda", requires_grad=True)
        mod.upsample2d(maps).sum().backward()
        pts = torch.rand(777, 3, device="cuda") * 120 - 60
        lg, sd = mod.query_points(sem, den, pts)
        (lg.sum() + sd.sum()).backward()
        ol, od = mod.occupancy(sem, den, mats["bda_mat"], LiftRenderB200.occ_coords())
        (ol.sum() + od.sum()).backward()
        img = torch.randn(B, cfg.num_cams, cfg.C, cfg.fH, cfg.fW, device="cuda", dtype=dt, requires_grad=True)
        mod.lift_pool_2d(img, mats).float().sum().backward()
        fr = (depth.unsqueeze(2) * ctx.unsqueeze(3)).detach().requires_grad_(True)
        mod.get_voxel_feats(fr, 0, mats).float().sum().backward()
        torch.cuda.synchronize()
        print("ok", cfg.density_mode, cfg.cat_seg, dt, flush=True)
print("memcheck probe done")
