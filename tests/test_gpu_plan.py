"""GPU: the cached projection / sort plan (vb200_lift_plan_build) and the plan-driven lift + pool.

The plan must hold exactly the integers and fractions of the strict index kernel (vb200_lift_indices, itself
SHA-pinned to the reference in test_gpu_geometry.py), and the planned forward / backward must be bit-identical
to the kernels that recompute the projection on every call."""
import numpy as np
import pytest
import torch

from helpers import Case, assert_close_scaled
from oracle import torch_path as tp

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module", params=["mini_val", "mini_stress"])
def case(request):
    return Case(request.param)


def _state(cfg, lift_2d=False):
    from vampire_b200 import ops
    cid = ops.register_config(cfg, lift_2d)
    return ops, cid, ops.state(cid)


def _batch(case, st):
    from vampire_b200.plan import LiftPlanBatch, build_lift_plans
    plans = build_lift_plans(st, case.prep.cuda(), True)
    return plans, LiftPlanBatch(plans, torch.device("cuda", torch.cuda.current_device()))


def test_plan_holds_the_strict_indices(case):
    """pairs / cell records == (valid, x0, y0, z0, fractions) of the bit-exact index kernel, compacted."""
    ops, cid, st = _state(case.cfg)
    cfg = case.cfg
    valid, i0, frac = ops.lift_indices(case.prep.cuda(), cid, True)
    plans, _ = _batch(case, st)
    nvox = cfg.vZ * cfg.vY * cfg.vX
    for b, p in enumerate(plans):
        v = valid[b].reshape(cfg.num_cams, nvox).cpu().numpy().astype(bool)
        ii = i0[b].reshape(cfg.num_cams, nvox, 3).cpu().numpy().astype(np.int64)
        ff = frac[b].reshape(cfg.num_cams, nvox, 3).cpu().numpy()
        assert p.num_pairs == int(v.sum())
        head = p.head.cpu().numpy().view(np.uint32)
        cnt = head & 15
        first = head >> 4
        assert np.array_equal(cnt, v.sum(0))
        assert np.array_equal(first, np.concatenate([[0], np.cumsum(cnt)[:-1]]))
        pairs = p.pairs.cpu().numpy().view(np.uint32)
        key = pairs[:p.num_pairs, 0]
        n = key >> 29
        z0 = ((key >> 20) & 511).astype(np.int64) - 1
        y0 = ((key >> 10) & 1023).astype(np.int64) - 1
        x0 = (key & 1023).astype(np.int64) - 1
        vox = np.repeat(np.arange(nvox), cnt)
        assert v[n, vox].all()
        # voxel-major, cameras ascending
        assert np.array_equal(np.stack([vox, n], 1), np.argwhere(v.T))
        assert np.array_equal(np.stack([x0, y0, z0], 1), ii[n, vox])
        fr = pairs[:p.num_pairs, 1:].view(np.float32)
        # vb200_lift_indices reports ix - floor(ix); the plan stores ix - (float)x0: the same subtraction
        assert np.array_equal(fr, ff[n, vox])
        # cell-major copy: same multiset of (voxel, z0, fractions), CSR by (n, y0+1, x0+1), sorted by (z0, voxel)
        off = p.cell_off.cpu().numpy()
        recs = p.cell_recs.cpu().numpy().view(np.uint32)[:p.num_pairs]
        assert off[0] == 0 and off[-1] == p.num_pairs and (np.diff(off) >= 0).all()
        cell_of_pair = (n.astype(np.int64) * (cfg.fH + 1) + (y0 + 1)) * (cfg.fW + 1) + (x0 + 1)
        assert np.array_equal(np.bincount(cell_of_pair, minlength=off.size - 1), np.diff(off))
        rkey = recs[:, 0].astype(np.int64)
        cell_of_rec = np.repeat(np.arange(off.size - 1), np.diff(off))
        order = np.lexsort((rkey, cell_of_rec))
        assert np.array_equal(order, np.arange(p.num_pairs)), "cell segments are not sorted by (z0, voxel)"
        want = np.lexsort((((z0 + 1) << 21) | vox, cell_of_pair))
        assert np.array_equal(rkey, (((z0 + 1) << 21) | vox)[want])
        assert np.array_equal(recs[:, 1:], pairs[:p.num_pairs, 1:][want])


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16, torch.float16])
@pytest.mark.parametrize("channels_last", [False, True])
def test_planned_forward_is_bit_identical(case, dtype, channels_last):
    ops, cid, st = _state(case.cfg)
    _, batch = _batch(case, st)
    d, c, m = case.depth.to(dtype).cuda(), case.ctx.to(dtype).cuda(), case.prep.cuda()
    a, ca = ops.lift_pool_fwd(d, c, m, cid, True, channels_last, True)
    b, cb = ops.lift_pool_fwd(d, c, m, cid, True, channels_last, True, batch.table)
    assert torch.equal(a, b) and torch.equal(ca, cb)
    assert b.permute(0, 2, 3, 4, 1).is_contiguous() == channels_last


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_planned_backward_is_bit_identical(case, dtype):
    ops, cid, st = _state(case.cfg)
    _, batch = _batch(case, st)
    cot = case.cotangents()[0].to(dtype).cuda()
    grads = []
    for plan in (None, batch.table):
        d = case.depth.to(dtype).cuda().requires_grad_(True)
        c = case.ctx.to(dtype).cuda().requires_grad_(True)
        out, _ = ops.lift_pool_fwd(d, c, case.prep.cuda(), cid, True, False, True, plan)
        grads.append(torch.autograd.grad((out * cot).sum(), [d, c]))
    for ga, gb in zip(*grads):
        assert torch.equal(ga, gb)


def test_planned_2d_lift_is_bit_identical(case):
    """The BaseBiLinear D = 1 mode goes through the same plan machinery."""
    ops, cid, st = _state(case.cfg, lift_2d=True)
    _, batch = _batch(case, st)
    c = case.ctx.cuda()
    ones = torch.ones(case.batch, case.cfg.num_cams, 1, case.cfg.fH, case.cfg.fW, device="cuda")
    a, _ = ops.lift_pool_fwd(ones, c, case.prep.cuda(), cid, True, False, False)
    b, _ = ops.lift_pool_fwd(ones, c, case.prep.cuda(), cid, True, False, False, batch.table)
    assert torch.equal(a, b)


def test_plan_cache_hits_and_module_modes(case):
    from vampire_b200.view_transform import LiftRenderB200
    outs = {}
    for mode in ("off", "always", "eval"):
        mod = LiftRenderB200(plans=mode, **case.conf).cuda()
        if mode == "eval":
            mod.eval()
        with torch.no_grad():
            outs[mode] = mod.lift_pool(case.depth.cuda(), case.ctx.cuda(), case.mats)
            again = mod.lift_pool(case.depth.cuda(), case.ctx.cuda(), case.mats)
        assert torch.equal(outs[mode], again)
        pc = mod.plan_cache
        if mode == "off":
            assert pc.hits == 0 and pc.misses == 0
        else:
            assert pc.misses == case.batch and pc.hits == case.batch and pc.nbytes() > 0
            # a permuted batch of known rigs is still a hit (per-sample keys)
            perm = {k: (v.flip(0) if v is not None else None) for k, v in case.mats.items()}
            with torch.no_grad():
                flipped = mod.lift_pool(case.depth.flip(0).cuda(), case.ctx.flip(0).cuda(), perm)
            assert pc.misses == case.batch
            assert torch.equal(flipped.flip(0), outs[mode])
    assert torch.equal(outs["off"], outs["always"]) and torch.equal(outs["off"], outs["eval"])
    # training mode never builds plans under plans="eval"
    mod = LiftRenderB200(plans="eval", **case.conf).cuda().train()
    d = case.depth.cuda().requires_grad_(True)
    mod.lift_pool(d, case.ctx.cuda(), case.mats).sum().backward()
    assert mod.plan_cache.misses == 0 and d.grad is not None


@pytest.mark.parametrize("cdt", [torch.bfloat16, torch.float16])
def test_amp_fp32_depth_with_16bit_ctx(case, cdt):
    """Reference under AMP (BV2:551-553): fp32 softmax output x 16-bit ctx promotes to fp32 -- the pooled volume is
    fp32 and the depth probabilities are NOT rounded to 16 bits.  Oracle: fp32 path fed the 16-bit-valued ctx."""
    ops, cid, st = _state(case.cfg)
    buf = tp.build_buffers(case.conf)
    ctx16 = case.ctx.to(cdt)
    d = case.depth.cuda().requires_grad_(True)
    c = ctx16.cuda().requires_grad_(True)
    out, _ = ops.lift_pool_fwd(d, c, case.prep.cuda(), cid, True, False, True)
    assert out.dtype == torch.float32
    dr = case.depth.clone().requires_grad_(True)
    cr = ctx16.float().requires_grad_(True)
    ref = tp.lift_pool(case.conf, buf, dr, cr, case.mats)
    assert_close_scaled(out.detach().cpu().numpy(), ref.detach().numpy(), 1e-5, "AMP pooled volume (fp32)")
    cot = case.cotangents()[0]
    gd, gc = torch.autograd.grad((out * cot.cuda()).sum(), [d, c])
    rd, rc = torch.autograd.grad((ref * cot).sum(), [dr, cr])
    assert gd.dtype == torch.float32 and gc.dtype == cdt
    assert_close_scaled(gd.cpu().numpy(), rd.numpy(), 2e-5, "AMP d_depth (fp32)")
    assert_close_scaled(gc.float().cpu().numpy(), rc.numpy(), 1e-2, "AMP d_ctx (16-bit)")
    # the module keeps the mix instead of rounding the probabilities down
    from vampire_b200.view_transform import LiftRenderB200
    mod = LiftRenderB200(plans="off", **case.conf).cuda()
    with torch.no_grad():
        got = mod.lift_pool(case.depth.cuda(), ctx16.cuda(), case.mats)
    assert got.dtype == torch.float32 and torch.equal(got, out.detach())


# ---- camera-march plan ------------------------------------------------------------------------------------------
def _render_batch(case, st):
    from vampire_b200.plan import LiftPlanBatch, build_render_plans
    plans = build_render_plans(st, case.prep.cuda(), True)
    return plans, LiftPlanBatch(plans, torch.device("cuda", torch.cuda.current_device()))


def test_render_plan_holds_the_strict_indices(case):
    """steps / delta / last == the mask, base voxel, fractions of the bit-exact index kernel (vb200_render_indices,
    SHA-pinned to the reference in test_gpu_geometry.py) and the step lengths of the bit-exact geometry."""
    ops, cid, st = _state(case.cfg)
    cfg = case.cfg
    prep = case.prep.cuda()
    mask, i0, frac = ops.render_indices(prep, cid, True)
    geom = torch.nan_to_num(ops.get_geometry(prep, cid, True, False), -1e3)
    dl_ref = torch.norm(geom[:, :, 1:] - geom[:, :, :-1], dim=-1).cpu().numpy()          # (B,N,S,fH,fW)
    plans, _ = _render_batch(case, st)
    S, fH, fW, N = cfg.S, cfg.fH, cfg.fW, cfg.num_cams
    pw, ph = 8, 4
    npx, npy = -(-fW // pw), -(-fH // ph)
    lane = np.arange(32)
    for b, p in enumerate(plans):
        steps = p.steps.cpu().numpy().view(np.uint32).reshape(N, npy, npx, S, 32, 4)
        delta = p.delta.cpu().numpy().reshape(N, npy, npx, S, 32)
        last = p.last.cpu().numpy().reshape(N, npy, npx, 32)
        m = mask[b].cpu().numpy().astype(bool)                                            # (N,S,fH,fW)
        ii = i0[b].cpu().numpy().astype(np.int64)
        ff = frac[b].cpu().numpy()
        for py in range(npy):
            for px in range(npx):
                w = px * pw + lane % pw
                h = py * ph + lane // pw
                act = (w < fW) & (h < fH)
                wc, hc = np.minimum(w, fW - 1), np.minimum(h, fH - 1)
                rec = steps[:, py, px]                                                    # (N,S,32,4)
                valid = (rec[..., 0] >> 31).astype(bool)
                want = m[:, :, hc, wc] & act                                              # (N,S,32)
                assert np.array_equal(valid, want)
                x0, y0, z0 = (ii[:, :, hc, wc, a] for a in range(3))
                sx, sy, sz = np.minimum(x0, cfg.vX - 2), np.minimum(y0, cfg.vY - 2), np.minimum(z0, cfg.vZ - 2)
                v0 = (sz * cfg.vY + sy) * cfg.vX + sx
                assert np.array_equal((rec[..., 0] & 0x1FFFFF)[valid], v0[valid])
                fr = rec[..., 1:].view(np.float32)
                exp = ff[:, :, hc, wc, :] + np.stack([x0 - sx, y0 - sy, z0 - sz], -1).astype(np.float32)
                assert np.array_equal(fr[valid], exp[valid])
                np.testing.assert_allclose(delta[:, py, px][:, :, act], dl_ref[b][:, :, hc, wc][:, :, act], rtol=2e-6)
                lv = np.where(want.any(1), S - 1 - np.argmax(want[:, ::-1], axis=1), -1)
                assert np.array_equal(last[:, py, px], lv)


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_planned_march_matches_the_recomputing_march(case, dtype):
    ops, cid, st = _state(case.cfg)
    _, batch = _render_batch(case, st)
    vols = [t.to(dtype).cuda() for t in (case.den, case.sem, case.rgb, case.feat)]
    beta = torch.tensor(0.1, device="cuda")
    a = ops.render_fwd(*vols, beta, case.prep.cuda(), None, cid, True, 3)
    b = ops.render_fwd(*vols, beta, case.prep.cuda(), None, cid, True, 3, batch.table)
    for name, x, y in zip(["rgb", "seg", "depth"], a[:3], b[:3]):
        # same samples, same weights; exp(-cumsum) as a running product and the exact tail step lengths move last bits
        assert_close_scaled(y.cpu().numpy(), x.cpu().numpy(), 3e-6, "planned march " + name)
    for x, y in zip(a[3:], b[3:]):
        assert torch.equal(x, y)                    # the BEV branch does not use the plan
    if dtype == torch.float32 and case.inputs_match_golden:
        from helpers import golden_value
        for name, y in zip(["rgb", "seg", "depth"], b[:3]):
            exp, got = golden_value(case.gold, "r_" + name, y.cpu().numpy())
            assert_close_scaled(got, exp, 1e-5, "planned march vs reference " + name)


def test_planned_march_nonfinite_volume_takes_the_nansafe_path(case):
    ops, cid, st = _state(case.cfg)
    _, batch = _render_batch(case, st)
    sem = case.sem.clone()
    sem[0, 3, 2, 5, 7] = float("nan")
    sem[-1, 0, 1, 1, 1] = float("inf")
    vols = [t.cuda() for t in (case.den, sem, case.rgb, case.feat)]
    beta = torch.tensor(0.1, device="cuda")
    a = ops.render_fwd(*vols, beta, case.prep.cuda(), None, cid, True, 1)
    b = ops.render_fwd(*vols, beta, case.prep.cuda(), None, cid, True, 1, batch.table)
    for x, y in zip(a[:3], b[:3]):
        assert torch.equal(x, y) and torch.isfinite(y).all()


def test_module_render_uses_cached_plans(case):
    from vampire_b200.view_transform import LiftRenderB200
    mod = LiftRenderB200(plans="eval", **case.conf).cuda().eval()
    vols = [t.cuda() for t in (case.den, case.sem, case.feat, case.rgb)]
    with torch.no_grad():
        a = mod.render(case.mats, *vols)
        b = mod.render(case.mats, *vols)
    assert mod.plan_cache.misses == case.batch and mod.plan_cache.hits == case.batch
    for x, y in zip(a, b):
        assert torch.equal(x, y)
    off = LiftRenderB200(plans="off", **case.conf).cuda().eval()
    with torch.no_grad():
        c = off.render(case.mats, *vols)
    for x, y in zip(a[:3], c[:3]):
        assert_close_scaled(x.cpu().numpy(), y.cpu().numpy(), 3e-6, "module planned render")


_STAGED_SCRIPT = r"""
import sys, torch
sys.path.insert(0, sys.argv[1]); sys.path.insert(0, sys.argv[1] + "/tests")
from helpers import Case
from vampire_b200 import ops
from vampire_b200.plan import PlanCache
case = Case(sys.argv[2])
cid = ops.register_config(case.cfg)
dt = {"bf16": torch.bfloat16, "fp16": torch.float16}[sys.argv[3]]
prep = case.prep.cuda()
tab = PlanCache().render(ops.state(cid), cid, prep, True).table
outs = ops.render_fwd(case.den.cuda().to(dt), case.sem.cuda().to(dt), case.rgb.cuda().to(dt), case.feat.cuda().to(dt),
                      torch.tensor(0.1, device="cuda"), prep, None, cid, True, 1, tab)
torch.cuda.synchronize()
torch.save([o.cpu() for o in outs[:3]], sys.argv[4])
"""


@pytest.mark.parametrize("name,dtype", [("mini_stress", "bf16"), ("mini_val", "fp16"), ("r50_val_digest", "bf16"),
                                        ("r50_stress_digest", "bf16")])
def test_staged_march_is_bit_identical(name, dtype, tmp_path):
    """VB200_MARCH_STAGED=1: the march whose voxel boxes are staged in shared memory by bulk asynchronous copies
    (cp.async.bulk + mbarrier, north-star kernel (c)) composites exactly what the direct-gather planned march does."""
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    outs = {}
    for tag in ("direct", "staged"):
        path = str(tmp_path / (tag + ".pt"))
        e = dict(os.environ)
        e.pop("VB200_MARCH_STAGED", None)
        if tag == "staged":
            e["VB200_MARCH_STAGED"] = "1"
        subprocess.run([sys.executable, "-c", _STAGED_SCRIPT, root, name, dtype, path], check=True, env=e, timeout=300)
        outs[tag] = torch.load(path)
    for n, a, b in zip(["rgb", "seg", "depth"], outs["direct"], outs["staged"]):
        assert torch.equal(a, b), n


# ---- depth-split march (thread-block clusters, vb200_render_set_march_split) ------------------------------------
@pytest.mark.parametrize("planned", [False, True])
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_depth_split_march_matches_the_unsplit_march(case, dtype, planned):
    """Clusters of 2 / 4 / 8 CTAs composite one part of the samples each and fold the parts front to back through
    distributed shared memory: same samples, same weights as the single-CTA march, another summation order."""
    from vampire_b200 import cabi
    ops, cid, st = _state(case.cfg)
    table = _render_batch(case, st)[1].table if planned else None
    vols = [t.to(dtype).cuda() for t in (case.den, case.sem, case.rgb, case.feat)]
    beta = torch.tensor(0.1, device="cuda")
    prep = case.prep.cuda()
    try:
        cabi.render_set_march_split(1)
        ref = [x.cpu().numpy() for x in ops.render_fwd(*vols, beta, prep, None, cid, True, 1, table)[:3]]
        for nseg in (2, 4, 8, -1):
            cabi.render_set_march_split(nseg)
            got = ops.render_fwd(*vols, beta, prep, None, cid, True, 1, table)[:3]
            for name, x, y in zip(["rgb", "seg", "depth"], ref, got):
                assert_close_scaled(y.cpu().numpy(), x, 3e-6, f"depth-split march {name}")
            if dtype == torch.float32 and case.inputs_match_golden and nseg == 8:
                from helpers import golden_value
                for name, y in zip(["rgb", "seg", "depth"], got):
                    exp, g = golden_value(case.gold, "r_" + name, y.cpu().numpy())
                    assert_close_scaled(g, exp, 1e-5, "depth-split march vs reference " + name)
    finally:
        cabi.render_set_march_split(0)
    with pytest.raises(RuntimeError):
        cabi.render_set_march_split(3)
