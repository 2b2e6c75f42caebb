"""GPU: API robustness of the ops / module -- camera counts, non-contiguous inputs, error behaviour."""
from dataclasses import replace

import pytest
import torch

from helpers import Case, assert_close_scaled
from oracle import torch_path as tp

pytestmark = pytest.mark.gpu


def test_four_camera_rig_matches_oracle():
    from vampire_b200 import ops, synth
    from vampire_b200.config import MINI
    from vampire_b200.matrices import prepare_matrices
    cfg = replace(MINI, num_cams=4)
    conf = cfg.backbone_kwargs()
    cid = ops.register_config(cfg)
    m = synth.make_mats(cfg, 2, "stress", seed=31)
    prep = prepare_matrices(m["sensor2ego_mats"][:, 0], m["intrin_mats"][:, 0], m["ida_mats"][:, 0], m["bda_mat"])
    depth, ctx = synth.make_lift_inputs(cfg, 2, seed=31)
    den, sem, feat, rgb = synth.make_render_inputs(cfg, 2, seed=31)
    buf = tp.build_buffers(conf)
    with torch.no_grad():
        ref_vox = tp.lift_pool(conf, buf, depth, ctx, m)
        ref = tp.render_from_mats(conf, buf, m, den, sem, feat, rgb, torch.tensor(0.1))
    vox, _ = ops.lift_pool_fwd(depth.cuda(), ctx.cuda(), prep.cuda(), cid, True, False, False)
    assert_close_scaled(vox.cpu().numpy(), ref_vox.numpy(), 1e-5, "vox, 4 cameras")
    outs = ops.render_fwd(den.cuda(), sem.cuda(), rgb.cuda(), feat.cuda(), torch.tensor(0.1, device="cuda"),
                          prep.cuda(), None, cid, True, 3)
    for o, r in zip(outs, ref):
        assert_close_scaled(o.cpu().numpy(), r.numpy(), 1e-5, "render, 4 cameras")


def test_non_contiguous_inputs_are_accepted():
    case = Case("mini_val")
    from vampire_b200 import ops
    cid = ops.register_config(case.cfg)
    depth = case.depth.cuda().permute(0, 1, 3, 4, 2).contiguous().permute(0, 1, 4, 2, 3)   # strided view
    ctx = case.ctx.cuda().flip(-1).flip(-1)
    assert not depth.is_contiguous()
    a, _ = ops.lift_pool_fwd(depth, ctx, case.prep.cuda(), cid, True, False, False)
    b, _ = ops.lift_pool_fwd(case.depth.cuda(), case.ctx.cuda(), case.prep.cuda(), cid, True, False, False)
    assert torch.equal(a, b)


def test_errors_are_loud():
    case = Case("mini_val")
    from vampire_b200 import ops
    cid = ops.register_config(case.cfg)
    d, c, p = case.depth.cuda(), case.ctx.cuda(), case.prep.cuda()
    with pytest.raises(ValueError):
        ops.lift_pool_fwd(d[:, :, :-1], c, p, cid, True, False, False)          # wrong depth planes
    with pytest.raises(TypeError):
        ops.lift_pool_fwd(d.half(), c, p, cid, True, False, False)              # dtype mismatch (16-bit depth, fp32 ctx)
    with pytest.raises(ValueError):
        ops.lift_pool_fwd(d, c, p[:, :, :5], cid, True, False, False)           # not a prepare_matrices block
    with pytest.raises(TypeError):
        ops.lift_pool_fwd(d.double(), c.double(), p, cid, True, False, False)   # unsupported dtype
    with pytest.raises(RuntimeError, match="CUDA"):
        ops.lift_pool_fwd(d.cpu(), c.cpu(), p.cpu(), cid, True, False, False)   # no CPU path
    den, sem, feat, rgb = (t.cuda() for t in (case.den, case.sem, case.feat, case.rgb))
    with pytest.raises(ValueError):
        ops.render_fwd(den, sem[:, :-1], rgb, feat, torch.tensor(0.1, device="cuda"), p, None, cid, True, 3)
    # forward without saved counts cannot be differentiated
    dd = d.clone().requires_grad_(True)
    out, _ = ops.lift_pool_fwd(dd, c, p, cid, True, False, False)
    with pytest.raises(RuntimeError, match="save_cnt"):
        out.sum().backward()


def test_trace_and_launch_counters():
    from vampire_b200 import cabi, ops
    case = Case("mini_val")
    cid = ops.register_config(case.cfg)
    n0 = cabi.launch_count()
    cabi.trace_enable(True)
    ops.lift_pool_fwd(case.depth.cuda(), case.ctx.cuda(), case.prep.cuda(), cid, True, False, False)
    tr = cabi.trace_collect()
    cabi.trace_enable(False)
    assert cabi.launch_count() - n0 == 2
    assert set(tr) == {"ctx_to_nhwc", "lift_pool_fwd"} and all(ms > 0 and n == 1 for ms, n in tr.values())


def test_forward_is_cuda_graph_capturable():
    """The ops allocate only from the caching allocator and launch on the current stream, so a whole
    lift+pool+render forward (on prepared matrices) can be captured into a CUDA graph and replayed -- the BEV
    side-stream fork is skipped under capture; replays must reproduce the eager results bit for bit."""
    from vampire_b200 import ops
    case = Case("mini_val")
    cid = ops.register_config(case.cfg)
    prep = case.prep.cuda()
    depth, ctx = case.depth.cuda(), case.ctx.cuda()
    den, sem, feat, rgb = case.den.cuda(), case.sem.cuda(), case.feat.cuda(), case.rgb.cuda()
    beta = torch.tensor(0.1, device="cuda")

    def step():
        vox, _ = ops.lift_pool_fwd(depth, ctx, prep, cid, True, False, False)
        return [vox] + list(ops.render_fwd(den, sem, rgb, feat, beta, prep, None, cid, True, 3))

    with torch.no_grad():
        eager = step()
        s = torch.cuda.Stream()          # warm-up on a side stream (allocator + lazy state), then capture
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            step()
        torch.cuda.current_stream().wait_stream(s)
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            captured = step()
        for _ in range(2):
            for t in captured:
                t.zero_()
            graph.replay()
            torch.cuda.synchronize()
            for a, b in zip(eager, captured):
                assert torch.equal(a, b)


def test_depth_softmax_and_2d_lift_reject_bad_input():
    from vampire_b200 import ops
    from vampire_b200.view_transform import LiftRenderB200
    case = Case("mini_val")
    mod = LiftRenderB200(**case.conf).cuda()
    with pytest.raises(RuntimeError, match="CUDA"):
        mod.depth_softmax(torch.zeros(2, 4, 4, 4))
    with pytest.raises(ValueError):
        ops.depth_softmax_fwd(torch.zeros(4, 4, device="cuda"), True)
    with pytest.raises(ValueError, match="do not match"):
        mod.lift_pool_2d(torch.zeros(1, case.cfg.num_cams, case.cfg.C, 3, 3, device="cuda"), case.mats)


@pytest.mark.parametrize("seed", [11, 12])
def test_module_api_with_host_matrices_vs_same_host_oracle(seed):
    """The e2e path: ``LiftRenderB200.lift_pool / .render`` handed a HOST mats_dict prepare the 4x4 matrices with this
    host's LAPACK -- exactly what the reference running on this host does -- so module and oracle must agree here
    whatever the host rounds the inverses to (the fixtures' prepared matrices are not involved)."""
    from oracle import torch_path as tp
    from vampire_b200 import synth
    from vampire_b200.config import MINI
    from vampire_b200.view_transform import LiftRenderB200
    cfg, conf = MINI, MINI.backbone_kwargs()
    mats = synth.make_mats(cfg, 2, "stress", seed=seed)
    depth, ctx = synth.make_lift_inputs(cfg, 2, seed=seed)
    den, sem, feat, rgb = synth.make_render_inputs(cfg, 2, seed=seed, field="surface")
    buf = tp.build_buffers(conf)
    with torch.no_grad():
        ref_vox = tp.lift_pool(conf, buf, depth, ctx, mats)
        ref = tp.render_from_mats(conf, buf, mats, den, sem, feat, rgb, torch.tensor(0.1))
    for plans in ("off", "always"):
        mod = LiftRenderB200(plans=plans, **conf).cuda().eval()
        with torch.no_grad():
            vox = mod.lift_pool(depth.cuda(), ctx.cuda(), mats)
            outs = mod.render(mats, den.cuda(), sem.cuda(), feat.cuda(), rgb.cuda())
        assert_close_scaled(vox.cpu().numpy(), ref_vox.numpy(), 1e-5, "module lift_pool (host mats_dict)")
        for n, o, r in zip(["rgb", "seg", "depth", "bev_rgb", "bev_seg", "bev_height", "voxel_density", "voxel_output"],
                           outs, ref):
            assert_close_scaled(o.cpu().numpy(), r.numpy(), 1e-5, "module render " + n)


@pytest.mark.filterwarnings("ignore:The AccumulateGrad node's stream does not match")
def test_graphed_train_step_matches_the_eager_step():
    """dp.GraphedTrainStep (two CUDA graphs around the all-reduce call) replays exactly dp.train_step: same outputs,
    same gradients, also after the producer has written new inputs and new matrices into the captured tensors."""
    from helpers import Case
    from vampire_b200 import synth
    from vampire_b200.dp import GradBucket, GraphedTrainStep, train_step
    from vampire_b200.view_transform import LiftRenderB200
    case = Case("mini_stress")
    cfg = case.cfg
    mod = LiftRenderB200(**case.conf).cuda().train()
    prep = case.prep.cuda().clone()
    leaves = [t.clone().cuda().requires_grad_(True) for t in (case.depth, case.ctx, case.den, case.sem, case.feat, case.rgb)]
    B = case.batch
    shapes = [(B, cfg.C, cfg.vZ, cfg.vY, cfg.vX), (B, cfg.num_cams, 3, cfg.fH, cfg.fW), (B, cfg.num_cams, cfg.K, cfg.fH, cfg.fW),
              (B, cfg.num_cams, 1, cfg.fH, cfg.fW), (B, 3, cfg.oY, cfg.oX), (B, cfg.K, cfg.oY, cfg.oX), (B, 1, cfg.oY, cfg.oX),
              (B, 1, cfg.oZ, cfg.oY, cfg.oX), (B, cfg.C, cfg.oZ, cfg.oY, cfg.oX)]
    cots = [c.cuda() for c in synth.make_cotangents(shapes)]
    bucket = GradBucket(torch.device("cuda"), 1)
    d, c, den, sem, feat, rgb = leaves
    graphed = GraphedTrainStep(mod, d, c, (den, sem, feat, rgb), prep, cots, bucket)

    def snapshot(vox, rend):
        return ([vox.detach().clone()] + [r.detach().clone() for r in rend],
                [t.grad.detach().clone() for t in leaves + [mod.density.beta]])

    for round_ in range(2):
        if round_ == 1:      # the next batch: new feature values and the other rig's matrices, written in place
            with torch.no_grad():
                for t in leaves:
                    t.mul_(0.5).add_(0.01)
                prep.copy_(Case("mini_val").prep.cuda())
        got = snapshot(*graphed())
        ref = snapshot(*train_step(mod, d, c, (den, sem, feat, rgb), prep, cots, bucket))
        for a, b in zip(got[0], ref[0]):
            assert torch.equal(a, b)
        for name, a, b in zip(("depth", "ctx", "den", "sem", "feat", "rgb", "beta"), got[1], ref[1]):
            if name in ("depth", "ctx"):
                assert torch.equal(a, b), name                   # deterministic lift backward
            else:                                                # vector atomics: order-dependent last bits
                assert torch.allclose(a, b, rtol=1e-4, atol=1e-5 * float(b.abs().max())), name
