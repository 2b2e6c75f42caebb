"""GPU: L1-L4 lift+pool and R1-R6 render forward through the C ABI / custom ops, against the
reference's golden outputs and the live torch oracle."""
import os

import numpy as np
import pytest
import torch

from helpers import CASES, Case, assert_close_scaled, golden_value
from oracle import torch_path as tp

pytestmark = pytest.mark.gpu

FP32_REL = 1e-5     # north star: 1e-5 relative in fp32
BF16_REL = 1e-2     # 1e-2 with bf16 features
NAMES = ["rgb", "seg", "depth", "bev_rgb", "bev_seg", "bev_height", "voxel_density", "voxel_output"]


@pytest.fixture(scope="module", params=list(CASES))
def case(request):
    return Case(request.param)


def _ops(cfg):
    from vampire_b200 import ops
    return ops, ops.register_config(cfg)


@pytest.mark.parametrize("channels_last", [False, True])
def test_lift_pool_fp32_vs_reference(case, channels_last):
    if not case.inputs_match_golden:
        pytest.skip("inputs differ from the fixture's")
    ops, cid = _ops(case.cfg)
    out, cnt = ops.lift_pool_fwd(case.depth.cuda(), case.ctx.cuda(), case.prep.cuda(), cid, True, channels_last, True)
    assert out.shape == (case.batch, case.cfg.C, case.cfg.vZ, case.cfg.vY, case.cfg.vX)
    assert out.permute(0, 2, 3, 4, 1).is_contiguous() == channels_last
    exp, got = golden_value(case.gold, "vox", out.contiguous().cpu().numpy())
    assert_close_scaled(got, exp, FP32_REL, "pooled voxel features")


def test_lift_pool_determinism(case):
    ops, cid = _ops(case.cfg)
    args = (case.depth.cuda(), case.ctx.cuda(), case.prep.cuda(), cid, True, False, False)
    a = ops.lift_pool_fwd(*args)[0]
    for _ in range(3):
        assert torch.equal(a, ops.lift_pool_fwd(*args)[0])


def test_lift_count_matches_oracle():
    """Per-channel non-zero camera count (BV2:509-512) incl. exact zeros in ctx."""
    from oracle import strict_np as sn
    case = Case("mini_stress")
    ops, cid = _ops(case.cfg)
    ctx = case.ctx.clone()
    ctx[:, :, 3] = 0.0           # a dead channel: count must be 0 -> output 0/(0+1e-6) = 0
    ctx[:, 1, 5] = 0.0           # dead in one camera only
    out, cnt = ops.lift_pool_fwd(case.depth.cuda(), ctx.cuda(), case.prep.cuda(), cid, True, False, True)
    lat, cfg = case.lat, case.cfg
    pix = sn.project_voxels(case.gold["prep"], lat.xs.numpy(), lat.ys.numpy(), lat.zs.numpy())
    ref, rcnt = sn.lift_pool_factorised(case.depth.numpy(), ctx.numpy(), pix, cfg.final_dim, cfg.d_bound)
    cnt = cnt.cpu().numpy().astype(np.uint64).reshape(case.batch, cfg.vZ, cfg.vY, cfg.vX)
    got = np.stack([(cnt >> np.uint64(4 * c)) & np.uint64(0xF) for c in range(cfg.C)], 1).astype(np.int32)
    assert np.array_equal(got, rcnt)
    assert_close_scaled(out.cpu().numpy(), ref, FP32_REL, "pooled features with dead channels")
    assert (out[:, 3] == 0).all()


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16])
def test_lift_pool_half_features(case, dtype):
    """bf16/fp16-valued features, fp32 arithmetic; oracle fed the same rounded inputs."""
    ops, cid = _ops(case.cfg)
    d, c = case.depth.to(dtype), case.ctx.to(dtype)
    out, _ = ops.lift_pool_fwd(d.cuda(), c.cuda(), case.prep.cuda(), cid, True, False, False)
    assert out.dtype == dtype
    buf = tp.build_buffers(case.conf)
    with torch.no_grad():
        ref = tp.lift_pool(case.conf, buf, d.float(), c.float(), case.mats)
    assert_close_scaled(out.float().cpu().numpy(), ref.numpy(), BF16_REL, f"lift {dtype}")


@pytest.mark.parametrize("from_tensor", [False, True])
def test_render_fp32_vs_reference(case, from_tensor):
    if not case.inputs_match_golden:
        pytest.skip("inputs differ from the fixture's")
    ops, cid = _ops(case.cfg)
    prep = case.prep.cuda()
    geom = ops.get_geometry(prep, cid, True, True) if from_tensor else None
    beta = torch.tensor(0.1, device="cuda")
    outs = ops.render_fwd(case.den.cuda(), case.sem.cuda(), case.rgb.cuda(), case.feat.cuda(), beta, prep, geom, cid,
                          True, 3)
    for n, o in zip(NAMES, outs):
        exp, got = golden_value(case.gold, "r_" + n, o.cpu().numpy())
        assert_close_scaled(got, exp, FP32_REL, n)


def test_render_early_termination_is_within_tolerance():
    """term_eps only skips samples whose weight is below 1e-8."""
    case = Case("mini_val")
    ops, cid = _ops(case.cfg)
    st = ops.state(cid)
    args = (case.den.cuda(), case.sem.cuda(), case.rgb.cuda(), case.feat.cuda(), torch.tensor(0.1, device="cuda"),
            case.prep.cuda(), None, cid, True, 1)
    old = st.term_eps
    try:
        st.term_eps = 0.0
        full = [o.clone() for o in ops.render_fwd(*args)]
        st.term_eps = 1e-8
        cut = ops.render_fwd(*args)
    finally:
        st.term_eps = old
    for n, a, b in zip(NAMES[:3], full, cut):
        assert_close_scaled(b.cpu().numpy(), a.cpu().numpy(), 1e-6, n)


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16])
def test_render_half_features(case, dtype):
    ops, cid = _ops(case.cfg)
    vols = [t.to(dtype) for t in (case.den, case.sem, case.rgb, case.feat)]
    beta = torch.tensor(0.1, device="cuda")
    outs = ops.render_fwd(*[v.cuda() for v in vols], beta, case.prep.cuda(), None, cid, True, 3)
    buf = tp.build_buffers(case.conf)
    with torch.no_grad():
        ref = tp.render_from_mats(case.conf, buf, case.mats, vols[0].float(), vols[1].float(), vols[3].float(),
                                  vols[2].float(), torch.tensor(0.1))
    assert outs[7].dtype == dtype
    for n, o, r in zip(NAMES, outs, ref):
        assert_close_scaled(o.float().cpu().numpy(), r.numpy(), BF16_REL, f"{n} {dtype}")


def test_module_interface_matches_reference_signatures():
    """LiftRenderB200 mirrors the backbone methods: same arguments, same return shapes/order."""
    from vampire_b200.view_transform import LiftRenderB200
    case = Case("mini_val")
    mod = LiftRenderB200(**case.conf).cuda()
    assert "density.beta" in mod.state_dict()
    a = case.mat_args()
    geom = mod.get_geometry(*a)
    pix = mod.get_pixel(*a)
    c = case.cfg
    assert geom.shape == (case.batch, 6, c.D, c.fH, c.fW, 3) and pix.shape == (case.batch, 6, c.vZ, c.vY, c.vX, 3)
    vox = mod.lift_pool(case.depth.cuda(), case.ctx.cuda(), case.mats)
    exp, got = golden_value(case.gold, "vox", vox.cpu().numpy())
    assert_close_scaled(got, exp, FP32_REL, "module lift_pool")
    vols = [t.cuda() for t in (case.den, case.sem, case.feat, case.rgb)]
    fused = mod.render(case.mats, *vols)
    ref_sig = mod.volume_rendering_from_multiple_views(torch.nan_to_num(geom, -1e3), *vols)
    assert len(fused) == len(ref_sig) == 8
    for n, x, y in zip(NAMES, fused, ref_sig):
        # same kernels; the fused path may finish rays that left the volume with a geometry-free tail
        assert_close_scaled(x.detach().cpu().numpy(), y.detach().cpu().numpy(), 1e-6, n + " fused vs signature")
        exp, got = golden_value(case.gold, "r_" + n, x.detach().cpu().numpy())
        assert_close_scaled(got, exp, FP32_REL, n)


ODD_CONFIGS = {
    # det grid coarser than the seg grid in xy: the BEV window assumption fails -> scalar gathers
    "coarse_det": dict(x_bound_det=(-51.2, 51.2, 6.4), y_bound_det=(-51.2, 51.2, 6.4)),
    # vX, oX not multiples of 4 -> the scalar BEV kernel; ragged feature map (fW=44 is not a multiple of 8)
    "ragged": dict(x_bound_seg=(-48.0, 48.0, 3.2), x_bound_det=(-48.0, 48.0, 3.2)),
    # det z-range not aligned with seg levels, fewer levels
    "shifted_z": dict(z_bound_det=(-2.2, 2.6, 1.6)),
    # BASELINE configs[3] in miniature: the frustum scaled 2x (128x352 input -> 32x88 feature map)
    "scaled_frustum": dict(final_dim=(128, 352)),
    # BASELINE configs[4] in miniature: more samples per ray (D = 129 planes, 128 samples)
    "dense_samples": dict(d_bound=(2.0, 58.0, 0.4375)),
    # fine seg z (40 rows) under coarse det levels: more than 17 z-rows touched -> BEV direct fallback
    "fine_seg_z": dict(z_bound_seg=(-5.0, 3.0, 0.2), z_bound_det=(-4.4, 2.8, 1.2)),
}


@pytest.mark.parametrize("name", list(ODD_CONFIGS))
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_odd_grids_vs_live_oracle(name, dtype):
    """Edge geometries the reference supports through its config: every kernel path (vectorised /
    scalar BEV, ragged ray patches) against the torch oracle."""
    from dataclasses import replace
    from vampire_b200 import synth
    from vampire_b200.config import MINI
    from vampire_b200.matrices import prepare_matrices
    cfg = replace(MINI, **ODD_CONFIGS[name])
    conf = cfg.backbone_kwargs()
    ops, cid = _ops(cfg)
    B = 2
    mats = synth.make_mats(cfg, B, "stress", seed=11)
    prep = prepare_matrices(mats["sensor2ego_mats"][:, 0], mats["intrin_mats"][:, 0], mats["ida_mats"][:, 0],
                            mats["bda_mat"])
    depth, ctx = synth.make_lift_inputs(cfg, B, seed=11, dtype=dtype)
    den, sem, feat, rgb = synth.make_render_inputs(cfg, B, seed=11, field="surface", dtype=dtype)
    buf = tp.build_buffers(conf)
    with torch.no_grad():
        ref_vox = tp.lift_pool(conf, buf, depth.float(), ctx.float(), mats)
        ref = tp.render_from_mats(conf, buf, mats, den.float(), sem.float(), feat.float(), rgb.float(),
                                  torch.tensor(0.1))
    rel = FP32_REL if dtype == torch.float32 else BF16_REL
    vox, _ = ops.lift_pool_fwd(depth.cuda(), ctx.cuda(), prep.cuda(), cid, True, False, False)
    assert_close_scaled(vox.float().cpu().numpy(), ref_vox.numpy(), rel, "vox")
    outs = ops.render_fwd(den.cuda(), sem.cuda(), rgb.cuda(), feat.cuda(), torch.tensor(0.1, device="cuda"),
                          prep.cuda(), None, cid, True, 3)
    for n, o, r in zip(NAMES, outs, ref):
        assert_close_scaled(o.float().cpu().numpy(), r.numpy(), rel, n)


def test_render_group_sizes_agree():
    """Packing/marching 1 sample per round (L2-resident) or all at once must not change results."""
    case = Case("mini_stress")
    ops, cid = _ops(case.cfg)
    st = ops.state(cid)
    args = (case.den.cuda(), case.sem.cuda(), case.rgb.cuda(), case.feat.cuda(), torch.tensor(0.1, device="cuda"),
            case.prep.cuda(), None, cid, True, 3)
    old = st.render_group
    try:
        st.render_group = 1
        a = [o.clone() for o in ops.render_fwd(*args)]
        st.render_group = 8
        b = ops.render_fwd(*args)
    finally:
        st.render_group = old
    for x, y in zip(a, b):
        assert torch.equal(x, y)


def test_get_voxel_feats_reference_signature():
    """The compatibility method takes the materialised 6-D frustum tensor like BV2:483 and agrees with
    both the reference and the fused lift_pool; its autograd matches the oracle's."""
    from vampire_b200.view_transform import LiftRenderB200
    case = Case("mini_stress")
    mod = LiftRenderB200(**case.conf).cuda()
    depth, ctx = case.depth.cuda(), case.ctx.cuda()
    fr = (depth.unsqueeze(2) * ctx.unsqueeze(3)).requires_grad_(True)
    out = mod.get_voxel_feats(fr, 0, case.mats)
    assert_close_scaled(out.detach().cpu().numpy(), case.gold["vox"], FP32_REL, "get_voxel_feats")
    fused = mod.lift_pool(depth, ctx, case.mats)
    assert_close_scaled(out.detach().cpu().numpy(), fused.cpu().numpy(), FP32_REL, "compat vs fused")
    cot = case.cotangents()[0]
    g = torch.autograd.grad((out * cot.cuda()).sum(), fr)[0]
    buf = tp.build_buffers(case.conf)
    fr_ref = (case.depth.unsqueeze(2) * case.ctx.unsqueeze(3)).requires_grad_(True)
    ref = tp.get_voxel_feats(case.conf, buf, fr_ref, case.mats)
    g_ref = torch.autograd.grad((ref * cot).sum(), fr_ref)[0]
    assert_close_scaled(g.cpu().numpy(), g_ref.numpy(), 2e-5, "d_frustum")


@pytest.mark.parametrize("seed", [200, 201, 202])
def test_fused_kernels_use_the_exact_indices_random_rigs(seed):
    """The conservative culls (per warp / per voxel) and the identity-bda shortcut must never drop a pair
    the strict projection calls valid: pooled features on random stress rigs vs the live oracle."""
    from vampire_b200 import synth
    from vampire_b200.config import MINI
    from vampire_b200.matrices import prepare_matrices
    cfg, conf = MINI, MINI.backbone_kwargs()
    ops, cid = _ops(cfg)
    m = synth.make_mats(cfg, 2, "stress", seed=seed)
    if seed == 202:
        m["bda_mat"] = torch.eye(4).expand(2, 4, 4).contiguous()     # identity-bda fast path
    prep = prepare_matrices(m["sensor2ego_mats"][:, 0], m["intrin_mats"][:, 0], m["ida_mats"][:, 0], m["bda_mat"])
    depth, ctx = synth.make_lift_inputs(cfg, 2, seed=seed)
    den, sem, feat, rgb = synth.make_render_inputs(cfg, 2, seed=seed, field="surface")
    buf = tp.build_buffers(conf)
    with torch.no_grad():
        ref_vox = tp.lift_pool(conf, buf, depth, ctx, m)
        ref = tp.render_from_mats(conf, buf, m, den, sem, feat, rgb, torch.tensor(0.1))
    vox, cnt = ops.lift_pool_fwd(depth.cuda(), ctx.cuda(), prep.cuda(), cid, True, False, True)
    assert_close_scaled(vox.cpu().numpy(), ref_vox.numpy(), FP32_REL, "vox")
    # the saved camera counts equal the number of strictly valid cameras per voxel
    valid, _, _ = ops.lift_indices(prep.cuda(), cid, True)
    nvalid = valid.sum(1).reshape(2, -1).cpu().numpy()
    got = (cnt.cpu().numpy().astype(np.uint64) & np.uint64(0xF)).astype(np.int64)
    assert np.array_equal(got, nvalid)
    outs = ops.render_fwd(den.cuda(), sem.cuda(), rgb.cuda(), feat.cuda(), torch.tensor(0.1, device="cuda"),
                          prep.cuda(), None, cid, True, 3)
    for n, o, r in zip(NAMES, outs, ref):
        assert_close_scaled(o.cpu().numpy(), r.numpy(), FP32_REL, n)


_TMA_SCRIPT = r"""
import sys, torch
sys.path.insert(0, sys.argv[1]); sys.path.insert(0, sys.argv[1] + "/tests")
from helpers import Case
from vampire_b200 import ops
case = Case(sys.argv[2])
cid = ops.register_config(case.cfg)
dt = {"fp32": torch.float32, "bf16": torch.bfloat16}[sys.argv[3]]
outs = ops.render_fwd(case.den.cuda().to(dt), case.sem.cuda().to(dt), case.rgb.cuda().to(dt), case.feat.cuda().to(dt),
                      torch.tensor(0.1, device="cuda"), case.prep.cuda(), None, cid, True, 2)
torch.save([o.cpu() for o in outs], sys.argv[4])
"""


@pytest.mark.parametrize("name,dtype", [("mini_stress", "fp32"), ("r50_val_digest", "bf16")])
def test_bev_tma_staged_kernel_is_bit_identical(name, dtype, tmp_path):
    """The opt-in TMA-staged BEV kernel (VB200_BEV_TMA=1: cp.async.bulk + mbarrier ring) does the same arithmetic
    in the same order as the default direct-load kernel: its outputs must be bit-identical."""
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    outs = {}
    for tag, env in (("direct", {}), ("tma", {"VB200_BEV_TMA": "1"})):
        path = str(tmp_path / (tag + ".pt"))
        e = dict(os.environ, **env)
        e.pop("VB200_BEV_TMA", None) if not env else None
        subprocess.run([sys.executable, "-c", _TMA_SCRIPT, root, name, dtype, path], check=True, env=e, timeout=300)
        outs[tag] = torch.load(path)
    for n, a, b in zip(NAMES, outs["direct"], outs["tma"]):
        if a.numel():
            assert torch.equal(a, b), n


def test_render_nonfinite_volume_takes_the_nan_safe_march():
    """A NaN / inf anywhere in the camera-branch volume raises the pack's flag and the march keeps
    torch.nan_to_num of the interpolated features (BV2:421); results must still match the reference ops."""
    case = Case("mini_val")
    ops, cid = _ops(case.cfg)
    den, sem, rgb, feat = case.den.clone(), case.sem.clone(), case.rgb.clone(), case.feat.clone()
    g = torch.Generator().manual_seed(5)
    for t, v in ((sem, float("nan")), (rgb, float("inf")), (sem, float("-inf"))):
        idx = torch.randint(0, t.numel(), (40,), generator=g)
        t.view(-1)[idx] = v
    buf = tp.build_buffers(case.conf)
    ref = tp.render_from_mats(case.conf, buf, case.mats, den, sem, feat, rgb, torch.tensor(0.1))
    st = ops.state(cid)
    old = st.term_eps
    try:
        st.term_eps = 0.0      # compare every sample: a skipped 1e-9 * FLT_MAX term is not small
        outs = ops.render_fwd(den.cuda(), sem.cuda(), rgb.cuda(), feat.cuda(), torch.tensor(0.1, device="cuda"),
                              case.prep.cuda(), None, cid, True, 1)
    finally:
        st.term_eps = old
    for n, o, r in list(zip(NAMES, outs, ref))[:3]:
        got, exp = o.cpu().numpy(), r.numpy()
        assert np.array_equal(np.isnan(got), np.isnan(exp)), n
        ok = ~np.isnan(exp)
        # weight * +-FLT_MAX sums: the weights agree to ~1e-6 ABSOLUTE (a weight of 1e-7 is one ulp of
        # 1 - exp(-sd) and may differ by 100 % relative), so compare in units of FLT_MAX
        big = ok & (np.abs(exp) > 1e25)
        fmax = float(np.finfo(np.float32).max)
        assert np.allclose(got[big].astype(np.float64) / fmax, exp[big].astype(np.float64) / fmax, rtol=1e-4, atol=2e-6), n
        small = ok & ~big
        assert_close_scaled(got[small], exp[small], FP32_REL, n, scale=max(np.abs(exp[small]).max(), 1e-30))


def _rig(cfg, seed):
    """Random stress rig + variations that push pairs against every validity boundary: doubled ida scale (zoom: the
    x / y image bounds cut through the volume), a camera dropped into the voxel lattice (depths around d_lo),
    mirrored images."""
    from vampire_b200 import synth
    m = synth.make_mats(cfg, 2, "stress", seed=seed)
    if seed % 3 == 1:
        m["ida_mats"][..., :2, :] *= 2.0
    if seed % 3 == 2:
        m["sensor2ego_mats"][:, :, 0, :3, 3] = torch.tensor([0.2, -0.6, -1.0])   # on a voxel centre
        m["ida_mats"][:, :, 1, 0, :] *= -1.0
        m["ida_mats"][:, :, 1, 0, 3] += cfg.final_dim[1] - 1.0
    return m


@pytest.mark.parametrize("lift_2d", [False, True])
@pytest.mark.parametrize("block", range(5))
def test_fused_cull_never_drops_a_valid_pair_50_rigs(block, lift_2d):
    """The fused lift's two-level conservative cull runs on an FMA-composed matrix with a guard band; a dropped valid
    pair would be a silent parity bug.  50 random rigs (MINI): the camera count the fused kernel saved per voxel
    equals the number of cameras the strict index kernel calls valid -- and the cached plan agrees."""
    from vampire_b200.config import MINI
    from vampire_b200.matrices import prepare_matrices
    from vampire_b200.plan import build_lift_plans
    cfg = MINI
    from vampire_b200 import ops
    # lift_2d: the BaseBiLinear mode, whose depth test z > 0 lets voxels right at the camera plane through -- the
    # case in which a guard band that reaches behind the camera once broke the warp-level cull
    cid = ops.register_config(cfg, lift_2d)
    ones_d = torch.ones(2, cfg.num_cams, 1 if lift_2d else cfg.D, cfg.fH, cfg.fW, device="cuda")
    ones_c = torch.ones(2, cfg.num_cams, cfg.C, cfg.fH, cfg.fW, device="cuda")
    total = 0
    for seed in range(300 + 10 * block, 310 + 10 * block):
        m = _rig(cfg, seed)
        prep = prepare_matrices(m["sensor2ego_mats"][:, 0], m["intrin_mats"][:, 0], m["ida_mats"][:, 0],
                                m["bda_mat"]).cuda()
        _, cnt = ops.lift_pool_fwd(ones_d, ones_c, prep, cid, True, False, True)
        valid, _, _ = ops.lift_indices(prep, cid, True)
        nvalid = valid.sum(1, dtype=torch.int64).reshape(2, -1)
        assert torch.equal(cnt & 0xF, nvalid), f"seed {seed}: the cull dropped (or invented) a pair"
        for b, p in enumerate(build_lift_plans(ops.state(cid), prep, True)):
            assert p.num_pairs == int(nvalid[b].sum())
        total += int(nvalid.sum())
    assert total > 0


def test_fused_cull_never_drops_a_valid_pair_full_size():
    """The same property at the full R50 geometry, val-mode and stress rig."""
    from vampire_b200 import synth
    from vampire_b200.config import R50_256x704 as cfg
    from vampire_b200.matrices import prepare_matrices
    ops, cid = _ops(cfg)
    ones_d = torch.ones(1, cfg.num_cams, cfg.D, cfg.fH, cfg.fW, device="cuda")
    ones_c = torch.ones(1, cfg.num_cams, cfg.C, cfg.fH, cfg.fW, device="cuda")
    for mode, seed in (("val", 1234), ("stress", 77)):
        m = synth.make_mats(cfg, 1, mode, seed=seed)
        prep = prepare_matrices(m["sensor2ego_mats"][:, 0], m["intrin_mats"][:, 0], m["ida_mats"][:, 0],
                                m["bda_mat"]).cuda()
        _, cnt = ops.lift_pool_fwd(ones_d, ones_c, prep, cid, True, False, True)
        valid, _, _ = ops.lift_indices(prep, cid, True)
        assert torch.equal(cnt & 0xF, valid.sum(1, dtype=torch.int64).reshape(1, -1)), mode


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("name", ["mini_stress", "r50_val_digest"])
def test_fused_tanh_epilogue(name, dtype):
    """voxel_output * bev_density.tanh() (BV2:627-630) folded into the BEV kernel == the product formed outside."""
    case = Case(name)
    ops, cid = _ops(case.cfg)
    vols = [t.to(dtype).cuda() for t in (case.den, case.sem, case.rgb, case.feat)]
    beta = torch.tensor(0.1, device="cuda")
    with torch.no_grad():
        raw = ops.render_fwd(*vols, beta, case.prep.cuda(), None, cid, True, 3)
        fused = ops.render_fwd(*vols, beta, case.prep.cuda(), None, cid, True, 3, None, True)
    for a, b in zip(raw[:7], fused[:7]):
        assert torch.equal(a, b)
    want = raw[7].float() * raw[6].tanh()
    assert fused[7].dtype == dtype
    assert_close_scaled(fused[7].float().cpu().numpy(), want.cpu().numpy(), 1e-6 if dtype == torch.float32 else 1e-2,
                        "fused tanh epilogue")
    v = vols[0].clone().requires_grad_(True)
    with pytest.raises(RuntimeError, match="forward-only"):
        ops.render_fwd(v, *vols[1:], beta, case.prep.cuda(), None, cid, True, 3, None, True)
