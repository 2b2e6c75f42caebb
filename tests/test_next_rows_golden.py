"""SURVEY §8f rows 1 and 4 against vectors produced by the REAL reference (oracle/gen_golden_next.py ->
tests/golden/mini_next_<mode>.npz): the CPU part pins the oracle restatements, the GPU part pins the kernels
on a box that has no /root/reference."""
import os

import numpy as np
import pytest
import torch

from helpers import GOLDEN_DIR, assert_close_scaled
from oracle import torch_path as tp
from oracle.gen_golden_next import next_inputs
from vampire_b200.config import MINI


def _gold(mode):
    return np.load(os.path.join(GOLDEN_DIR, f"mini_next_{mode}.npz"), allow_pickle=False)


def _inputs(mode):
    mats, ctx, logits, cot = next_inputs(mode)
    gold = _gold(mode)
    chk = np.array([ctx.double().sum().item(), logits.double().sum().item(), cot.double().sum().item()])
    if not np.allclose(chk, gold["in_checksum"], rtol=1e-7, atol=0):
        pytest.skip("torch CPU RNG no longer reproduces the fixture's inputs")
    return mats, ctx, logits, cot, gold


@pytest.mark.parametrize("mode", ["val", "stress"])
def test_oracle_matches_reference_vectors(mode):
    from vampire_b200.matrices import prepare_matrices
    mats, ctx, logits, cot, gold = _inputs(mode)
    assert_close_scaled(tp.depth_softmax(logits).numpy(), gold["softmax_out"], 1e-6, "softmax (oracle)")
    prep = prepare_matrices(mats["sensor2ego_mats"][:, 0], mats["intrin_mats"][:, 0], mats["ida_mats"][:, 0],
                            mats["bda_mat"]).numpy()
    if not np.array_equal(prep.view(np.uint32), gold["prep"].view(np.uint32)):
        pytest.xfail("this host's LAPACK rounds the 4x4 inverses differently from the build container")
    conf = MINI.backbone_kwargs()
    x = ctx.clone().requires_grad_(True)
    out = tp.get_voxel_feats_2d(conf, tp.build_buffers(conf), x, mats)
    assert_close_scaled(out.detach().numpy(), gold["lift2d_out"], 1e-6, "2-D lift (oracle)")
    grad, = torch.autograd.grad((out * cot).sum(), x)
    assert_close_scaled(grad.numpy(), gold["lift2d_grad"], 1e-5, "2-D lift backward (oracle)")


@pytest.mark.gpu
@pytest.mark.parametrize("mode", ["val", "stress"])
def test_kernels_match_reference_vectors(mode):
    """Fed the fixture's prepared matrices (like every golden index test), so the host's LAPACK plays no part."""
    from vampire_b200 import ops
    mats, ctx, logits, cot, gold = _inputs(mode)
    sm = ops.depth_softmax_fwd(logits.cuda(), True)
    assert_close_scaled(sm.cpu().numpy(), gold["softmax_out"], 1e-6, "softmax")
    cid = ops.register_config(MINI, lift_2d=True)
    prep = torch.from_numpy(gold["prep"]).cuda()
    x = ctx.cuda().requires_grad_(True)
    ones = torch.ones(ctx.shape[0], ctx.shape[1], 1, MINI.fH, MINI.fW, device="cuda")
    out, _ = ops.lift_pool_fwd(ones, x, prep, cid, True, False, True)
    assert_close_scaled(out.detach().cpu().numpy(), gold["lift2d_out"], 1e-5, "2-D lift")
    grad, = torch.autograd.grad((out * cot.cuda()).sum(), x)
    assert_close_scaled(grad.cpu().numpy(), gold["lift2d_grad"], 2e-5, "2-D lift backward")
