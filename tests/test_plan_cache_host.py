"""CPU: the host logic of the plan cache (keying by matrix content, de-duplication inside a batch, LRU eviction,
batch-table reuse) with a stub builder -- the plans themselves are built by CUDA kernels and tested in test_gpu_plan.py."""
import torch

from vampire_b200 import plan as vplan


class _FakePlan:
    def __init__(self, tag):
        self.tag = tag
        self.buf = torch.zeros(4, dtype=torch.int32)

    def pointers(self):
        return [self.buf.data_ptr()] * 4

    def nbytes(self):
        return 16


def _mats(*vals):
    return torch.stack([torch.full((6, 6, 4, 4), float(v)) for v in vals])


def test_plan_cache_keys_by_content_and_evicts_lru(monkeypatch):
    built = []

    def builder(state, mats, has_bda):
        out = []
        for b in range(mats.shape[0]):
            built.append(float(mats[b, 0, 0, 0, 0]))
            out.append(_FakePlan(built[-1]))
        return out

    monkeypatch.setattr(vplan, "build_lift_plans", builder)
    pc = vplan.PlanCache(max_samples=3)
    b1 = pc.lift(None, 7, _mats(1, 2, 1), True)               # two distinct rigs in a batch of three
    assert built == [1.0, 2.0] and (pc.hits, pc.misses) == (0, 3)    # the statistics count samples, the builds rigs
    assert [p.tag for p in b1.plans] == [1.0, 2.0, 1.0] and b1.plans[0] is b1.plans[2]
    assert b1.table.shape == (3, 4) and b1.table._vb200_keepalive is b1.plans
    assert pc.lift(None, 7, _mats(1, 2, 1), True) is b1       # same batch: the table itself is reused
    assert built == [1.0, 2.0]
    pc.lift(None, 7, _mats(2), True)                          # known sample in another batch: no build
    assert built == [1.0, 2.0]
    pc.lift(None, 7, _mats(1), False)                         # has_bda is part of the key
    assert built == [1.0, 2.0, 1.0]
    pc.lift(None, 8, _mats(1), True)                          # ... and so is the config
    assert built == [1.0, 2.0, 1.0, 1.0]
    assert len(pc._lru) == 3                                  # capacity 3: the least recently used rig (1, cfg 7, bda) left
    pc.lift(None, 7, _mats(1), True)
    assert built[-1] == 1.0 and len(built) == 5
    # a clone with equal bytes is the same key; a changed element is not
    m = _mats(5)
    pc.lift(None, 7, m, True)
    n = len(built)
    pc.lift(None, 7, m.clone(), True)
    assert len(built) == n
    m2 = m.clone()
    m2[0, 3, 2, 1, 0] += 1e-3
    pc.lift(None, 7, m2, True)
    assert len(built) == n + 1
    pc.clear()
    assert pc.nbytes() == 0


def test_prepared_matrices_cache_follows_identity_and_version():
    """LiftRenderB200 prepares the 4x4 matrices once per distinct mats_dict contents: same tensors at the same version hit,
    an in-place write or a new tensor object misses, a dict without bda_mat is its own entry."""
    from vampire_b200 import synth
    from vampire_b200.config import MINI
    from vampire_b200.view_transform import LiftRenderB200
    mod = LiftRenderB200(**MINI.backbone_kwargs())
    mats = synth.make_mats(MINI, 2, "stress")
    a = mod._prep_dict_cached(mats, 0, "cpu")
    assert mod._prep_dict_cached(mats, 0, "cpu") is a and a[1] is True and a[2] is not None
    assert mod._prep_dict_cached(dict(mats), 0, "cpu") is a                 # another dict, the same tensor objects
    mats["ida_mats"][0, 0, 0, 0, 3] += 1.0                                   # in-place write bumps the version counter
    b = mod._prep_dict_cached(mats, 0, "cpu")
    assert b is not a and not torch.equal(b[0], a[0])
    clone = {k: v.clone() for k, v in mats.items()}                          # equal values, other objects: recomputed
    c = mod._prep_dict_cached(clone, 0, "cpu")
    assert c is not b and torch.equal(c[0], b[0])
    nobda = {k: v for k, v in mats.items() if k != "bda_mat"}
    d = mod._prep_dict_cached(nobda, 0, "cpu")
    assert d[1] is False and d is not b
    assert len(mod._prep_cache) <= 4
