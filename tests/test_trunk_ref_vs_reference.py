"""CPU: oracle/trunk_ref.TrunkRef (the fp32 trunk the tcgen05 kernels and the CPU step are compared with) against the live,
unmodified reference modules (models/detector/yolo_ssod.py:105-118 + backbone/neck/head) on the same state_dict and input:
identical outputs in eval and in train mode, and this repo's Model has exactly the reference's state_dict keys / shapes.
Needs /root/reference (absent on the GPU box -> skipped there)."""
import os

import pytest
import torch

from oracle import ref_harness

pytestmark = pytest.mark.skipif(not ref_harness.reference_available(), reason="reference checkout not present")


@pytest.mark.parametrize("yaml_rel,depth,nd", [("configs/ssod/coco-standard/yolov5l_coco_ssod_10_percent.yaml", (3, 6, 9, 3), 3)])
def test_trunk_ref_equals_live_reference(yaml_rel, depth, nd):
    from oracle.trunk_ref import TrunkRef
    ns = ref_harness.load_reference()
    cfg = ref_harness.make_cfg(yaml_rel)
    torch.manual_seed(0)
    ref = ns.SSODModel(cfg)
    sd = {k: v.clone() for k, v in ref.state_dict().items()}
    x = torch.rand(1, 3, 128, 128, generator=torch.Generator().manual_seed(3))
    # eval (teacher pass): reference returns ((pred, raw_list), features)
    ref.eval()
    with torch.no_grad():
        (pred, raw_ref), feat_ref = ref(x)
        raw, feat = TrunkRef(sd, depth, nd).forward(x, train=False)
    for a, b in zip(raw, raw_ref):
        assert a.shape == b.shape and float((a - b).abs().max()) == 0.0
    for a, b in zip(feat, feat_ref):
        assert float((a - b).abs().max()) == 0.0
    # train (student pass): batch statistics + running-stat update with the reference's momentum 0.03
    ref.train()
    x2 = torch.rand(2, 3, 128, 128, generator=torch.Generator().manual_seed(4))
    sd_t = {k: v.clone() for k, v in sd.items()}
    out_ref, feat_ref = ref(x2)
    raw, feat = TrunkRef(sd_t, depth, nd, bn_momentum=0.03).forward(x2, train=True)
    for a, b in zip(raw, out_ref):
        assert float((a - b).abs().max()) == 0.0
    for a, b in zip(feat, feat_ref):
        assert float((a - b).abs().max()) == 0.0
    after = ref.state_dict()
    for k in sd_t:
        if "running_" in k:
            assert torch.equal(sd_t[k], after[k]), k


def test_model_state_dict_keys_equal_reference():
    from efficientteacher_b200.config import yolov5_ssod_cfg, yolov5_sup_cfg
    from efficientteacher_b200.model import Model, SupModel
    ns = ref_harness.load_reference()
    ref = ns.SSODModel(ref_harness.make_cfg("configs/ssod/coco-standard/yolov5l_coco_ssod_10_percent.yaml"))
    mine = Model(yolov5_ssod_cfg('l'))
    a, b = ref.state_dict(), mine.state_dict()
    assert list(a.keys()) == list(b.keys())
    assert all(a[k].shape == b[k].shape and a[k].dtype == b[k].dtype for k in a)
    ref_s = ns.SupModel(ref_harness.make_cfg("configs/sup/public/yolov5s_coco.yaml"))
    mine_s = SupModel(yolov5_sup_cfg('s'))
    a, b = ref_s.state_dict(), mine_s.state_dict()
    assert sorted(a.keys()) == sorted(b.keys()) and all(a[k].shape == b[k].shape for k in a)
