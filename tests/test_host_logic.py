"""CPU: host-side mirror of the reference interface -- registry names, constructor arguments, config restatements,
argument errors of the training-side entry points (no compute: there is no GPU in the build container)."""
import ctypes as C

import pytest
import torch


def test_registry_has_the_reference_names():
    import unidet3d_b200 as u
    for name in ("SpConvUNet", "UniDet3DEncoder", "UniDet3D", "UniDet3DCriterion"):
        assert name in u.MODELS, name          # unidet3d/spconv_unet.py:94, encoder.py:113, unidet3d.py:20, criterion.py:7


def test_criterion_takes_the_reference_config():
    """configs/unidet3d_1xb8_scannet.py:60-89 and the joint config's criterion dict build unchanged."""
    import unidet3d_b200 as u
    from unidet3d_b200 import configs
    c = u.MODELS.build(configs.criterion_cfg(("scannet",)))
    assert (c.w_cls, c.w_box, c.loss_weight, c.non_object_weight, c.topk, c.iter_matcher) == (0.5, 2.0, [0.5, 1.0], 0.1, [6], True)
    j = u.MODELS.build(configs.criterion_cfg(configs.JOINT))
    assert j.topk == [6, 6, 3, 3, 3, 3] and j.datasets == list(configs.JOINT) and j.datasets_weights == [1.0] * 6
    with pytest.raises(NotImplementedError):
        bad = configs.criterion_cfg(("scannet",))
        bad["bbox_loss_simple"] = dict(type="UniDet3DAxisAlignedIoULoss", mode="iou", reduction="none")
        u.MODELS.build(bad)


def test_detector_builds_with_criterion_and_has_no_cpu_path():
    import unidet3d_b200 as u
    from unidet3d_b200 import configs, _lib
    from unidet3d_b200.structures import Det3DDataSample, InstanceData, PointData
    cfg = configs.model_cfg(("scannet",), num_planes=[32, 64], num_layers=1, d_model=128, num_heads=4, hidden_dim=128)
    model = u.MODELS.build(cfg).eval()
    assert isinstance(model.criterion, u.UniDet3DCriterion) and model.train_cfg["topk"] == 6
    pts = torch.rand(500, 6)
    sample = Det3DDataSample(lidar_path="data/scannet/points/x.bin",
                             gt_pts_seg=PointData(sp_pts_mask=torch.randint(0, 20, (500,)), pts_instance_mask=torch.randint(-1, 3, (500,))),
                             gt_instances_3d=InstanceData(labels_3d=torch.tensor([1, 2, 3]), sp_masks=torch.zeros(3, 20, dtype=torch.bool)))
    with pytest.raises((_lib.Ud3dError, RuntimeError)):     # CPU tensors: the ops refuse, nothing falls back
        model.loss(dict(points=[pts]), [sample])
    with pytest.raises(RuntimeError):
        model.forward_scenes([pts.numpy()], [torch.zeros(500, dtype=torch.int64).numpy()], ["scannet"])


def test_criterion_entry_point_argument_errors():
    from unidet3d_b200 import _lib
    lib = _lib.load()
    assert lib.ud3d_criterion_layer(None, None, 0, None) == -1 and b"NULL" in lib.ud3d_last_error()
    a = _lib.CriterionArgs()
    buf = (C.c_float * 64)()
    a.logits = C.addressof(buf); a.ld_logits = 4; a.T = 3; a.C1 = 4
    a.boxes = C.addressof(buf); a.box_dim = 5
    a.sums = C.addressof(buf)
    assert lib.ud3d_criterion_layer(C.byref(a), None, 0, None) == -1 and b"box_dim" in lib.ud3d_last_error()
    a.box_dim = 6; a.topk = 99
    assert lib.ud3d_criterion_layer(C.byref(a), None, 0, None) == -1 and b"topk" in lib.ud3d_last_error()
    assert lib.ud3d_criterion_workspace_bytes(3000, 40) >= 2 * 3000 * 4 + 40 * 4
    assert lib.ud3d_subm3_tile_order_workspace_bytes(1000) >= 65536 * 4 + 2000
    assert lib.ud3d_subm3_tile_order(None, 5, None, None, None, None, 0, None) == -1


def test_training_entry_points_have_no_cpu_path_and_check_their_arguments():
    """The training step's host logic: loss_backward needs train mode, the criterion gradients need iter_matcher=True
    (every reference config), and the new ops refuse CPU tensors (no fallback)."""
    import types
    import unidet3d_b200 as u
    from unidet3d_b200 import _lib, configs, ops, train
    model = u.MODELS.build(configs.model_cfg(("scannet",))).eval()
    with pytest.raises(RuntimeError, match="train"):
        train.loss_backward(model, dict(points=[]), [])
    crit = u.MODELS.build(configs.criterion_cfg(("scannet",)))
    crit.iter_matcher = False
    with pytest.raises(NotImplementedError):
        train.criterion_backward(crit, dict(cls_preds=[], bboxes=[], aux_outputs=[]), [], [])
    T, G = 8, 2
    logits, boxes = torch.zeros(T, 19), torch.zeros(T, 6)
    with pytest.raises(_lib.Ud3dError):
        ops.criterion_layer_grad(logits, boxes, torch.zeros(G, 6), torch.zeros(G, dtype=torch.long), torch.zeros(T, G, dtype=torch.bool),
                                 torch.zeros(4), torch.zeros(2), 0.1)
    with pytest.raises(_lib.Ud3dError):
        ops.head_backward(torch.zeros(T, 8), None, False, None, torch.zeros(19, dtype=torch.int32), torch.zeros(T, 8), torch.zeros(T, 19))
    with pytest.raises(_lib.Ud3dError):
        ops.attention_backward(torch.zeros(T, 768), torch.tensor([0, T], dtype=torch.int32), 8, torch.zeros(T, 256), torch.zeros(T, 256))
    # the dataset-weight vector of a batch is cached per batch composition
    w1 = crit.scene_weights(torch.zeros(2, 4), ["scannet", "scannet"])
    assert w1 is crit.scene_weights(torch.zeros(2, 4), ["scannet", "scannet"]) and w1.tolist() == [1.0, 1.0]


def test_synthetic_scannet_annotations_are_consistent():
    """unidet3d_b200.synthetic.make_scannet_gt (bench.py --workload train_b8): instances are unions of superpoints, every
    instance owns at least one, the per-point instance ids agree with the superpoint masks."""
    import numpy as np
    from unidet3d_b200.synthetic import make_scene, make_scannet_gt, SCENE_PRESETS
    n, v, a, c = SCENE_PRESETS["tiny"]
    pts, sp = make_scene(3, n, a, c)
    labels, sp_masks, inst = make_scannet_gt(sp, 6, 11)
    assert labels.shape == (6,) and sp_masks.shape == (6, int(sp.max()) + 1) and inst.shape == sp.shape
    assert sp_masks.any(1).all() and sp_masks.sum(0).max() <= 1 and inst.min() >= -1 and inst.max() == 5
    for k in range(6):
        assert np.array_equal(inst == k, sp_masks[k][sp])
    again = make_scannet_gt(sp, 6, 11)
    assert all(np.array_equal(x, y) for x, y in zip((labels, sp_masks, inst), again))
