"""Model configurations of the reference (configs/unidet3d_1xb8_scannet.py:28-96 and
configs/unidet3d_1xb8_scannet_s3dis_multiscan_3rscan_scannetpp_arkitscenes.py:28-96), restated as
plain dicts so tests and bench can build the models without mmengine's config loader."""
from __future__ import annotations

import copy

CLASSES = dict(
    scannet=['cabinet', 'bed', 'chair', 'sofa', 'table', 'door', 'window', 'bookshelf', 'picture', 'counter', 'desk',
             'curtain', 'refrigerator', 'showercurtrain', 'toilet', 'sink', 'bathtub', 'otherfurniture'],
    s3dis=['table', 'chair', 'sofa', 'bookcase', 'board'],
    multiscan=['door', 'table', 'chair', 'cabinet', 'window', 'sofa', 'microwave', 'pillow', 'tv_monitor', 'curtain',
               'trash_can', 'suitcase', 'sink', 'backpack', 'bed', 'refrigerator', 'toilet'],
    scannetpp=['table', 'door', 'ceiling lamp', 'cabinet', 'blinds', 'curtain', 'chair', 'storage cabinet',
               'office chair', 'bookshelf', 'whiteboard', 'window', 'box', 'monitor', 'shelf', 'heater',
               'kitchen cabinet', 'sofa', 'bed', 'trash can', 'book', 'plant', 'blanket', 'tv', 'computer tower',
               'refrigerator', 'jacket', 'sink', 'bag', 'picture', 'pillow', 'towel', 'suitcase', 'backpack', 'crate',
               'keyboard', 'rack', 'toilet', 'printer', 'poster', 'painting', 'microwave', 'shoes', 'socket', 'bottle',
               'bucket', 'cushion', 'basket', 'shoe rack', 'telephone', 'file folder', 'laptop', 'plant pot',
               'exhaust fan', 'cup', 'coat hanger', 'light switch', 'speaker', 'table lamp', 'kettle',
               'smoke detector', 'container', 'power strip', 'slippers', 'paper bag', 'mouse', 'cutting board',
               'toilet paper', 'paper towel', 'pot', 'clock', 'pan', 'tap', 'jar', 'soap dispenser', 'binder', 'bowl',
               'tissue box', 'whiteboard eraser', 'toilet brush', 'spray bottle', 'headphones', 'stapler', 'marker'],
    arkitscenes=['cabinet', 'refrigerator', 'shelf', 'stove', 'bed', 'sink', 'washer', 'toilet', 'bathtub', 'oven',
                 'dishwasher', 'fireplace', 'stool', 'chair', 'table', 'tv_monitor', 'sofa'],
)
CLASSES['3rscan'] = CLASSES['scannet']


def model_cfg(datasets=('scannet',), num_channels=32, voxel_size=0.02, num_planes=None, num_layers=6, d_model=256,
              num_heads=8, hidden_dim=1024, topk_insts=1000):
    """``datasets=('scannet',)`` -> unidet3d_1xb8_scannet.py; the 6-name tuple -> the joint config."""
    per = dict(scannet=(True, False, True, True, False, 0.5), s3dis=(True, False, True, False, False, 0.55),
               multiscan=(False, True, True, True, False, 0.55), **{'3rscan': (False, True, False, True, False, 0.55)},
               scannetpp=(False, True, False, True, False, 0.55), arkitscenes=(False, True, False, None, True, 0.55))
    ds = list(datasets)
    planes = num_planes or [num_channels * (i + 1) for i in range(5)]
    return dict(
        type='UniDet3D', data_preprocessor=dict(type='Det3DDataPreprocessor_'), in_channels=6,
        num_channels=planes[0], voxel_size=voxel_size, min_spatial_shape=128, query_thr=3000,
        bbox_by_mask=[per[d][0] for d in ds], target_by_distance=[per[d][1] for d in ds],
        use_superpoints=[per[d][2] for d in ds], fast_nms=[per[d][3] for d in ds],
        backbone=dict(type='SpConvUNet', num_planes=planes, return_blocks=True),
        decoder=dict(type='UniDet3DEncoder', num_layers=num_layers, datasets_classes=[CLASSES[d] for d in ds],
                     in_channels=planes[0], d_model=d_model, num_heads=num_heads, hidden_dim=hidden_dim, dropout=0.0,
                     activation_fn='gelu', datasets=ds, angles=[per[d][4] for d in ds]),
        criterion=criterion_cfg(ds), train_cfg=dict(topk=6),
        test_cfg=dict(low_sp_thr=0.18, up_sp_thr=0.81, topk_insts=topk_insts, score_thr=0,
                      iou_thr=[per[d][5] for d in ds]))


def criterion_cfg(datasets):
    """configs/unidet3d_1xb8_scannet.py:60-89 / the joint config's :60-92."""
    diou = lambda t: dict(type=t, mode='diou', reduction='none')
    simple, rotated = diou('UniDet3DAxisAlignedIoULoss'), diou('UniDet3DRotatedIoU3DLoss')
    topk = dict(scannet=6, s3dis=6, multiscan=3, scannetpp=3, arkitscenes=3, **{'3rscan': 3})
    return dict(type='UniDet3DCriterion', datasets=list(datasets), datasets_weights=[1] * len(datasets),
                bbox_loss_simple=simple, bbox_loss_rotated=rotated,
                matcher=dict(type='UniMatcher', costs=[dict(type='QueryClassificationCost', weight=0.5),
                                                       dict(type='BboxCostJointTraining', weight=2.0, loss_simple=simple,
                                                            loss_rotated=rotated)]),
                loss_weight=[0.5, 1.0], non_object_weight=0.1, topk=[topk[d] for d in datasets], iter_matcher=True)


JOINT = ('scannet', 's3dis', 'multiscan', '3rscan', 'scannetpp', 'arkitscenes')


def oracle_cfg(cfg):
    """The dict oracle.detector.forward_scenes expects, derived from a model cfg."""
    d = cfg['decoder']
    return dict(voxel_size=cfg['voxel_size'], min_spatial_shape=cfg['min_spatial_shape'],
                encoder=dict(num_layers=d['num_layers'], num_heads=d['num_heads'], activation_fn=d['activation_fn'],
                             datasets=d['datasets'], datasets_classes=d['datasets_classes'], angles=d['angles']),
                test_cfg=copy.deepcopy(cfg['test_cfg']), fast_nms=cfg['fast_nms'], use_superpoints=cfg['use_superpoints'])
