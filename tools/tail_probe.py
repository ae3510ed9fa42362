"""One launch of every HBM-bound kernel of the path (for `ncu -k regex:...` captures):
DfMBackbone + DepthHead at the benchmarked KITTI size, one host-pipeline frame
(FrustumToVoxel gather / pool) and one Waymo 2 x 5-view lifting."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import bench  # noqa: E402
from depth_from_motion_b200 import capi, modules  # noqa: E402
from depth_from_motion_b200 import synthetic as syn  # noqa: E402

capi.lib()
H, W, D, C = bench.H, bench.W, bench.D, bench.C
cur, prev, metas, params = syn.make_kitti_pair(100, H, W, D, ori_shape=bench.ORI_SHAPE)
metas[0]['cam2img'] = syn.KITTI_P2.astype('float32').tolist()
cfg = syn.depth_cfg_for(D)
model = modules.DfMBackbone(in_channels=C, depth_cfg=cfg).cuda().eval()
model.load_state_dict(params, strict=True)
model.downsampled_depth = bench._depths(cfg, 4)
head = modules.DepthHead(
    depth_cfg=dict(mode='UD', num_bins=cfg['num_bins'], min_depth=2, max_depth=59.6),
    with_convs=False, num_views=1, depth_loss=dict(type='ce', loss_weight=1.0))
head.depth_samples = bench._depths(cfg, 1)
head.downsample_factor = 4
cost, stereo, mono = model(cur.cuda(), prev.cuda(), metas)
head(cost)
fc = syn.make_frustum_case(7, H, W, D, (288, 304, 20))
frustum = modules.FrustumToVoxel().eval()
frustum.load_state_dict(fc['params'], strict=True)
frustum = frustum.cuda()
frustum.coordinates_3d = fc['coordinates_3d']
frustum.depth_cfg = cfg
pipe = modules.HotPathPipeline(model, head, frustum)
pipe(cur.pin_memory(), prev.pin_memory(), fc['sem'].contiguous().pin_memory(), metas)
feats, meta = syn.make_waymo_sample(200, 2, 5)
modules.multiview_lift(feats.cuda(), meta, list(syn.WAYMO_N_VOXELS), list(syn.WAYMO_RANGE), 5, 2,
                       'concat')
torch.cuda.synchronize()
print('tail_probe done')
