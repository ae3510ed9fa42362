"""One full-size DfMBackbone forward; with DFM_TC_ROLE_CYCLES=1 the library prints per-role
busy / wait cycles of every tensor-core conv launch (stderr)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import bench  # noqa: E402
from depth_from_motion_b200 import capi, modules  # noqa: E402
from depth_from_motion_b200 import synthetic as syn  # noqa: E402

capi.lib()
cur, prev, metas, params = syn.make_kitti_pair(100, bench.H, bench.W, bench.D,
                                               ori_shape=bench.ORI_SHAPE)
cfg = syn.depth_cfg_for(bench.D)
model = modules.DfMBackbone(in_channels=bench.C, depth_cfg=cfg).cuda().eval()
model.load_state_dict(params, strict=True)
model.downsampled_depth = bench._depths(cfg, 4)
model(cur.cuda(), prev.cuda(), metas)
torch.cuda.synchronize()
