"""Time DepthHead.forward (both volumes + depth_preds) at the benchmarked size."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import bench  # noqa: E402
from depth_from_motion_b200 import capi, modules  # noqa: E402
from depth_from_motion_b200 import synthetic as syn  # noqa: E402

capi.lib()
cfg = syn.depth_cfg_for(bench.D)
head = modules.DepthHead(
    depth_cfg=dict(mode='UD', num_bins=cfg['num_bins'], min_depth=2, max_depth=59.6),
    with_convs=False, num_views=1, depth_loss=dict(type='ce', loss_weight=1.0))
head.depth_samples = bench._depths(cfg, 1)
head.downsample_factor = 4
cost = torch.randn(1, 1, bench.D, bench.H // 4, bench.W // 4, device='cuda')
for _ in range(5):
    head(cost)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
torch.cuda.synchronize()
e0.record()
n = 30
for _ in range(n):
    head(cost)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / n
print(f'depth_head {ms:.4f} ms  {2 * 4 * bench.D * bench.H * bench.W * 4 * 4 / ms / 1e6:.0f} GB/s')
