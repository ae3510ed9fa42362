"""Shared helpers for the test-suite."""
import os

import numpy as np
import torch

from depth_from_motion_b200 import synthetic as syn

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')

# must match tests/golden/make_golden.py
KITTI_CASES = {
    'kitti_plain': (11, 32, 64, 8, False, (0, 0), 1.0, None),
    'kitti_aug': (12, 32, 64, 8, True, (10, 40), 1.03, (375, 1242, 3)),
}


def load_kitti_case(name):
    seed, h, w, d, flip, crop, scale, ori = KITTI_CASES[name]
    cur, prev, metas, params = syn.make_kitti_pair(
        seed, h, w, d, flip=flip, crop_offset=crop, scale=scale, ori_shape=ori)
    gold = dict(np.load(os.path.join(GOLDEN, name + '.npz')))
    # the stored inputs are authoritative (regeneration is only a convenience)
    cur = torch.from_numpy(gold['cur'])
    prev = torch.from_numpy(gold['prev'])
    return cur, prev, metas, params, syn.depth_cfg_for(d), gold


def rel_err(a, b):
    """max |a-b| / max |b|  -- the normalised max-norm error used for the 1e-3 bar."""
    a = torch.as_tensor(a, dtype=torch.float64).cpu()
    b = torch.as_tensor(b, dtype=torch.float64).cpu()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


# must match tests/golden/make_golden.py
FRUSTUM_CASE = dict(seed=31, h=32, w=64, num_planes=8, n_voxels=(24, 20, 8))


def load_frustum_case():
    c = syn.make_frustum_case(**FRUSTUM_CASE)
    gold = dict(np.load(os.path.join(GOLDEN, 'frustum.npz')))
    for k in ('stereo', 'cost', 'sem'):
        c[k] = torch.from_numpy(gold[k])
    return c, gold
