"""Deterministic synthetic inputs for tests and bench.

There is no dataset or checkpoint on the GPU box, so the workload is: smooth
(band-limited) random 32-channel stereo features at the KITTI shape, the real
KITTI demo-sample geometry (numbers below were read once from the reference's
``demo/data/kitti/kitti_000008_infos.pkl``: ``P2`` and
``inv(prev_cam2global) @ cur_cam2global`` as ``VideoPipeline`` derives it,
mmdet3d/datasets/pipelines/loading.py:530-537), and random weights keyed by the
reference ``state_dict`` names.  Everything is generated from NumPy's legacy
MT19937 ``RandomState`` so every process regenerates identical tensors.
"""
import math

import numpy as np
import torch
import torch.nn.functional as F

# KITTI sample 000008, camera 2 projection padded to 4x4 (calib['P2'])
KITTI_P2 = np.array(
    [[721.5377, 0.0, 609.5593, 44.85728],
     [0.0, 721.5377, 172.854, 0.2163791],
     [0.0, 0.0, 1.0, 0.002745884],
     [0.0, 0.0, 0.0, 1.0]], dtype=np.float64)

# cur -> prev camera transforms for the three previous sweeps of the demo sample
KITTI_CUR2PREV = np.array([
    [[0.999302046, -0.001022011, 0.037342192, 0.048377033],
     [0.000957212, 0.99999813, 0.00173758, 0.015017807],
     [-0.037344459, -0.001701675, 0.999300674, 0.321288688],
     [0.0, 0.0, 0.0, 1.0]],
    [[0.99658589, -0.004466793, 0.082434821, 0.122045509],
     [0.004160789, 0.999984333, 0.003887874, 0.034121847],
     [-0.082450365, -0.003532619, 0.996588961, 0.662983686],
     [0.0, 0.0, 0.0, 1.0]],
    [[0.991866141, -0.009969419, 0.126890392, 0.213423159],
     [0.009329447, 0.999941439, 0.005634116, 0.059576494],
     [-0.126938255, -0.004405767, 0.991900695, 0.958234025],
     [0.0, 0.0, 0.0, 1.0]]], dtype=np.float64)


def depth_cfg_for(num_planes, downsample_factor=4):
    """Model-level depth_cfg (configs/dfm/dfm_r34_1x8_kitti-3d-3class.py:4-9)
    with num_bins chosen so that D = num_bins / downsample_factor planes."""
    return dict(mode='UD', num_bins=num_planes * downsample_factor,
                depth_min=2, depth_max=59.6,
                downsample_factor=downsample_factor)


def smooth_field(rng, c, h, w, cell=8):
    """Band-limited random feature map [1, c, h, w] fp32 (N(0,1) on a 1/cell
    lattice, bicubic-upsampled) -- bilinear taps of it are not noise-amplifying
    ('sampling-coordinate reproducibility')."""
    lh, lw = math.ceil(h / cell) + 3, math.ceil(w / cell) + 3
    low = torch.from_numpy(
        rng.standard_normal((1, c, lh, lw)).astype(np.float32))
    up = F.interpolate(low, size=(lh * cell, lw * cell), mode='bicubic',
                       align_corners=False)
    return up[:, :, cell:cell + h, cell:cell + w].contiguous()


def make_img_meta(h, w, sweep=2, flip=False, crop_offset=(0, 0), scale=1.0,
                  ori_shape=None):
    """The img_meta keys DfMBackbone.forward reads (dfm_backbone.py:150-172)."""
    if ori_shape is None:
        ori_shape = (h, w, 3)
    return dict(
        ori_cam2img=KITTI_P2.astype(np.float32).tolist(),
        cur2prevs=torch.from_numpy(
            KITTI_CUR2PREV[sweep:sweep + 1].astype(np.float32)),
        ori_shape=tuple(ori_shape),
        pad_shape=(h, w, 3),
        img_shape=(h, w, 3),
        flip=flip,
        crop_offset=list(crop_offset),
        scale_factor=[scale, scale, scale, scale])


def _kaiming(rng, shape, fan_in, gain=1.0):
    std = gain * math.sqrt(2.0 / fan_in)
    return (rng.standard_normal(shape) * std).astype(np.float32)


def make_backbone_params(rng, num_planes, in_channels=32, cv=32):
    """Random DfMBackbone parameters keyed by the reference state_dict names
    (SURVEY.md section 8a 'State').  GroupNorm affine is randomised so gamma/beta
    are exercised."""
    p = {}

    def gn(name, c):
        p[name + '.weight'] = (0.5 + rng.random_sample(c)).astype(np.float32)
        p[name + '.bias'] = (0.2 * rng.standard_normal(c)).astype(np.float32)

    def convmod(name, cin, cout):
        p[name + '.conv.weight'] = _kaiming(rng, (cout, cin, 3, 3, 3), cin * 27)
        gn(name + '.gn', cout)

    def hg(name, c):
        for sub, ci, co, seq in (('conv1', c, 2 * c, True),
                                 ('conv2', 2 * c, 2 * c, False),
                                 ('conv3', 2 * c, 2 * c, True),
                                 ('conv4', 2 * c, 2 * c, True)):
            pre = f'{name}.{sub}.0' if seq else f'{name}.{sub}'
            p[pre + '.0.weight'] = _kaiming(rng, (co, ci, 3, 3, 3), ci * 27)
            gn(pre + '.1', co)
        # ConvTranspose3d weight layout is (in, out, kd, kh, kw)
        p[f'{name}.conv5.0.weight'] = _kaiming(
            rng, (2 * c, 2 * c, 3, 3, 3), 2 * c * 27 / 8)
        gn(f'{name}.conv5.1', 2 * c)
        p[f'{name}.conv6.0.weight'] = _kaiming(
            rng, (2 * c, c, 3, 3, 3), 2 * c * 27 / 8)
        gn(f'{name}.conv6.1', c)

    for sfx, cin in (('', 2 * in_channels), ('_mono', in_channels)):
        convmod('dres0' + sfx, cin, cv)
        convmod('dres1' + sfx, cv, cv)
        tower = 'mono' if sfx else 'stereo'
        hg(f'hg_{tower}.0', cv)
        convmod(f'pred_{tower}.0.0', cv, cv)
        p[f'pred_{tower}.0.1.weight'] = _kaiming(rng, (1, cv, 3, 3, 3), cv * 27)
    p['aggregate_cost.weight'] = _kaiming(
        rng, (num_planes, 2 * num_planes, 1, 1), 2 * num_planes, gain=0.7)
    return {k: torch.from_numpy(v) for k, v in p.items()}


def make_kitti_pair(seed, h, w, num_planes, c=32, sweep=2, flip=False,
                    crop_offset=(0, 0), scale=1.0, ori_shape=None):
    """One synthetic (cur, prev) stereo-feature pair + img_metas + weights."""
    rng = np.random.RandomState(seed)
    cur = smooth_field(rng, c, h, w)
    prev = smooth_field(rng, c, h, w)
    params = make_backbone_params(rng, num_planes, c, 32)
    metas = [make_img_meta(h, w, sweep, flip, crop_offset, scale, ori_shape)]
    return cur, prev, metas, params


def make_neck_params(rng, template_state_dict):
    """Random DfMNeck / OutdoorImVoxelNeck parameters shaped like (and ordered as)
    ``template_state_dict``; BatchNorm running statistics are randomised so the
    eval-mode affine is exercised."""
    out = {}
    for k, v in template_state_dict.items():
        shape = tuple(v.shape)
        if k.endswith('num_batches_tracked'):
            out[k] = torch.zeros(shape, dtype=v.dtype)
        elif k.endswith('conv.weight'):
            fan_in = shape[1] * 27
            out[k] = torch.from_numpy(_kaiming(rng, shape, fan_in))
        elif k.endswith('aggregate_layer.weight'):
            out[k] = torch.from_numpy(_kaiming(rng, shape, shape[1], gain=0.5))
        elif k.endswith('bn.weight'):
            out[k] = torch.from_numpy(
                (0.5 + rng.random_sample(shape)).astype(np.float32))
        elif k.endswith('bn.bias') or k.endswith('running_mean'):
            out[k] = torch.from_numpy(
                (0.2 * rng.standard_normal(shape)).astype(np.float32))
        elif k.endswith('running_var'):
            out[k] = torch.from_numpy(
                (0.5 + rng.random_sample(shape)).astype(np.float32))
        else:
            raise KeyError(k)
    return out


# KITTI model-level voxel grid (configs/dfm/dfm_r34_1x8_kitti-3d-3class.py:1,10-11)
KITTI_POINT_CLOUD_RANGE = [2, -30.4, -3, 59.6, 30.4, 1]
KITTI_VOXEL_SIZE = [0.2, 0.2, 0.2]


def frustum_voxel_centres(point_cloud_range, n_voxels):
    """The three axes of ``DfM.prepare_coordinates_3d`` (detectors/dfm.py:193-211):
    linspace of voxel centres, (nx, ny, nz) cells."""
    pcr = point_cloud_range
    axes = []
    for a, n in enumerate(n_voxels):
        vs = (pcr[3 + a] - pcr[a]) / n
        axes.append(torch.linspace(pcr[a] + vs / 2., pcr[3 + a] - vs / 2., n,
                                   dtype=torch.float32))
    return axes


def frustum_coordinates(point_cloud_range, n_voxels):
    """coordinates_3d [nz, ny, nx, 3] holding (x, y, z), detectors/dfm.py:208-211."""
    xs, ys, zs = frustum_voxel_centres(point_cloud_range, n_voxels)
    zz, yy, xx = torch.meshgrid(zs, ys, xs, indexing='ij')
    return torch.stack([xx, yy, zz], dim=-1).float()


def make_frustum_case(seed, h, w, num_planes, n_voxels, num_3dconvs=1,
                      cat_img_feature=True):
    """Synthetic FrustumToVoxel inputs: ``h x w`` is the padded network input,
    the plane-sweep volume is [32, num_planes, h/4, w/4].  cam2img is KITTI P2
    rescaled to the h x w image so that part of the grid projects outside."""
    rng = np.random.RandomState(seed)
    ho, wo = h // 4, w // 4
    stereo = torch.cat([smooth_field(rng, 32, ho, wo, cell=4)
                        for _ in range(num_planes)], 0)
    stereo = stereo.permute(1, 0, 2, 3)[None].contiguous()  # [1,32,D,ho,wo]
    cost = torch.cat([smooth_field(rng, 1, ho, wo, cell=4)
                      for _ in range(num_planes)], 1)[:, None] * 2.0
    sem = smooth_field(rng, 32, ho, wo, cell=4)
    s = w / 1248.0
    P = KITTI_P2.copy()
    P[:2] *= s
    P[1, 2] = 0.45 * h
    params = {}
    cin = 64 if cat_img_feature else 32
    for i in range(num_3dconvs):
        ci = cin if i == 0 else 32
        params[f'voxel_convs.{i}.0.conv.weight'] = torch.from_numpy(
            _kaiming(rng, (32, ci, 3, 3, 3), ci * 27))
        params[f'voxel_convs.{i}.0.gn.weight'] = torch.from_numpy(
            (0.5 + rng.random_sample(32)).astype(np.float32))
        params[f'voxel_convs.{i}.0.gn.bias'] = torch.from_numpy(
            (0.2 * rng.standard_normal(32)).astype(np.float32))
    metas = [dict(cam2img=P.astype(np.float32).tolist(), pad_shape=(h, w, 3))]
    return dict(stereo=stereo, cost=cost.contiguous(), sem=sem, metas=metas,
                params=params,
                coordinates_3d=frustum_coordinates(KITTI_POINT_CLOUD_RANGE,
                                                   n_voxels),
                depth_cfg=depth_cfg_for(num_planes))


# Waymo multi-view workload (configs/dfm/multiview-dfm_r101_dcn_2x16_waymoD5-3d-3class_camsync
# [_10sweeps].py: input 832x1248 after MultiViewImageResize3D, FPN level-0 features at stride 4,
# n_voxels [220, 300, 12] over [-35, -75, -2, 75, 75, 4]; detectors/multiview_dfm.py:54-61)
WAYMO_N_VOXELS = [220, 300, 12]
WAYMO_RANGE = [-35.0, -75.0, -2.0, 75.0, 75.0, 4.0]
WAYMO_INPUT_HW = (832, 1248)
WAYMO_FEAT_HW = (208, 312)


def waymo_lidar2img(num_frames, num_views=5, ego_shift=1.0):
    """Synthetic pinhole rig: cameras yawed over the front half-circle, focal length
    2055 px * 0.65 (resize), principal point at the image centre; earlier frames are
    shifted `ego_shift` metres backwards.  [T*Nv, 4, 4] float64."""
    mats = []
    for f in range(num_frames):
        for v in range(num_views):
            yaw = (v - (num_views - 1) / 2) * 0.7
            r = np.array([[np.cos(yaw), np.sin(yaw), 0], [-np.sin(yaw), np.cos(yaw), 0],
                          [0, 0, 1]])
            # lidar (x fwd, y left, z up) -> camera (x right, y down, z fwd)
            l2c = np.array([[0, -1, 0], [0, 0, -1], [1, 0, 0]], dtype=np.float64) @ r
            ext = np.eye(4)
            ext[:3, :3] = l2c
            ext[:3, 3] = l2c @ np.array([-ego_shift * f, 0.03 * v, -1.5])
            k = np.array([[1335.75, 0, 624, 0], [0, 1335.75, 416, 0], [0, 0, 1, 0],
                          [0, 0, 0, 1]])
            mats.append(k @ ext)
    return np.array(mats)


def make_waymo_sample(seed, num_frames, num_views=5, channels=64, feat_hw=WAYMO_FEAT_HW,
                      input_hw=WAYMO_INPUT_HW, flip=False, scale=1.0, crop=(0.0, 0.0)):
    """One sample of MultiViewDfM.feature_transformation's inputs: [T*Nv, C, Hf, Wf] features
    and the img_meta keys multiview_dfm.py:139-170 reads."""
    rng = np.random.RandomState(seed)
    s = num_frames * num_views
    feats = torch.from_numpy(
        rng.standard_normal((s, channels) + tuple(feat_hw)).astype(np.float32))
    meta = dict(ori_lidar2img=waymo_lidar2img(num_frames, num_views),
                input_shape=tuple(input_hw),
                img_shape=[(input_hw[0], input_hw[1], 3)] * s,
                scale_factor=np.array([scale, scale, scale, scale], dtype=np.float32),
                img_crop_offset=list(crop), flip=flip, num_views=num_views,
                num_ref_frames=num_frames - 1)
    return feats, meta


def make_bev_params(rng, in_channels=160, c=64):
    """Random BEVHourglass parameters keyed like the reference state_dict
    (backbones/bev_hourglass.py:25-32, 53-119; GroupNorm variant of the KITTI config)."""
    p = {}

    def gn(name, ch):
        p[name + '.weight'] = (0.5 + rng.random_sample(ch)).astype(np.float32)
        p[name + '.bias'] = (0.2 * rng.standard_normal(ch)).astype(np.float32)

    p['compress_conv.conv.weight'] = _kaiming(rng, (c, in_channels, 3, 3), in_channels * 9)
    gn('compress_conv.gn', c)
    hg = 'bev_hourglass.'
    for sub, ci, co, seq in (('conv1', c, 2 * c, True), ('conv2', 2 * c, 2 * c, False),
                             ('conv3', 2 * c, 2 * c, True), ('conv4', 2 * c, 2 * c, True)):
        pre = f'{hg}{sub}.0' if seq else f'{hg}{sub}'
        p[pre + '.0.weight'] = _kaiming(rng, (co, ci, 3, 3), ci * 9)
        gn(pre + '.1', co)
    p[hg + 'conv5.0.weight'] = _kaiming(rng, (2 * c, 2 * c, 3, 3), 2 * c * 9 / 4)  # (in, out, k, k)
    gn(hg + 'conv5.1', 2 * c)
    p[hg + 'conv6.0.weight'] = _kaiming(rng, (2 * c, c, 3, 3), 2 * c * 9 / 4)
    gn(hg + 'conv6.1', c)
    return {k: torch.from_numpy(v) for k, v in p.items()}


def make_anchor_head_params(rng, c=64, num_convs=2, num_anchors=6, num_classes=3,
                            box_code_size=7):
    """Random LIGAAnchor3DHead parameters (dense_heads/liga_anchor3d_head.py:37-75): 6 anchors
    (3 classes x 2 rotations, KITTI config) -> 18 class / 42 box / 12 direction channels."""
    p = {}
    for br in ('cls_convs', 'reg_convs'):
        for i in range(num_convs):
            p[f'{br}.{i}.conv.weight'] = _kaiming(rng, (c, c, 3, 3), c * 9)
            p[f'{br}.{i}.gn.weight'] = (0.5 + rng.random_sample(c)).astype(np.float32)
            p[f'{br}.{i}.gn.bias'] = (0.2 * rng.standard_normal(c)).astype(np.float32)
    for name, co, k in (('conv_cls', num_anchors * num_classes, 3),
                        ('conv_reg', num_anchors * box_code_size, 3),
                        ('conv_dir_cls', num_anchors * 2, 1)):
        p[name + '.weight'] = _kaiming(rng, (co, c, k, k), c * k * k, gain=0.7)
        p[name + '.bias'] = (0.1 * rng.standard_normal(co)).astype(np.float32)
    return {k: torch.from_numpy(v) for k, v in p.items()}


BEV_CASE = dict(seed=91, nz=5, ny=44, nx=36)   # 3 x 5 tiles of the 16 x 8 conv tile, ragged


def make_bev_case(seed=91, nz=5, ny=44, nx=36, cv=32):
    """volume_feat [1, cv, nz, ny, nx] (what FrustumToVoxel returns) + parameters of the 2-D
    stage; the KITTI config has nz=5, ny=304, nx=288."""
    rng = np.random.RandomState(seed)
    vol = torch.from_numpy(rng.standard_normal((1, cv, nz, ny, nx)).astype(np.float32)).relu()
    return dict(volume=vol, bev=make_bev_params(rng, cv * nz), head=make_anchor_head_params(rng))
