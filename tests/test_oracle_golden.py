"""CPU tests: the oracle restatement against (1) the reference's own golden vectors
for this path and (2) fixtures produced by the unmodified
reference sources (tests/golden/make_golden.py)."""
import os

import numpy as np
import pytest
import torch

from depth_from_motion_b200 import synthetic as syn
from oracle import dfm_oracle as O
from tests.util import GOLDEN, KITTI_CASES, load_kitti_case


def test_points_img2cam_kat():
    # reference tests/test_utils/test_utils.py:186-193
    points = torch.tensor([[0.5764, 0.9109, 0.7576], [0.6656, 0.5498, 0.9813]])
    cam2img = torch.tensor([[700., 0., 450., 0.], [0., 700., 200., 0.],
                            [0., 0., 1., 0.]])
    expected = torch.tensor([[-0.4864, -0.2155, 0.7576],
                             [-0.6299, -0.2796, 0.9813]])
    assert torch.allclose(O.points_img2cam(points, cam2img), expected, atol=1e-3)


def test_points_cam2img_kat():
    # reference tests/test_utils/test_box3d.py:1653-1680
    torch.manual_seed(0)
    points = torch.rand([5, 3])
    proj_mat = torch.rand([4, 4])
    expected = torch.tensor([[0.5832, 0.6496], [0.6146, 0.7910],
                             [0.6994, 0.7782], [0.5623, 0.6303],
                             [0.4359, 0.6532]])
    assert torch.allclose(O.points_cam2img(points, proj_mat), expected, 1e-3)
    expected_d = torch.tensor([[0.5832, 0.6496, 1.7577], [0.6146, 0.7910, 1.5477],
                               [0.6994, 0.7782, 2.0091], [0.5623, 0.6303, 1.8739],
                               [0.4359, 0.6532, 1.2056]])
    assert torch.allclose(O.points_cam2img(points, proj_mat, with_depth=True),
                          expected_d, 1e-3)


def test_point_sample_kat():
    # reference tests/test_models/test_fusion/test_point_fusion.py:13-41 (the
    # no-3D-augmentation half; PointFusion.sample_single -> point_sample with
    # scale 1, crop 0, no flip, bilinear, align_corners=True)
    lidar2img = torch.tensor(
        [[6.0294e+02, -7.0791e+02, -1.2275e+01, -1.7094e+02],
         [1.7678e+02, 8.8088e+00, -7.0794e+02, -1.0257e+02],
         [9.9998e-01, -1.5283e-03, -5.2907e-03, -3.2757e-01],
         [0.0000e+00, 0.0000e+00, 0.0000e+00, 1.0000e+00]])
    img_feat = torch.arange(370 * 1224)[None, ...].view(
        370, 1224)[None, None, ...].float() / (370 * 1224)
    pts = torch.tensor([[8.356, -4.312, -0.445], [11.777, -6.724, -0.564],
                        [6.453, 2.53, -1.612], [6.227, -3.839, -0.563]])
    out = O.point_sample(img_feat, pts, lidar2img, pts.new_tensor([1., 1.]),
                         0, False, (370, 1224), (370, 1224), aligned=True)
    expected = torch.tensor([0.5560822, 0.5476625, 0.9687978, 0.6241757])
    assert torch.allclose(expected, out.squeeze(), 1e-4)


@pytest.mark.parametrize('name', sorted(KITTI_CASES))
def test_backbone_matches_reference_fixture(name):
    cur, prev, metas, params, cfg, gold = load_kitti_case(name)
    spec = KITTI_CASES[name]
    with torch.no_grad():
        vol = O.build_dfm_cost(
            cur, prev, O.downsampled_depth(cfg), 1, 4,
            torch.as_tensor(np.array([metas[0]['ori_cam2img']]),
                            dtype=torch.float32), metas[0]['cur2prevs'],
            metas[0]['ori_shape'][:2], spec[4], metas[0]['crop_offset'],
            img_scale_factor=spec[6])
        cost, stereo, mono = O.dfm_backbone_forward(params, cur, prev, metas, cfg)
        _, sm, preds = O.depth_head_forward(cost, O.depth_samples(cfg))
    # same ATen ops, same order, same machine class: tiny tolerance only for
    # thread-count dependent summation order
    for got, key in ((vol, 'volume'), (cost, 'cost'), (stereo, 'stereo'),
                     (mono, 'mono'), (preds, 'depth_preds')):
        ref = torch.from_numpy(gold[key])
        assert got.shape == ref.shape, key
        assert torch.allclose(got, ref, rtol=1e-4, atol=1e-5), key
    assert torch.allclose(sm[0, 0, :, ::8, ::8],
                          torch.from_numpy(gold['softmax_slice']), atol=1e-6)


def test_generated_inputs_match_stored_inputs():
    # fixture inputs regenerate bit-identically from their seed
    for name, spec in KITTI_CASES.items():
        seed, h, w, d = spec[:4]
        cur, prev, _, _ = syn.make_kitti_pair(seed, h, w, d)
        gold = np.load(os.path.join(GOLDEN, name + '.npz'))
        assert np.array_equal(cur.numpy(), gold['cur'])
        assert np.array_equal(prev.numpy(), gold['prev'])


@pytest.mark.parametrize('name', ['neck_dfm', 'neck_imvoxel'])
def test_neck_matches_reference_fixture(name):
    from depth_from_motion_b200 import modules
    gold = np.load(os.path.join(GOLDEN, name + '.npz'))
    rng = np.random.RandomState(21)
    dfm = modules.DfMNeck(64, 256, num_frames=2)
    imv = modules.OutdoorImVoxelNeck(64, 256)
    # make_golden draws the DfMNeck parameters first, then its input, then the
    # OutdoorImVoxelNeck parameters: replay the same stream
    sd_dfm = syn.make_neck_params(rng, dfm.state_dict())
    x_dfm = rng.standard_normal((1, 128, 6, 5, 12)).astype(np.float32)
    sd_imv = syn.make_neck_params(rng, imv.state_dict())
    x = torch.from_numpy(gold['x'])
    if name == 'neck_dfm':
        assert np.array_equal(x_dfm, gold['x'])
        with torch.no_grad():
            y = O.dfm_neck_forward(sd_dfm, x, 64)[0]
    else:
        with torch.no_grad():
            y = O.imvoxel_neck_forward(sd_imv, x)[0]
    assert torch.allclose(y, torch.from_numpy(gold['y']), rtol=1e-4, atol=1e-4)


@pytest.mark.parametrize('name', ['neck_dfm_mt', 'neck_imvoxel_mt'])
def test_neck_multitile_fixture(name):
    """The multi-tile neck fixtures (37 x 21 x 12 voxels = 3 x 3 tiles of the CUDA kernel,
    ragged edges): inputs regenerate from the seed (checksums stored), the oracle restatement
    reproduces the verbatim reference output."""
    from depth_from_motion_b200 import modules
    from tests.util import make_neck_mt_case
    gold = np.load(os.path.join(GOLDEN, name + '.npz'))
    rng, x = make_neck_mt_case(name)
    assert float(x.double().sum()) == float(gold['x_sum'])
    assert float(x.double().abs().sum()) == float(gold['x_abs'])
    mod = (modules.DfMNeck(64, 256, num_frames=2) if name == 'neck_dfm_mt'
           else modules.OutdoorImVoxelNeck(64, 256))
    sd = syn.make_neck_params(rng, mod.state_dict())
    with torch.no_grad():
        y = (O.dfm_neck_forward(sd, x, 64) if name == 'neck_dfm_mt'
             else O.imvoxel_neck_forward(sd, x))[0]
    ref = torch.from_numpy(gold['y'])
    assert y.shape == ref.shape == (1, 256, 21, 37)
    assert torch.allclose(y, ref, rtol=1e-4, atol=1e-4)


def test_depth_tables():
    cfg = syn.depth_cfg_for(72)
    d = O.downsampled_depth(cfg)
    assert d.shape == (72,) and abs(float(d[0]) - (2 + 0.5 * 4 * 0.2)) < 1e-5
    s = O.depth_samples(cfg)
    assert s.shape == (288,) and abs(float(s[-1]) - (59.6 - 0.1)) < 1e-4


def test_frustum_to_voxel_matches_reference_fixture():
    """oracle restatement (DepthHead -> FrustumToVoxel)
    against the verbatim reference run stored in tests/golden/frustum.npz."""
    from tests.util import load_frustum_case
    c, gold = load_frustum_case()
    cfg = c['depth_cfg']
    _, sm, _ = O.depth_head_forward(c['cost'], O.depth_samples(cfg), 4)
    assert torch.equal(sm, torch.from_numpy(gold['softmax']))
    out = O.frustum_to_voxel_forward(c['params'], c['stereo'], sm, c['metas'], c['sem'],
                                     c['coordinates_3d'], cfg)
    ref = torch.from_numpy(gold['out'])
    assert out.shape == ref.shape == (1, 32, 2, 20, 24)
    assert float((out - ref).abs().max()) <= 1e-6
    # the grid of the shipped config (detectors/dfm.py:174-211)
    full = O.frustum_coordinates_3d(dict(point_cloud_range=syn.KITTI_POINT_CLOUD_RANGE,
                                         voxel_size=syn.KITTI_VOXEL_SIZE))
    assert tuple(full.shape) == (20, 304, 288, 3)
    assert torch.equal(full, syn.frustum_coordinates(syn.KITTI_POINT_CLOUD_RANGE,
                                                     (288, 304, 20)))


@pytest.mark.skipif(not os.path.isdir('/root/reference/mmdet3d'), reason='reference tree not mounted')
@pytest.mark.parametrize('flip,aligned', [(False, True), (True, True), (False, False)])
def test_voxel_sample_matches_reference_source(flip, aligned):
    """SURVEY.md row a8 (oracle only): the restatement against the reference function executed
    verbatim (point_fusion.py:324-410)."""
    import torch.nn.functional as F
    from oracle.ref_loader import load_reference, reference_function
    ns = load_reference()
    ref = reference_function('mmdet3d/models/fusion_layers/point_fusion.py', 'voxel_sample',
                             dict(torch=torch, F=F, points_img2cam=ns.points_img2cam))
    g = torch.Generator().manual_seed(3)
    vox = torch.randn(1, 6, 20, 16, 8, generator=g)
    vrange, vsize = [0.0, -8.0, -2.0, 20.0, 8.0, 2.0], [1.0, 1.0, 0.5]
    depths = torch.linspace(2.0, 18.0, 16)
    # lidar -> image: camera looks along +x
    k = torch.tensor([[40., 0, 32, 0], [0, 40., 16, 0], [0, 0, 1, 0], [0, 0, 0, 1]])
    l2c = torch.tensor([[0., -1, 0, 0], [0, 0, -1, 0.3], [1, 0, 0, 0.1], [0, 0, 0, 1]])
    proj = k @ l2c
    args = (vox, vrange, vsize, depths, proj, 4, torch.tensor([1.02, 0.98]),
            torch.tensor([1.0, 2.0]), flip, (32, 64), (30, 62))
    a = ref(*args, aligned=aligned)
    b = O.voxel_sample(*args, aligned=aligned)
    assert a.shape == b.shape == (1, 6, 4, 8, 16)
    assert torch.equal(a, b)
    assert float(a.abs().sum()) > 0


def test_bev_stage_matches_reference_fixture():
    """oracle restatement of BEVHourglass + LIGAAnchor3DHead
    (class scores, 3-D box regressions, direction logits) against the verbatim reference run."""
    gold = np.load(os.path.join(GOLDEN, 'bev_stage.npz'))
    c = syn.make_bev_case(**syn.BEV_CASE)
    assert float(c['volume'].double().sum()) == float(gold['x_sum'])
    v = c['volume']
    with torch.no_grad():
        prehg, bev = O.bev_hourglass_forward(c['bev'], v.reshape(1, -1, v.shape[3], v.shape[4]))
        outs = O.dfm_bev_stage(c['bev'], c['head'], v)
    for got, key in ((prehg, 'prehg'), (bev, 'bev'), (outs[0], 'cls_score'),
                     (outs[1], 'bbox_pred'), (outs[2], 'dir_cls_preds')):
        ref = torch.from_numpy(gold[key])
        assert got.shape == ref.shape, key
        assert torch.allclose(got, ref, rtol=1e-4, atol=1e-5), key
    assert outs[0].shape[1] == 18 and outs[1].shape[1] == 42 and outs[2].shape[1] == 12


@pytest.mark.skipif(not os.path.isdir('/root/reference/mmdet3d'), reason='reference tree not mounted')
def test_bev_stage_oracle_equals_reference_source():
    """Bit-for-bit against the reference classes executed in place (only where mounted)."""
    from oracle.ref_loader import load_reference
    ns = load_reference()
    c = syn.make_bev_case(seed=5, nz=5, ny=12, nx=16)
    gn = dict(type='GN', num_groups=32, requires_grad=True)
    bev = ns.BEVHourglass(160, 64, norm_cfg=gn).eval()
    head = ns.LIGAAnchor3DHead(3, 64, 64, 6, norm_cfg=gn).eval()
    bev.load_state_dict(c['bev'], strict=True)
    head.load_state_dict(c['head'], strict=True)
    x = c['volume'].reshape(1, 160, 12, 16)
    with torch.no_grad():
        _, feat = bev(x)
        ref = head.forward_single(feat)
        got = O.dfm_bev_stage(c['bev'], c['head'], c['volume'])
    for a, b in zip(got, ref):
        assert torch.equal(a, b)


@pytest.mark.skipif(not os.path.isdir('/root/reference/mmdet3d'), reason='reference tree not mounted')
def test_spp_unet_lastconv_equals_reference_module():
    """the oracle restatement of SPPUNetNeck.lastconv against the
    reference class executed in place (spp_unet_neck.py:60-75, :110)."""
    from oracle.ref_loader import load_reference
    ns = load_reference()
    gn = dict(type='GN', num_groups=32, requires_grad=True)
    m = ns.SPPUNetNeck(in_channels=[3, 64, 128, 128, 128], start_level=2, sem_channels=[128, 32],
                       stereo_channels=[32, 32], with_upconv=True, cat_img_feature=True,
                       norm_cfg=gn).eval()
    p = {k: v for k, v in m.state_dict().items() if k.startswith('lastconv')}
    assert sorted(p) == ['lastconv.0.conv.weight', 'lastconv.0.gn.bias', 'lastconv.0.gn.weight',
                         'lastconv.1.weight']
    x = torch.randn(1, 32, 24, 40, generator=torch.Generator().manual_seed(4))
    with torch.no_grad():
        assert torch.equal(m.lastconv(x), O.spp_unet_lastconv(p, x))
    # the mirror takes the same keys
    from depth_from_motion_b200 import modules
    modules.SPPUNetNeckTail().load_state_dict(p, strict=True)
