"""CPU tests of the input-side metadata against the reference's
own VideoPipeline / RandomCrop3D code executed verbatim on the demo KITTI sample."""
import copy
import os
import pickle

import numpy as np
import pytest

from depth_from_motion_b200 import pipeline_meta as pm
from depth_from_motion_b200 import synthetic as syn

REF = '/root/reference'
needs_ref = pytest.mark.skipif(not os.path.isdir(os.path.join(REF, 'mmdet3d')),
                               reason='reference tree not mounted')


def test_cur2prevs_matches_recorded_demo_geometry():
    # the constants in synthetic.py were read from the demo sample through this very formula
    c2p = syn.KITTI_CUR2PREV
    cur = np.eye(4)
    prevs = [np.linalg.inv(m) for m in c2p]      # prev_cam2global with cur at the origin
    out = pm.cur2prevs(cur, prevs)
    assert out.shape == (3, 4, 4)
    assert np.allclose(out, c2p, atol=1e-9)


def test_select_ref_frames_modes():
    assert pm.select_ref_frames(3, 1, random=False).tolist() == [2]      # test: the last one
    assert pm.select_ref_frames(3, 2, random=False).tolist() == [1, 2]
    assert pm.select_ref_frames(3, -1).size == 0 and pm.select_ref_frames(0, 2).size == 0
    rng = np.random.RandomState(0)
    ids = pm.select_ref_frames(3, 5, random=True, rng=rng)               # with replacement
    assert len(ids) == 5 and set(ids.tolist()) <= {0, 1, 2}


def test_quaternion_matrix():
    m = pm.quaternion_matrix([np.cos(0.3), 0, 0, np.sin(0.3)])          # yaw 0.6 rad
    assert np.allclose(m[:2, :2], [[np.cos(0.6), -np.sin(0.6)], [np.sin(0.6), np.cos(0.6)]])
    assert np.allclose(m[:3, :3] @ m[:3, :3].T, np.eye(3))


@needs_ref
def test_video_meta_matches_reference_pipeline():
    from oracle.ref_loader import reference_class
    info = pickle.load(open(os.path.join(REF, 'demo/data/kitti/kitti_000008_infos.pkl'), 'rb'))[0]
    img_info = dict(filename='x.png', cam2global=info['image']['cam2global'],
                    sweeps=[dict(data_path=s['data_path'], cam2global=s['cam2global'])
                            for s in info['image']['sweeps']])

    class Compose:   # the image transforms are out of scope: identity
        def __init__(self, t):
            pass

        def __call__(self, r):
            r['img'] = 0
            return r

    VP = reference_class('mmdet3d/datasets/pipelines/loading.py', 'VideoPipeline',
                         dict(np=np, copy=copy, Compose=Compose))
    for nref, rand in ((1, False), (3, False), (2, True)):
        np.random.seed(7)
        ref = VP([], num_ref_imgs=nref, random=rand)(dict(img_info=copy.deepcopy(img_info)))
        got = pm.video_meta(img_info, nref, rand, rng=np.random.RandomState(7))
        assert np.array_equal(ref['cur2prevs'], got['cur2prevs'])
        assert [m for m in got['ref_filenames']] == \
            [img_info['sweeps'][i]['data_path'] for i in got['ref_ids']]
    # the three sweeps are what synthetic.KITTI_CUR2PREV records (nearest first)
    allp = pm.video_meta(img_info, 3, False)['cur2prevs']
    assert np.allclose(allp, syn.KITTI_CUR2PREV, atol=1e-6)


@needs_ref
def test_crop3d_meta_matches_reference():
    from oracle.ref_loader import reference_class

    class RandomCrop:   # mmdet base: only what _crop_data touches
        def __init__(self, **kw):
            self.bbox_clip_border = kw.get('bbox_clip_border', True)
            self.bbox2label, self.bbox2mask = {}, {}

    RC = reference_class('mmdet3d/datasets/pipelines/transforms_3d.py', 'RandomCrop3D',
                         dict(np=np, RandomCrop=RandomCrop))
    rc = RC(crop_size=(320, 1280), rel_offset_h=(0.3, 1.0))
    img = np.zeros((375, 1242, 3), dtype=np.uint8)
    np.random.seed(3)
    ref = rc._crop_data(dict(img=img, cam2img=syn.KITTI_P2.copy()), (320, 1280), True)
    rng = np.random.RandomState(3)
    x1, y1 = pm.random_crop_offsets(img.shape, (320, 1280), (0.3, 1.0), (0., 1.), rng)
    cam, off = pm.crop3d_meta(syn.KITTI_P2, x1, y1)
    assert off == ref['crop_offset']
    assert np.allclose(cam, ref['cam2img'], atol=1e-9)
    assert ref['img_shape'][:2] == (320, 1242)


def test_backbone_img_meta_feeds_geometry_packing():
    from depth_from_motion_b200 import modules
    c2p = pm.cur2prevs(np.eye(4), [np.linalg.inv(syn.KITTI_CUR2PREV[2])])
    meta = pm.backbone_img_meta(syn.KITTI_P2, c2p, (375, 1242, 3), (320, 1280, 3),
                                crop_offset=(0, 55))
    g = modules.geometry_from_meta(meta)
    assert abs(g.cur2prev[11] - syn.KITTI_CUR2PREV[2][2, 3]) < 1e-6
    assert (g.crop_x, g.crop_y) == (0.0, 55.0) and g.org_w == 1242
