/*
 * dfm_b200.h -- C ABI of the B200-native DfM plane-sweep cost-volume path.
 *
 * The reference (Tai-Wang/Depth-from-Motion @ e2321189) has no FFI: its hot path
 * is a chain of PyTorch calls inside mmcv-registry modules.  This header is the
 * boundary a maintainer binds instead of those chains; every entry point names the
 * reference interface it replaces (file:line relative to the reference checkout).
 * INTEGRATION.md shows the ctypes binding used by depth_from_motion_b200/modules.py
 * and the ten-line patch that makes the reference's own modules call it.
 *
 * Conventions
 *   - plain C types only; all tensors are dense fp32, layouts stated per argument;
 *   - pointers named d_* are CUDA device pointers, h_* are host pointers;
 *   - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream);
 *     device entry points are asynchronous on that stream, *_host entry points
 *     synchronise the stream before returning;
 *   - every function returns DFM_OK (0) or a DFM_ERR_* code; dfm_last_error()
 *     returns a thread-local human-readable message for the last failure;
 *   - there is no CPU fallback anywhere behind this ABI.
 */
#ifndef DFM_B200_H_
#define DFM_B200_H_

#ifdef __cplusplus
extern "C" {
#endif

#define DFM_OK 0
#define DFM_ERR_INVALID 1 /* bad argument / unsupported shape           */
#define DFM_ERR_CUDA 2    /* CUDA runtime error (message has the cause)  */
#define DFM_ERR_STATE 3   /* missing parameter, depths not set, ...      */
#define DFM_ERR_NOGPU 4   /* no sm_100 device visible                    */

/* conv implementation selector (dfm_backbone_desc_t.conv_impl, dfm_op_conv3d) */
#define DFM_CONV_AUTO 0 /* tcgen05 tensor-core kernels where implemented, SIMT else */
#define DFM_CONV_SIMT 1 /* fp32 CUDA-core kernels everywhere (bring-up / cross-check) */
#define DFM_CONV_TC 2   /* tcgen05 only; error if a layer has no tensor-core kernel   */
#define DFM_CONV_TC_NECK 3 /* dfm_op_conv3d only: force the K-outer tcgen05 kernel of the BEV
                              necks (64..256 channels, W <= 16, strides (1,1,1) / (1,1,2))  */
#define DFM_CONV_TC_NECK_DHW 4 /* dfm_op_conv3d only: force the same kernel in [D][H][W]
                                  orientation, D cut into windows (stride 1, pad 1, Cin >= 64) */

/* output-selection flags for the *_host entry points */
#define DFM_OUT_COST 1   /* gated depth logits           [1,1,D,Ho,Wo] */
#define DFM_OUT_STEREO 2 /* stereo tower feature         [1,32,D,Ho,Wo] */
#define DFM_OUT_MONO 4   /* mono tower feature           [1,32,D,Ho,Wo] */

const char* dfm_last_error(void);
int dfm_version(void);
/* Reports the visible device; DFM_ERR_NOGPU when there is none. */
int dfm_device_info(int* sm_count, int* cc_major, int* cc_minor, long long* l2_bytes);

/* ------------------------------------------------------------------------------------
 * Geometry of one (cur, prev) pair -- the img_meta fields DfMBackbone.forward reads
 * (mmdet3d/models/backbones/dfm_backbone.py:150-172).
 * ---------------------------------------------------------------------------------- */
typedef struct dfm_geometry {
  double cam2img[16];  /* img_metas[0]['ori_cam2img'], 4x4 row-major              */
  double cur2prev[16]; /* img_metas[0]['cur2prevs'][0], 4x4 row-major            */
  double crop_x;       /* img_metas[0]['crop_offset'][0]                          */
  double crop_y;       /* img_metas[0]['crop_offset'][1]                          */
  double scale;        /* img_metas[0].get('scale_factor', [1.0])[0]             */
  double org_w;        /* img_metas[0]['ori_shape'][1] (used only when flipped)   */
  int flip;            /* img_metas[0].get('flip', False)                         */
  int reserved;
} dfm_geometry_t;

/* ------------------------------------------------------------------------------------
 * DfMBackbone  (replaces mmdet3d/models/backbones/dfm_backbone.py:14-314:
 * build_dfm_cost + dres0/dres1 + hourglass + depth-pred convs + mono/stereo gate)
 * ---------------------------------------------------------------------------------- */
typedef struct dfm_backbone dfm_backbone_t;

typedef struct dfm_backbone_desc {
  int in_channels;        /* DfMBackbone(in_channels=32)                       */
  int cv_channels;        /* cv_channels=32                                    */
  int feat_h, feat_w;     /* stereo feature size H x W (full image resolution) */
  int num_planes;         /* D = depth_cfg.num_bins / depth_cfg.downsample_factor */
  int cost_sample_factor; /* 4                                                 */
  int feat_sample_factor; /* 1                                                 */
  int conv_impl;          /* DFM_CONV_*                                        */
} dfm_backbone_desc_t;

int dfm_backbone_create(const dfm_backbone_desc_t* desc, dfm_backbone_t** out);
int dfm_backbone_destroy(dfm_backbone_t* bb);
/* Upload one parameter by its reference state_dict key (SURVEY.md 8a "State"), e.g.
 * "dres0.conv.weight" (32,64,3,3,3), "hg_stereo.0.conv5.0.weight" (ConvTranspose3d
 * layout in x out), "aggregate_cost.weight" (D,2D,1,1).  h_data: host fp32, reference
 * layout; the library repacks it for its kernels. */
int dfm_backbone_set_param(dfm_backbone_t* bb, const char* name, const float* h_data,
                           long long numel);
/* The injected attribute DfMBackbone.downsampled_depth (detectors/dfm.py:160-168). */
int dfm_backbone_set_depths(dfm_backbone_t* bb, const float* h_depths, int n);
/* Number of parameters still missing (0 = ready). */
int dfm_backbone_missing_params(const dfm_backbone_t* bb);
long long dfm_backbone_workspace_bytes(const dfm_backbone_t* bb);
/* DfMBackbone.forward (dfm_backbone.py:143-214), batch 1 (the reference supports
 * only B=1, :160).  d_cur/d_prev: [1,C,H,W] NCHW.  Outputs (any may be NULL):
 * d_cost [1,1,D,Ho,Wo], d_stereo / d_mono [1,32,D,Ho,Wo], NCDHW like the reference. */
int dfm_backbone_forward(dfm_backbone_t* bb, const float* d_cur, const float* d_prev,
                         const dfm_geometry_t* geom, float* d_cost, float* d_stereo,
                         float* d_mono, void* stream);
/* Same call for stereo features that are already channels-last [H][W][C] (what
 * dfm_stereo_tail_forward emits): skips the two NCHW -> NHWC transposes. */
int dfm_backbone_forward_cl(dfm_backbone_t* bb, const float* d_cur_cl, const float* d_prev_cl,
                            const dfm_geometry_t* geom, float* d_cost, float* d_stereo,
                            float* d_mono, void* stream);
/* Same call with HOST buffers: copies the two feature maps host->device, runs the
 * path, copies the outputs selected by out_flags (DFM_OUT_*) device->host, and
 * synchronises.  Buffers should be page-locked for full PCIe bandwidth. */
/* Optional: starts copying the NEXT pair host->device on a side stream and returns at once
 * (page-locked buffers).  A later dfm_backbone_forward_host with the same two host pointers
 * consumes the staged copy instead of copying again, so the transfer of pair i+1 overlaps the
 * processing of pair i.  Two pairs can be staged; the host buffers must stay unchanged until
 * the forward that consumes them returns. */
int dfm_backbone_prefetch_host(dfm_backbone_t* bb, const float* h_cur, const float* h_prev);
int dfm_backbone_forward_host(dfm_backbone_t* bb, const float* h_cur, const float* h_prev,
                              const dfm_geometry_t* geom, int out_flags, float* h_cost,
                              float* h_stereo, float* h_mono, void* stream);
/* Device pointer of the gated logits kept inside the handle after a forward
 * (lets a host-buffer caller chain dfm_depth_head_forward without a round trip). */
const float* dfm_backbone_cost_device(const dfm_backbone_t* bb);
/* Channels-last [D][Ho][Wo][cv] device copy of stereo_feat kept by the last forward (valid until
 * the next one): hand it to dfm_frustum_forward with DFM_LAYOUT_DHWC to skip a transpose. */
const float* dfm_backbone_stereo_feat_device(const dfm_backbone_t* bb);
/* Test hook: copies a named intermediate (channels-last [D][H][W][C]) to d_out.
 * Names: "raw0", "raw1", "c1".."c6", "p0", "logit", with suffix "_mono" for the mono tower
 * (which may hold the z-shortened volume). */
int dfm_backbone_debug_tensor(dfm_backbone_t* bb, const char* name, float* d_out,
                              long long numel, void* stream);
/* Synchronises `stream` and reports asynchronous failures of this library's kernels
 * (CUDA errors, or an mbarrier hand-over that timed out inside a tensor-core kernel). */
int dfm_sync_check(void* stream);
/* Per-kernel device timing: when enabled, every conv launch is bracketed by CUDA events
 * on its own stream.  dfm_profile_report synchronises the device, writes one JSON object
 * {"<kernel class>": {"launches": n, "ms": total, "flops": algorithmic}, ...} into buf
 * (NUL-terminated, truncated to cap) and clears the record. */
int dfm_profile_enable(int on);
int dfm_profile_report(char* buf, int cap);
/* Counters since creation: kernels launched by this library / of which tcgen05. */
int dfm_launch_counters(long long* launches, long long* tc_launches);

/* ------------------------------------------------------------------------------------
 * build_dfm_cost alone (dfm_backbone.py:217-314), materialising the reference's
 * [1,2C,D,Ho,Wo] NCDHW volume.  Parity/bring-up op: the backbone never calls it (the
 * volume is consumed on the fly), tests use it to pin rows a1/a9 against the oracle.
 * h_depths: host [D].
 * ---------------------------------------------------------------------------------- */
int dfm_op_build_cost_volume(const float* d_cur, const float* d_prev, int C, int H, int W,
                             const float* h_depths, int D, int cost_sample_factor,
                             int feat_sample_factor, const dfm_geometry_t* geom,
                             float* d_volume, void* stream);

/* ------------------------------------------------------------------------------------
 * Generic 3x3x3 conv3d / conv_transpose3d building block (the cuDNN calls behind
 * models/utils/conv_modules.py:27-43,104-127 and mmcv ConvModule(Conv3d)).
 * d_x: NCDHW [1,Cin,Di,Hi,Wi]; h_w: host weight in the reference layout
 * ((Cout,Cin,3,3,3) or, transposed, (Cin,Cout,3,3,3)); d_y: NCDHW output.
 * stride/pad per (D,H,W); transposed uses stride 2, pad 1, output_padding 1.
 * ---------------------------------------------------------------------------------- */
int dfm_op_conv3d(const float* d_x, int Cin, int Di, int Hi, int Wi, const float* h_w,
                  int Cout, const int stride[3], const int pad[3], int transposed,
                  int conv_impl, float* d_y, void* stream);

/* ------------------------------------------------------------------------------------
 * DepthHead.forward with with_convs=False (mmdet3d/models/dense_heads/depth_head.py:
 * 190-212): x`factor` trilinear upsample (align_corners) -> softmax over depth ->
 * expectation over d_depth_samples [factor*D].  d_cost [1,1,D,Ho,Wo].
 * d_volume / d_softmax [1,1,fD,fHo,fWo] and d_preds [1,1,fHo,fWo]; any may be NULL
 * (skipping the two 4-D volumes is the fast path when only depth_preds is consumed).
 * ---------------------------------------------------------------------------------- */
int dfm_depth_head_forward(const float* d_cost, const float* d_depth_samples, int D, int Ho,
                           int Wo, int factor, float* d_volume, float* d_softmax,
                           float* d_preds, void* stream);

/* ------------------------------------------------------------------------------------
 * MultiViewDfM.feature_transformation lifting step (mmdet3d/models/detectors/
 * multiview_dfm.py:119-209 calling fusion_layers/point_fusion.py:14-106 with
 * aligned=False, valid_flag=True), one sample: nearest-tap gather of every voxel
 * centre into every (frame, view), valid-count averaging, temporal 'mean' or 'concat'.
 * d_feats [T*Nv, C, Hf, Wf]; h_lidar2img [T*Nv][16] row-major; the voxel-centre
 * coordinates per axis are passed in (the caller computes them exactly as
 * AlignedAnchor3DRangeGenerator does, core/anchor/anchor_3d_generator.py:283-310),
 * points are ordered z-major, then y, then x (the generator's permute at :327).  d_volume: [C*(concat?T:1), Nx, Ny, Nz].
 * ---------------------------------------------------------------------------------- */
typedef struct dfm_lift_desc {
  int num_frames, num_views, channels, feat_h, feat_w;
  int n_voxels[3];            /* Nx, Ny, Nz                                  */
  float scale_x, scale_y;     /* img_meta['scale_factor'][:2]                */
  float crop_x, crop_y;       /* img_meta['img_crop_offset']                 */
  int flip;
  int input_h, input_w;       /* img_meta['input_shape']                     */
  int concat;                 /* temporal_aggregate == 'concat'              */
} dfm_lift_desc_t;
int dfm_multiview_lift(const dfm_lift_desc_t* desc, const float* d_feats,
                       const double* h_lidar2img, const int* h_img_w /* [T*Nv] img_shape w */,
                       const float* h_xs /* [Nx] */, const float* h_ys /* [Ny] */,
                       const float* h_zs /* [Nz] voxel-centre coordinates */,
                       float* d_volume, void* stream);
/* Same lifting, channels-last output [Nx][Ny][Nz][C*(concat?T:1)] -- the layout the necks' conv
 * loaders read (dfm_neck_forward_cl): the voxel's channels are one contiguous row, written by one
 * coalesced warp store, and the neck's NCDHW -> channels-last pass disappears.  Bit-identical
 * values.  (C == 64 and (concat or T == 1) run the fused kernel; other configurations run the
 * reference-layout kernel plus one transpose.) */
int dfm_multiview_lift_cl(const dfm_lift_desc_t* desc, const float* d_feats,
                          const double* h_lidar2img, const int* h_img_w, const float* h_xs,
                          const float* h_ys, const float* h_zs, float* d_volume_cl,
                          void* stream);

/* ------------------------------------------------------------------------------------
 * DfMNeck / OutdoorImVoxelNeck, eval mode (mmdet3d/models/necks/dfm_neck.py:10-122,
 * imvoxel_neck.py:8-117): BatchNorm3d folded into per-channel affine.
 * ---------------------------------------------------------------------------------- */
typedef struct dfm_neck dfm_neck_t;
typedef struct dfm_neck_desc {
  int in_channels;  /* per-frame channels (64)                               */
  int out_channels; /* 256                                                   */
  int num_frames;   /* DfMNeck: stereo tower sees in_channels*num_frames; 0 => OutdoorImVoxelNeck */
  int nx, ny, nz;
  int conv_impl;
} dfm_neck_desc_t;
int dfm_neck_create(const dfm_neck_desc_t* desc, dfm_neck_t** out);
int dfm_neck_destroy(dfm_neck_t* neck);
/* Reference state_dict keys: "mono_layers.0.conv0.conv.weight", "...bn.weight/bias/
 * running_mean/running_var", "stereo_layers...", "aggregate_layer.weight"; for
 * OutdoorImVoxelNeck the prefix is "model.". */
int dfm_neck_set_param(dfm_neck_t* neck, const char* name, const float* h_data,
                       long long numel);
int dfm_neck_missing_params(const dfm_neck_t* neck);
/* d_x [1, Cin_total, Nx, Ny, Nz] -> d_bev [1, out_channels, Ny, Nx]. */
int dfm_neck_forward(dfm_neck_t* neck, const float* d_x, float* d_bev, void* stream);
/* d_x_cl: channels-last [Nx][Ny][Nz][C*T] (dfm_multiview_lift_cl's output) */
int dfm_neck_forward_cl(dfm_neck_t* neck, const float* d_x_cl, float* d_bev, void* stream);

/* ------------------------------------------------------------------------------------
 * FrustumToVoxel (mmdet3d/models/necks/feature_transformation.py:12-173), the stage that
 * consumes the hot path's outputs: the pseudo-lidar voxel grid
 * (detectors/dfm.py:174-211) is projected with cam2img[:3] (:175-187), stereo_feat / the depth
 * distribution / cur_sem_feats are grid-sampled there (:127-158), then voxel_convs
 * (Conv3d k3 + GroupNorm(32) + ReLU, :49-62) and AvgPool3d((4,1,1)) (:63,166).
 * ---------------------------------------------------------------------------------- */
#define DFM_LAYOUT_NCDHW 0 /* [C][D][H][W], the reference tensor layout   */
#define DFM_LAYOUT_DHWC 1  /* channels-last [D][H][W][C]                   */
typedef struct dfm_frustum dfm_frustum_t;
typedef struct dfm_frustum_desc {
  int num_3dconvs;       /* feature_transformation.py:17 (1..4)            */
  int cv_channels;       /* 32                                              */
  int out_channels;      /* 32                                              */
  int in_sem_channels;   /* 32                                              */
  int sem_atten_feat, stereo_atten_feat, cat_img_feature; /* :20-22        */
  int num_planes;        /* D of stereo_feat [1, cv, D, feat_h, feat_w]     */
  int feat_h, feat_w;
  int sem_h, sem_w;      /* cur_sem_feats [1, sem, sem_h, sem_w]            */
  int depth_factor;      /* softmax volume is [f*D][f*feat_h][f*feat_w]     */
  int nx, ny, nz;        /* voxel grid; coordinates_3d is [nz][ny][nx][3]   */
  float depth_min, depth_max; /* depth_cfg                                  */
  int conv_impl;         /* DFM_CONV_*                                      */
} dfm_frustum_desc_t;
/* h_xs [nx], h_ys [ny], h_zs [nz]: the separable pseudo-lidar voxel centres of
 * coordinates_3d (x = [0,0,:,0], y = [0,:,0,1], z = [:,0,0,2]). */
int dfm_frustum_create(const dfm_frustum_desc_t* desc, const float* h_xs, const float* h_ys,
                       const float* h_zs, dfm_frustum_t** out);
int dfm_frustum_destroy(dfm_frustum_t* f);
/* Reference state_dict keys: "voxel_convs.<i>.0.conv.weight", "voxel_convs.<i>.0.gn.weight",
 * "voxel_convs.<i>.0.gn.bias". */
int dfm_frustum_set_param(dfm_frustum_t* f, const char* name, const float* h_data,
                          long long numel);
int dfm_frustum_missing_params(const dfm_frustum_t* f);
/* One sample.  d_stereo_feat: device, layout as flagged.  The depth distribution is either the
 * materialised DepthHead output d_softmax [f*D][f*H][f*W] (the reference's argument) or, with
 * d_softmax == NULL, rebuilt on the fly from the low-res logits d_cost_logits [D][H][W]
 * (DfMBackbone's first output) so the 0.5-0.9 GB volume never exists; that pass is
 * DepthHead.forward's reduction, so with d_depth_samples [f*D] and d_depth_preds [f*H][f*W]
 * (both optional) it also returns the DepthHead's depth_preds and no separate
 * dfm_depth_head_forward call is needed.  d_sem: [sem][sem_h][sem_w] (NULL when
 * !cat_img_feature).  cam2img: 16 doubles, row-major img_meta['cam2img'].
 * pad_h/pad_w: img_metas[0]['pad_shape'].  d_out: [out_channels][nz/4][ny][nx]. */
int dfm_frustum_forward(dfm_frustum_t* f, const float* d_stereo_feat, int stereo_layout,
                        const float* d_softmax, const float* d_cost_logits,
                        const float* d_depth_samples, float* d_depth_preds, const float* d_sem,
                        const double* cam2img, int pad_h, int pad_w, float* d_out,
                        void* stream);

/* ------------------------------------------------------------------------------------
 * The hot-path segment of DfM.simple_test as one call with HOST buffers
 * (mmdet3d/models/detectors/dfm.py:296 `backbone_stereo(...)`, :420 `depth_head(...)`,
 * :423-425 `feature_transformation(...)`): (cur, prev) stereo features + cur semantic
 * features in, what the BEV stage consumes out -- voxel features [out][nz/4][ny][nx] and
 * DepthHead's depth_preds [fH][fW] (optionally the gated logits [D][Ho][Wo]).  Host->device
 * copies, the whole path and device->host copies run on `stream`; the call synchronises.
 * The pair may have been staged with dfm_backbone_prefetch_host.  stereo_feat never leaves
 * the device and the x4-upsampled DepthHead volumes are never built.  `fr` must have been
 * created for the backbone's volume shape (num_planes, feat_h = Ho, feat_w = Wo).
 * h_depth_samples: host [depth_factor * D] (DepthHead.depth_samples).
 * ---------------------------------------------------------------------------------- */
int dfm_pipeline_forward_host(dfm_backbone_t* bb, dfm_frustum_t* fr, const float* h_cur,
                              const float* h_prev, const float* h_sem,
                              const dfm_geometry_t* geom, const double* cam2img, int pad_h,
                              int pad_w, const float* h_depth_samples, float* h_voxel,
                              float* h_depth_preds, float* h_cost, void* stream);

/* ------------------------------------------------------------------------------------
 * The 2-D BEV stage behind FrustumToVoxel (north_star's "3D box
 * regressions"): BEVHourglass (mmdet3d/models/backbones/bev_hourglass.py:11-137, GroupNorm
 * variant of configs/dfm/dfm_r34_1x8_kitti-3d-3class.py:146-150) and LIGAAnchor3DHead's
 * forward (mmdet3d/models/dense_heads/liga_anchor3d_head.py:37-128).  All tensors NCHW fp32.
 * ---------------------------------------------------------------------------------- */
typedef struct dfm_bev_hourglass dfm_bev_hourglass_t;
typedef struct dfm_bev_desc {
  int in_channels;  /* 160 = 32 channels x 5 height slices (detectors/dfm.py:427-428) */
  int out_channels; /* 64                                                            */
  int ny, nx;       /* BEV grid, both multiples of 4 (304 x 288)                      */
  int conv_impl;    /* DFM_CONV_*                                                     */
} dfm_bev_desc_t;
int dfm_bev_hourglass_create(const dfm_bev_desc_t* desc, dfm_bev_hourglass_t** out);
int dfm_bev_hourglass_destroy(dfm_bev_hourglass_t* b);
/* Reference state_dict keys: "compress_conv.conv.weight", "compress_conv.gn.{weight,bias}",
 * "bev_hourglass.conv1.0.0.weight", "bev_hourglass.conv1.0.1.{weight,bias}",
 * "bev_hourglass.conv2.0.weight", ..., "bev_hourglass.conv5.0.weight" (ConvTranspose2d layout
 * in x out), "bev_hourglass.conv6.1.bias". */
int dfm_bev_hourglass_set_param(dfm_bev_hourglass_t* b, const char* name, const float* h_data,
                                long long numel);
int dfm_bev_hourglass_missing_params(const dfm_bev_hourglass_t* b);
/* BEVHourglass.forward (:39-50): d_x [in][ny][nx] -> d_prehg (optional) and d_out, both
 * [out][ny][nx] (spatial_features_2d_prehg, spatial_features_2d). */
int dfm_bev_hourglass_forward(dfm_bev_hourglass_t* b, const float* d_x, float* d_prehg,
                              float* d_out, void* stream);

typedef struct dfm_anchor_head dfm_anchor_head_t;
typedef struct dfm_anchor_head_desc {
  int in_channels, feat_channels; /* 64, 64                                            */
  int num_convs;                  /* cls_convs / reg_convs depth (2)                   */
  int cls_channels;               /* num_anchors * num_classes      (18)               */
  int reg_channels;               /* num_anchors * box_code_size    (42)               */
  int dir_channels;               /* num_anchors * 2, 0 without direction classifier   */
  int ny, nx;
  int conv_impl;
} dfm_anchor_head_desc_t;
int dfm_anchor_head_create(const dfm_anchor_head_desc_t* desc, dfm_anchor_head_t** out);
int dfm_anchor_head_destroy(dfm_anchor_head_t* h);
/* Keys: "cls_convs.<i>.conv.weight", "cls_convs.<i>.gn.{weight,bias}", "reg_convs.<i>...",
 * "conv_cls.{weight,bias}", "conv_reg.{weight,bias}", "conv_dir_cls.{weight,bias}". */
int dfm_anchor_head_set_param(dfm_anchor_head_t* h, const char* name, const float* h_data,
                              long long numel);
int dfm_anchor_head_missing_params(const dfm_anchor_head_t* h);
/* LIGAAnchor3DHead.forward_single (:108-128): d_x [64][ny][nx] -> cls_score
 * [cls_channels][ny][nx], bbox_pred [reg_channels][ny][nx], dir_cls_preds
 * [dir_channels][ny][nx]. */
int dfm_anchor_head_forward(dfm_anchor_head_t* h, const float* d_x, float* d_cls, float* d_bbox,
                            float* d_dir, void* stream);

/* ------------------------------------------------------------------------------------
 * voxel_sample (mmdet3d/models/fusion_layers/point_fusion.py:324-410): frustum-from-voxel
 * resampling for an optional depth head of MultiViewDfM (detectors/multiview_dfm.py:220-256;
 * no shipped config enables it).  d_voxel [C][Nx][Ny][Nz] -> d_out [C][num_depths][out_h][out_w],
 * out_h/out_w = round(img_pad_shape / downsample_factor); h_depths: host [num_depths] =
 * depth_samples[::downsample_factor]; h_proj: 16 doubles, the (lidar/cam)2img matrix.
 * grid_sample semantics: zeros padding, align_corners=True, trilinear (aligned) or nearest.
 * ---------------------------------------------------------------------------------- */
typedef struct dfm_voxel_sample_desc {
  int channels, nx, ny, nz;
  float voxel_range[6];
  float voxel_size[3];
  int num_depths, out_h, out_w, downsample_factor;
  float scale_x, scale_y;  /* img_scale_factor */
  float crop_x, crop_y;    /* img_crop_offset  */
  int flip, img_w;         /* img_flip, img_shape[1] */
  int aligned;             /* 1: trilinear, 0: nearest */
} dfm_voxel_sample_desc_t;
int dfm_voxel_sample(const dfm_voxel_sample_desc_t* desc, const float* d_voxel,
                     const float* h_depths, const double* h_proj, float* d_out, void* stream);

/* Asynchronous form of the same call: submit enqueues the host->device copies, the path and
 * the device->host copies of the outputs (those on a side stream, so they overlap the NEXT
 * frame's compute) and returns at once; dfm_pipeline_wait blocks until the outputs of the
 * oldest submitted frame are in host memory and reports asynchronous failures.  At most two
 * frames may be in flight; host input buffers must stay unchanged until the wait for their
 * frame returns, output buffers must not be read before it. */
int dfm_pipeline_submit_host(dfm_backbone_t* bb, dfm_frustum_t* fr, const float* h_cur,
                             const float* h_prev, const float* h_sem,
                             const dfm_geometry_t* geom, const double* cam2img, int pad_h,
                             int pad_w, const float* h_depth_samples, float* h_voxel,
                             float* h_depth_preds, void* stream);
int dfm_pipeline_wait(dfm_backbone_t* bb);
/* dfm_backbone_prefetch_host for the pipeline: also stages the NEXT frame's sem features
 * (`sem_numel` floats, may be NULL / 0) together with its pair on the side stream.  A small
 * host->device copy issued at submit time would queue on the copy engine behind this bulk
 * copy and stall the compute stream (measured: ~1 ms per frame), so everything a frame reads
 * from the host should be handed over here, one frame ahead. */
int dfm_pipeline_prefetch_host(dfm_backbone_t* bb, const float* h_cur, const float* h_prev,
                               const float* h_sem, long long sem_numel);

/* ------------------------------------------------------------------------------------
 * The tail of SPPUNetNeck (mmdet3d/models/necks/spp_unet_neck.py:60-75 `lastconv`, applied at
 * :110): Conv2d 3x3 (32->32) + GroupNorm(32) + ReLU + Conv2d 1x1
 * (32->32, no bias) producing the full-resolution stereo feature that build_dfm_cost samples.
 * d_x [32][H][W] (the up-convolved feature, NCHW) -> d_out_cl [H][W][32] channels-last (feed it
 * to dfm_backbone_forward_cl) and / or d_out_nchw [32][H][W] (the reference's return value).
 * Keys: "lastconv.0.conv.weight" (32,32,3,3), "lastconv.0.gn.weight", "lastconv.0.gn.bias",
 * "lastconv.1.weight" (32,32,1,1).
 * ---------------------------------------------------------------------------------- */
typedef struct dfm_stereo_tail dfm_stereo_tail_t;
int dfm_stereo_tail_create(int H, int W, int conv_impl, dfm_stereo_tail_t** out);
int dfm_stereo_tail_destroy(dfm_stereo_tail_t* t);
int dfm_stereo_tail_set_param(dfm_stereo_tail_t* t, const char* name, const float* h_data,
                              long long numel);
int dfm_stereo_tail_missing_params(const dfm_stereo_tail_t* t);
int dfm_stereo_tail_forward(dfm_stereo_tail_t* t, const float* d_x, float* d_out_cl,
                            float* d_out_nchw, void* stream);

/* Re-entrancy: handles may live on different devices and be driven from different host
 * threads only if each thread owns its device; per-device scratch (K-slice partial sums,
 * lifting staging, the host-copy side stream) and the profiling record are shared by all
 * handles of a device and are NOT locked -- the same one-thread-per-process model as the
 * reference (tools/slurm_train.sh:15-24, SURVEY.md section 8b "Threading"). */

#ifdef __cplusplus
}
#endif
#endif /* DFM_B200_H_ */
