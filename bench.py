#!/usr/bin/env python
"""bench.py -- frames/s of the DfM plane-sweep cost-volume path (BASELINE.json metric).

A step = one pass of the hot path (DfMBackbone: warp + volume + 3-D aggregation +
gate, then DepthHead) over one synthetic KITTI-shape pair: 370x1224 padded to
384x1248, D=112 planes (BASELINE.json configs[1]).

  python bench.py --gpus N --steps K --warmup W          # the CUDA path
  python bench.py --impl reference ...                   # the reference's CPU path

N > 1 is launched by torchrun, one rank per GPU: the pairs are sharded over ranks
(independent frames; the reference cannot batch, dfm_backbone.py:160), no data-path
collective, "scaling": "weak".

One JSON line on stdout (rank 0).  Keys beyond the base contract:
  roofline      dominant kernel (tcgen05 3x3x3 conv, 32->32 full resolution): algorithmic
                FLOPs per launch / mean launch duration (CUDA events inside the timed
                region, on the launching stream) against the measured bf16 peak
  cpu_baseline  the oracle (PyTorch-CPU port of the reference) timed on this host
  e2e           same metric through the C-ABI host-buffer entry point, pinned host
                buffers, H2D + D2H copies inside the timed region
"""
import argparse
import copy
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

H, W, D, C = 384, 1248, 112, 32
HO, WO = H // 4, W // 4
ORI_SHAPE = (370, 1224, 3)
V = D * HO * WO
FLOPS_PER_FRAME = 521856.0 * V
IO_BYTES_PER_FRAME = 2 * 32 * H * W * 4 + 5.25e6 + 33 * V * 4
WORKLOAD = 'dfm_r34_1x8_kitti-3d-3class D=112 384x1248 batch=1'
NCU_DOMINANT_TRAFFIC_BYTES = 973.3e6   # 587.2 MB read + 386.1 MB written, profiles/r01_ncu_conv_tc.csv


def peaks():
    p = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.isfile(p):
        j = json.load(open(p))
        return dict(bf16=j.get('bf16_tflops_sustained', j.get('bf16_tflops')),
                    hbm=j.get('hbm_gbs'), src='measured')
    return dict(bf16=1400.0, hbm=6650.0, src='fallback')  # B200_PROFILING.md


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region."""

    def __init__(self, index):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        q = ('clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,'
             'clocks_event_reasons.hw_thermal_slowdown,'
             'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap')
        try:
            self.proc = subprocess.Popen(
                ['nvidia-smi', f'--id={self.index}', f'--query-gpu={q}',
                 '--format=csv,noheader,nounits', '-lms', '100'],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(',')])

    def stop(self):
        if self.proc is None:
            return None
        time.sleep(0.15)
        self.proc.terminate()
        rows = [r for r in self.rows if len(r) >= 7 and r[0].isdigit()]
        if not rows:
            return None
        sm = [int(r[0]) for r in rows]
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        reasons = [n for i, n in enumerate(names) if any(r[3 + i] == 'Active' for r in rows)]
        return dict(sm_mhz=int(statistics.median(sm)), sm_max_mhz=int(rows[0][1]),
                    samples=len(rows), reasons=reasons)


def oracle_frame_seconds(planes, threads, repeats=1):
    """Time of one reference-path frame (backbone + depth head) on the host cores."""
    import torch

    from depth_from_motion_b200 import synthetic as syn
    from oracle import dfm_oracle as O
    torch.set_num_threads(threads)
    cur, prev, metas, params = syn.make_kitti_pair(0, H, W, planes, ori_shape=ORI_SHAPE)
    cfg = syn.depth_cfg_for(planes)
    ts = []
    with torch.no_grad():
        for _ in range(repeats):
            t0 = time.perf_counter()
            cost, _, _ = O.dfm_backbone_forward(params, cur, prev, metas, cfg)
            O.depth_head_forward(cost, O.depth_samples(cfg))
            ts.append(time.perf_counter() - t0)
    return statistics.median(ts)


def run_reference(args):
    """--impl reference: the reference's own PyTorch CPU path (oracle port; the
    reference has no native code to compile) on the host cores."""
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    slab = 16
    t_full = oracle_frame_seconds(D, cores)           # one whole D=112 frame
    t_slab0 = oracle_frame_seconds(slab, cores)       # also warms the slab path
    ratio = t_full / t_slab0
    for _ in range(max(args.warmup - 1, 0)):
        oracle_frame_seconds(slab, cores)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        oracle_frame_seconds(slab, cores)
    per_step = (time.perf_counter() - t0) / args.steps
    fps = 1.0 / (per_step * ratio)
    line = dict(
        impl='reference', metric='frames/sec', value=fps, unit='frames/s', n_gpus=args.gpus,
        steps=args.steps, warmup=args.warmup, ms_per_step=per_step * ratio * 1e3,
        higher_is_better=True, scaling='weak', vs_baseline=None, dtype='f32',
        data='synthetic', config=dict(workload=WORKLOAD),
        cpu_baseline=dict(value=fps, unit='frames/s', cores=cores, kind='port',
                          sample=f'each step = the same pair at D={slab} planes '
                                 f'({per_step:.2f} s), scaled by the measured full-frame/'
                                 f'slab time ratio {ratio:.2f} (one D=112 frame: {t_full:.1f} s)'),
        e2e=dict(value=fps, unit='frames/s', h2d_bytes_per_step=0, d2h_bytes_per_step=0))
    print(json.dumps(line), flush=True)


def run_ours(args):
    import torch
    import torch.distributed as dist

    from depth_from_motion_b200 import capi, modules
    from depth_from_motion_b200 import synthetic as syn

    rank = int(os.environ.get('RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    if not torch.cuda.is_available():
        raise SystemExit('bench.py needs a B200: there is no CPU path (use --impl reference)')
    torch.cuda.set_device(local)
    if world > 1:
        os.environ['NCCL_DEBUG'] = 'WARN'  # keep NCCL's version banner off stdout (one JSON line)
        dist.init_process_group('nccl', device_id=torch.device('cuda', local))
    capi.lib()

    # two different pairs per rank so consecutive steps never re-read the same inputs;
    # the per-step working set (~7 GB of activations) is far larger than the 126 MB L2
    pairs = []
    for i in range(2):
        cur, prev, metas, params = syn.make_kitti_pair(100 + 2 * rank + i, H, W, D,
                                                       ori_shape=ORI_SHAPE)
        pairs.append((cur.cuda(), prev.cuda(), metas, cur.pin_memory(), prev.pin_memory()))
    cfg = syn.depth_cfg_for(D)
    model = modules.DfMBackbone(in_channels=C, depth_cfg=cfg).cuda().eval()
    model.load_state_dict(params, strict=True)
    model.downsampled_depth = _depths(cfg, 4)
    head = modules.DepthHead(
        depth_cfg=dict(mode='UD', num_bins=cfg['num_bins'], min_depth=2, max_depth=59.6),
        with_convs=False, num_views=1, depth_loss=dict(type='ce', loss_weight=1.0))
    head.depth_samples = _depths(cfg, 1)
    head.downsample_factor = 4

    def step(i):
        cur, prev, metas, _, _ = pairs[i % 2]
        cost, stereo, mono = model(cur, prev, metas)
        return head(cost)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    with torch.no_grad():
        for i in range(args.warmup):
            step(i)
        capi.sync_check()
        l0, tc0 = capi.launch_counters()
        capi.profile_enable(True)
        capi.profile_report()
        sampler = ClockSampler(local)
        barrier()
        if rank == 0:
            sampler.start()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(args.steps):
            step(i)
        e1.record()
        barrier()
        clocks = sampler.stop() if rank == 0 else None
        ms = e0.elapsed_time(e1)
        prof = capi.profile_report()
        capi.profile_enable(False)
        l1, tc1 = capi.launch_counters()
        capi.sync_check()
    from depth_from_motion_b200.sharding import reduce_step_time
    ms_total = reduce_step_time(ms, 'cuda')
    fps = world * args.steps / (ms_total * 1e-3)

    # ---- e2e: host buffers through the C-ABI, copies inside the timed region ----
    import ctypes
    L = capi.lib()
    h_cost = torch.empty((1, 1, D, HO, WO)).pin_memory()
    h_pred = torch.empty((1, 1, H, W)).pin_memory()
    d_pred = torch.empty((1, 1, H, W), device='cuda')
    samples_dev = head.depth_samples.cuda()

    def prefetch(i):
        _, _, _, hc, hp = pairs[i % 2]
        capi.check(L.dfm_backbone_prefetch_host(
            model._handle, ctypes.c_void_p(hc.data_ptr()), ctypes.c_void_p(hp.data_ptr())),
            'dfm_backbone_prefetch_host')

    use_prefetch = [True]

    def e2e_step(i):
        _, _, metas, hc, hp = pairs[i % 2]
        g = modules.geometry_from_meta(metas[0])
        stream = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
        # every step starts the host->device copy of the NEXT step's pair (side stream), then
        # processes its own pair, whose copy was started one step earlier: one pair copied per
        # step, inside the timed region, overlapped with compute
        if use_prefetch[0]:
            prefetch(i + 1)
        capi.check(L.dfm_backbone_forward_host(
            model._handle, ctypes.c_void_p(hc.data_ptr()), ctypes.c_void_p(hp.data_ptr()),
            ctypes.byref(g), capi.DFM_OUT_COST, ctypes.c_void_p(h_cost.data_ptr()), None, None,
            stream), 'dfm_backbone_forward_host')
        capi.check(L.dfm_depth_head_forward(
            L.dfm_backbone_cost_device(model._handle), ctypes.c_void_p(samples_dev.data_ptr()),
            D, HO, WO, 4, None, None, ctypes.c_void_p(d_pred.data_ptr()), stream),
            'dfm_depth_head_forward')
        h_pred.copy_(d_pred, non_blocking=True)
        torch.cuda.current_stream().synchronize()

    # self-check of the prefetched path against the plain call (same CUDA kernels either way);
    # on any disagreement the e2e loop copies synchronously inside forward_host instead
    use_prefetch[0] = False
    e2e_step(0)
    ref_cost = h_cost.clone()
    try:
        prefetch(0)
        use_prefetch[0] = True
        e2e_step(0)
        if not torch.allclose(h_cost, ref_cost, rtol=1e-5, atol=1e-6):
            raise RuntimeError('prefetched result differs')
    except RuntimeError as exc:
        print(f'[bench] prefetch path disabled: {exc}', file=sys.stderr)
        use_prefetch[0] = False
    if use_prefetch[0]:
        prefetch(0)
    nwarm = min(args.warmup, 3)
    for i in range(nwarm):
        e2e_step(i)
    barrier()
    t0 = time.perf_counter()
    ee0, ee1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ee0.record()
    for i in range(nwarm, nwarm + args.steps):  # step indices continue: pair i was prefetched
        e2e_step(i)
    ee1.record()
    barrier()
    e2e_ms = reduce_step_time(max(ee0.elapsed_time(ee1), (time.perf_counter() - t0) * 1e3),
                              'cuda')
    e2e_fps = world * args.steps / (e2e_ms * 1e-3)
    h2d = 2 * C * H * W * 4
    d2h = (D * HO * WO + H * W) * 4

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    pk = peaks()
    # dominant kernel: the full-resolution 32->32 tensor-core conv
    dom_key = f'conv_tc<32->32,s1,src>@{D}x{HO}x{WO}'
    roof = None
    tc_ms = sum(v['ms'] for k, v in prof.items() if k.startswith('conv_tc'))
    conv_ms = sum(v['ms'] for k, v in prof.items() if k.startswith('conv_'))
    if dom_key in prof:
        r = prof[dom_key]
        per_launch_flops = r['flops'] / r['launches']
        per_launch_s = r['ms'] * 1e-3 / r['launches']
        ach = per_launch_flops / per_launch_s / 1e12
        roof = dict(bound='tensor', kernel=dom_key, achieved=round(ach, 2), peak=pk['bf16'],
                    unit='TFLOP/s', frac=round(ach / pk['bf16'], 4),
                    # dram__bytes_read.sum + dram__bytes_write.sum of this kernel, one launch,
                    # from the committed ncu --set full capture (profiles/r01_ncu_conv_tc.csv);
                    # algorithmic bytes: 429.4 MB in + 429.4 MB out
                    traffic=NCU_DOMINANT_TRAFFIC_BYTES,
                    executed_bf16_tflops=round(3 * ach, 1),
                    executed_frac_of_peak=round(3 * ach / pk['bf16'], 4),
                    peak_source=pk['src'] + ' bf16 dense (sustained)',
                    launches_per_step=r['launches'] / args.steps,
                    ms_per_launch=round(per_launch_s * 1e3, 4),
                    note='achieved = algorithmic fp32 conv FLOPs (2*V*27*Cin*Cout); the kernel '
                         'executes 3 bf16 MMAs per product (hi/lo split) for fp32-class '
                         'accuracy, so the tensor pipe runs at 3x this rate',
                    share_of_step=round(r['ms'] / ms_total, 4))
    else:
        simt = {k: v for k, v in prof.items() if k.startswith('conv_simt')}
        if simt:
            k, r = max(simt.items(), key=lambda kv: kv[1]['ms'])
            ach = r['flops'] / (r['ms'] * 1e-3) / 1e12
            roof = dict(bound='tensor', kernel=k, achieved=round(ach, 2), peak=pk['bf16'],
                        unit='TFLOP/s', frac=round(ach / pk['bf16'], 4), traffic=None,
                        peak_source=pk['src'])
    cores = os.cpu_count() or 1
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        t_cpu = oracle_frame_seconds(D, cores)
        cpu = dict(value=1.0 / t_cpu, unit='frames/s', cores=cores, kind='port',
                   sample=f'one whole frame of this workload (D={D}, {H}x{W}) through the '
                          f'oracle (PyTorch-CPU restatement of the reference path), '
                          f'{t_cpu:.1f} s')
    line = dict(
        metric='frames/sec', value=fps, unit='frames/s', n_gpus=world, steps=args.steps,
        warmup=args.warmup, ms_per_step=ms_total / args.steps, higher_is_better=True,
        scaling='weak', vs_baseline=None, dtype='f32 (bf16x2 split operands, fp32 accumulate)',
        data='synthetic',
        config=dict(workload=WORKLOAD, pairs_per_step=world, planes=D, feature_hw=[H, W],
                    l2='per-step working set ~7 GB >> 126 MB L2; two input pairs alternate',
                    outputs='cost + stereo_feat + mono_feat + DepthHead(volume, softmax, preds)'),
        clocks=clocks,
        e2e=dict(value=e2e_fps, unit='frames/s', h2d_bytes_per_step=h2d,
                 d2h_bytes_per_step=d2h,
                 prefetch=use_prefetch[0],
                 what='dfm_backbone_prefetch_host(next pair) + dfm_backbone_forward_host (pinned '
                      'host features in, logits out) + dfm_depth_head_forward '
                      '(depth_preds out); stereo_feat stays on device for the next stage'),
        gpu_launches=l1 - l0, tc_launches=tc1 - tc0,
        roofline=roof,
        tensor=dict(achieved_tflops=round(FLOPS_PER_FRAME * fps / world / 1e12, 2),
                    frac_of_bf16_peak=round(FLOPS_PER_FRAME * fps / world / 1e12 / pk['bf16'], 4)),
        hbm=dict(compulsory_gbs=round(IO_BYTES_PER_FRAME * fps / world / 1e9, 1),
                 frac_of_peak=round(IO_BYTES_PER_FRAME * fps / world / 1e9 / pk['hbm'], 4)),
        conv_ms_per_step=round(conv_ms / args.steps, 3),
        tc_conv_ms_per_step=round(tc_ms / args.steps, 3),
        kernels={k: dict(launches=v['launches'] // args.steps,
                         ms=round(v['ms'] / args.steps, 4),
                         tflops=round(v['flops'] / max(v['ms'], 1e-9) / 1e9, 1))
                 for k, v in sorted(prof.items(), key=lambda kv: -kv[1]['ms'])},
        cpu_baseline=cpu)
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def _depths(cfg, ds):
    import torch
    nb = cfg['num_bins']
    interval = (cfg['depth_max'] - cfg['depth_min']) / nb
    d = torch.zeros(nb // ds, dtype=torch.float32)
    for i in range(nb // ds):
        d[i] = (i + 0.5) * ds * interval + cfg['depth_min']
    return d


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=20)
    ap.add_argument('--warmup', type=int, default=5)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--no-cpu-baseline', action='store_true')
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    if args.impl == 'reference':
        run_reference(args)
    else:
        run_ours(args)


if __name__ == '__main__':
    main()
