#!/usr/bin/env python
"""Precision tax of the tensor-core convs, measured with the oracle.

The CUDA kernels evaluate every conv product as x_hi*w_hi + x_lo*w_hi + x_hi*w_lo (bf16 pairs,
fp32 accumulate): 3 bf16 MMAs per product.  Cheaper candidates per layer:
    tf32   one kind::tf32 MMA, both operands rounded to 10 mantissa bits  (2 bf16-MMA units)
    w_bf16 x_hi*w_hi + x_lo*w_hi: weights rounded to bf16, activations 16 bits (2 units)
    x_bf16 x_hi*w_hi + x_hi*w_lo: activations rounded to bf16, weights 16 bits (2 units)
For each of the 22 conv layers this script applies ONE candidate to that layer only and reports
the normalised max-norm error of the gated depth logits (the north_star metric) against the
fp32 oracle, then the greedy largest layer set that stays under a 5e-4 budget.
CPU only; prints a markdown table (pasted into DESIGN.md)."""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from depth_from_motion_b200 import synthetic as syn  # noqa: E402
from oracle import dfm_oracle as O  # noqa: E402

LAYERS = []
for sfx, hg, pr in (('', 'hg_stereo.0', 'pred_stereo.0'), ('_mono', 'hg_mono.0', 'pred_mono.0')):
    LAYERS += ['dres0' + sfx, 'dres1' + sfx] + [f'{hg}.conv{i}' for i in range(1, 7)] + \
        [pr + '.0', pr + '.1']


def split16(x):
    """bf16 hi + bf16 lo: what the kernels' 3-term scheme keeps of an operand."""
    hi = O.bf16_round(x)
    return hi + O.bf16_round(x - hi)


class LayerQ:
    def __init__(self, table):
        self.table = table   # layer -> mode

    def round(self, t, layer, kind):
        mode = self.table.get(layer, 'split3')
        if mode == 'fp32':
            return t
        if mode == 'split3':
            return split16(t)
        if mode == 'tf32':
            return O.tf32_round(t)
        if mode == 'w_bf16':
            return O.bf16_round(t) if kind == 'w' else split16(t)
        if mode == 'x_bf16':
            return O.bf16_round(t) if kind == 'x' else split16(t)
        raise KeyError(mode)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--planes', type=int, default=16)
    ap.add_argument('--h', type=int, default=96)
    ap.add_argument('--w', type=int, default=192)
    ap.add_argument('--seeds', type=int, default=2)
    ap.add_argument('--json', default=None)
    args = ap.parse_args()
    torch.set_num_threads(os.cpu_count() or 8)
    res = {m: {l: 0.0 for l in LAYERS} for m in ('tf32', 'w_bf16', 'x_bf16')}
    base = {'split3_all': 0.0, 'tf32_all': 0.0}
    for seed in range(args.seeds):
        cur, prev, metas, params = syn.make_kitti_pair(300 + seed, args.h, args.w, args.planes)
        cfg = syn.depth_cfg_for(args.planes)

        def run(table):
            with torch.no_grad():
                return O.dfm_backbone_forward(params, cur, prev, metas, cfg, q=LayerQ(table))[0]
        ref = run({l: 'fp32' for l in LAYERS}).double()
        scale = float(ref.abs().max())

        def err(table):
            return float((run(table).double() - ref).abs().max()) / scale
        base['split3_all'] = max(base['split3_all'], err({}))
        base['tf32_all'] = max(base['tf32_all'], err({l: 'tf32' for l in LAYERS}))
        for mode in res:
            for l in LAYERS:
                res[mode][l] = max(res[mode][l], err({l: mode}))
    print(f'| layer | tf32 | w_bf16 | x_bf16 |  (normalised max error of the gated logits, '
          f'{args.h}x{args.w}, D={args.planes}, worst of {args.seeds} seeds)')
    print('|---|---|---|---|')
    for l in LAYERS:
        print(f'| {l} | {res["tf32"][l]:.1e} | {res["w_bf16"][l]:.1e} | {res["x_bf16"][l]:.1e} |')
    print(f'\nall layers 3-term split: {base["split3_all"]:.1e};  all layers tf32: '
          f'{base["tf32_all"]:.1e}')
    # greedy: cheapest layers first under a 5e-4 budget (verify the set jointly)
    order = sorted(LAYERS, key=lambda l: res['tf32'][l])
    cur, prev, metas, params = syn.make_kitti_pair(300, args.h, args.w, args.planes)
    cfg = syn.depth_cfg_for(args.planes)
    with torch.no_grad():
        ref = O.dfm_backbone_forward(params, cur, prev, metas, cfg,
                                     q=LayerQ({l: 'fp32' for l in LAYERS}))[0].double()
    chosen = []
    for l in order:
        trial = chosen + [l]
        with torch.no_grad():
            out = O.dfm_backbone_forward(params, cur, prev, metas, cfg,
                                         q=LayerQ({k: 'tf32' for k in trial}))[0].double()
        e = float((out - ref).abs().max() / ref.abs().max())
        if e <= 5e-4:
            chosen = trial
            last = e
    print(f'greedy tf32 set under 5e-4 ({len(chosen)} of {len(LAYERS)} layers, joint error '
          f'{last:.1e}): {chosen}')
    if args.json:
        json.dump(dict(per_layer=res, base=base, tf32_set=chosen), open(args.json, 'w'), indent=1)


if __name__ == '__main__':
    main()
