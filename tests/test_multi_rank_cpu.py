"""CPU test of the N>1 host logic (gloo, world_size 2): the pairs are sharded over
ranks with no data-path collective; only the timing reduction (MAX over ranks) and the
rank-0 report use torch.distributed."""
import os
import socket
import subprocess
import sys
import textwrap

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = textwrap.dedent('''
    import os, sys, json
    sys.path.insert(0, %r)
    import torch, torch.distributed as dist
    from depth_from_motion_b200 import synthetic as syn
    from depth_from_motion_b200.sharding import shard_pairs, reduce_step_time
    dist.init_process_group('gloo')
    rank, world = dist.get_rank(), dist.get_world_size()
    mine = shard_pairs(list(range(8)), rank, world)
    # every pair is owned by exactly one rank
    owned = [None] * world
    dist.all_gather_object(owned, mine)
    flat = sorted(sum(owned, []))
    assert flat == list(range(8)), flat
    # rank-local synthetic inputs differ between ranks (different seeds)
    cur, prev, metas, params = syn.make_kitti_pair(100 + rank, 32, 64, 8)
    sig = float(cur.sum())
    sigs = [None] * world
    dist.all_gather_object(sigs, sig)
    assert len(set(sigs)) == world
    # step time is the max over ranks
    t = reduce_step_time(10.0 + rank, 'cpu')
    assert abs(t - (10.0 + world - 1)) < 1e-6
    if rank == 0:
        print(json.dumps({'ok': True, 'world': world}))
    dist.destroy_process_group()
''') % ROOT


def test_two_rank_sharding_gloo(tmp_path):
    script = tmp_path / 'worker.py'
    script.write_text(WORKER)
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    port = s.getsockname()[1]
    s.close()
    cmd = [sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node=2',
           '--master-addr', '127.0.0.1', '--master-port', str(port), str(script)]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
    assert '"ok": true' in r.stdout


def test_shard_pairs_edge_cases():
    """Ragged and empty partitions: every pair has exactly one owner, order is kept, ranks
    beyond the number of pairs get nothing."""
    from depth_from_motion_b200.sharding import reduce_step_time, shard_pairs
    for n, world in ((0, 2), (1, 4), (5, 2), (8, 8), (9, 4)):
        pairs = list(range(n))
        parts = [shard_pairs(pairs, r, world) for r in range(world)]
        assert sorted(sum(parts, [])) == pairs
        assert all(p == sorted(p) for p in parts)
        assert max(len(p) for p in parts) - min(len(p) for p in parts) <= 1
    # without an initialised process group the reduction is the identity
    assert reduce_step_time(3.5, 'cpu') == 3.5
