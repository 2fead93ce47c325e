"""N > 1 host logic on CPU (gloo, world_size 2): the path does not shard, so ranks are independent replicas; what is
shared is the timing rule (max over ranks of the device-timed duration, one all-reduce) and the reference arm's
"rank 0 alone runs and prints" contract.  No GPU, no compute through the CUDA library."""
import json
import os
import socket
import subprocess
import sys

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), '..'))


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    import bench
    dist.init_process_group('gloo', init_method='tcp://127.0.0.1:%d' % port, rank=rank, world_size=world)
    # rank r "measured" (10 + r, 20 - r) ms: the job's value must use the slowest rank of each region
    out = bench.reduce_times([10.0 + rank, 20.0 - rank], world)
    dist.barrier()
    dist.destroy_process_group()
    q.put((rank, out))


def test_times_are_max_over_ranks_gloo():
    world, port = 2, _free_port()
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=300) for _ in range(world))
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    assert res[0] == res[1] == [11.0, 20.0]


def test_single_rank_needs_no_process_group():
    sys.path.insert(0, ROOT)
    import bench
    assert bench.reduce_times([3.0, 4.0], 1) == [3.0, 4.0]


def test_reference_arm_under_torchrun_rank0_only():
    """`bench.py --impl reference` launched the way the driver launches N > 1: rank 0 prints ONE json line, the other rank
    exits 0 without work (tiny plumbing workload so the oracle port finishes in seconds)."""
    cmd = [sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node', '2', '--master-addr', '127.0.0.1',
           '--master-port', str(_free_port()), os.path.join(ROOT, 'bench.py'), '--impl', 'reference', '--gpus', '2',
           '--steps', '1', '--warmup', '0', '--workload', 'tiny']
    r = subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.startswith('{')]
    assert len(lines) == 1, r.stdout
    d = json.loads(lines[0])
    assert d['impl'] == 'reference' and d['n_gpus'] == 2 and d['unit'] == 'frames/s' and d['higher_is_better'] is True
    assert d['value'] > 0 and d['cpu_baseline']['kind'] in ('reference', 'port') and d['cpu_baseline']['value'] == d['value']
    assert d['e2e'] == {'value': d['value'], 'unit': 'frames/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}


def _flat_worker(rank, world, port, q):
    import torch
    import torch.distributed as dist
    import torch.nn as nn
    from vid2vid_b200.trainer import FlatGrads
    dist.init_process_group('gloo', init_method='tcp://127.0.0.1:%d' % port, rank=rank, world_size=world)
    torch.manual_seed(0)
    g = nn.Sequential(nn.Conv2d(3, 4, 3), nn.BatchNorm2d(4))
    d = nn.Sequential(nn.Conv2d(4, 2, 3))
    fg = FlatGrads([list(g.parameters()), list(d.parameters())])
    x = torch.randn(2, 3, 8, 8, generator=torch.Generator().manual_seed(10 + rank))       # this rank's shard of the batch
    loss = d(g(x)).pow(2).mean()
    fg.zero()
    loss.backward()
    local = fg.flat.clone()
    fg.all_reduce_mean(world)
    # (plain lists, not tensors: a tensor travels through the queue as a shared-memory handle that dies with this process)
    q.put((rank, local.tolist(), fg.flat.clone().tolist(), [p.grad.data_ptr() == fg.flat[o:].data_ptr() for p, o in
                                          zip(list(g.parameters()) + list(d.parameters()), _offsets(list(g.parameters()) + list(d.parameters())))]))
    dist.destroy_process_group()


def _offsets(params):
    off, out = 0, []
    for p in params:
        out.append(off)
        off += p.numel()
    return out


def test_flat_gradient_all_reduce_world2_gloo():
    """one flat [G | D | ...] gradient buffer, one all-reduce(SUM)/world per step: the result on every rank is the
    mean of the per-rank gradients (= the gradient of the mean of the per-replica losses, train.py:65,78), and every
    parameter's .grad stays a view into the flat buffer."""
    import torch.multiprocessing as mp
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_flat_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=300) for _ in procs], key=lambda t: t[0])
    for p in procs:
        p.join(timeout=120)
    import torch
    (r0, l0, a0, v0), (r1, l1, a1, v1) = [(r, torch.tensor(l), torch.tensor(a), v) for r, l, a, v in res]
    assert torch.allclose(a0, (l0 + l1) / 2, atol=1e-7) and torch.equal(a0, a1)
    assert all(v0) and all(v1)
    assert not torch.allclose(l0, l1)
