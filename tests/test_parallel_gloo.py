"""World-size-2 gloo test (CPU) of the N>1 host logic: frame sharding + the single waypoint gather."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        return s.getsockname()[1]


def _worker(rank, world, port, total, q):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    from thinktwice_b200.parallel import gather_waypoints, shard_batch, shard_range
    batch = {'img': torch.arange(total, dtype=torch.float32).view(total, 1), 'speed': torch.arange(total) * 2.0,
             'img_metas': [[{'id': i}] for i in range(total)], 'scalar': 7}
    local = shard_batch(batch, rank, world)
    lo, hi = shard_range(total, rank, world)
    assert [m[0]['id'] for m in local['img_metas']] == list(range(lo, hi)) and local['scalar'] == 7
    # the "forward": waypoints that encode the global frame id
    wp = local['img'].view(-1, 1, 1, 1).repeat(1, 6, 4, 2)
    allwp = gather_waypoints(wp, world)
    q.put((rank, allwp[:, 0, 0, 0].tolist()))
    dist.destroy_process_group()


def _run(total):
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = _free_port()
    ps = [ctx.Process(target=_worker, args=(r, 2, port, total, q)) for r in range(2)]
    for p in ps:
        p.start()
    res = [q.get(timeout=120) for _ in ps]
    for p in ps:
        p.join(timeout=60)
        assert p.exitcode == 0
    for _, ids in res:
        assert ids == [float(i) for i in range(total)]          # every rank sees all frames, in frame order


def test_even_split_gather():
    _run(4)


def test_ragged_split_gather():
    _run(5)
