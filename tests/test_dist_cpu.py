"""world_size-2 gloo test of the one collective on the path (dist.all_merge)."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import rnf_oracle as orc
from rotationnormflow_b200 import dist as rdist


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    try:
        _worker_body(rank, world, port, q)
    except Exception as e:  # surface failures instead of letting the parent wait for its timeout
        q.put((rank, repr(e)))


def _worker_body(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    gen = torch.Generator().manual_seed(7)
    B, G = 6, 1001
    logp = torch.randn(B, G, generator=gen) * 4
    logp[0, 3] = logp[0, 700] = 40.0
    b, e = rdist.shard_range(G, rank, world)
    part = logp[:, b:e]
    m = part.max(1).values
    out = rdist.all_merge(m, part.argmax(1) + b, torch.exp(part - m[:, None]).sum(1))
    ridx, rmx, rlme = orc.grid_reduce(logp)
    ok = torch.equal(out[1], ridx) and torch.equal(out[0], rmx) and \
        float((rdist.log_normaliser(out[0], out[2], G) - rlme).abs().max()) < 1e-5
    # the same collective with the spread numerators riding along ([B,4] instead of [B,3])
    samples = orc.random_rotations(G, gen)
    gt = orc.random_rotations(B * 2, gen).reshape(B, 2, 3, 3)
    prod = torch.einsum("gij,bkij->bgk", samples.double(), gt.double())
    d = torch.acos(torch.clip((prod.max(-1).values - 1.0) / 2.0, -1.0, 1.0))
    e_part = torch.exp(part.double() - m[:, None].double())
    out4 = rdist.all_merge(m, part.argmax(1) + b, e_part.sum(1).float(), sn=(e_part * d[:, b:e]).sum(1).float())
    want = orc.spread(logp, samples, gt)
    ok = ok and torch.equal(out4[1], ridx) and float(((out4[3].double() / out4[2].double()) - want).abs().max()) < 1e-5
    q.put((rank, bool(ok)))
    dist.destroy_process_group()


def test_all_merge_world2_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert sorted(res) == [(0, True), (1, True)]
