"""world_size-2 gloo test of the multi-GPU host logic (sharding + scalar reduction), run on CPU.

The forward itself needs a B200; here every rank's shard is decoded by the numpy oracle instead, which is exactly
what the sharding logic must be indifferent to."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import las_testlib as tl
from las_pytorch_b200 import parallel
from las_pytorch_b200.solver import LetterErrorRate
from oracle import las_oracle as O


def _oracle_forward(sd, cfg, labels, S):
    def fn(x_shard):
        out = O.las_forward(x_shard.numpy(), sd, cfg["L"], cfg["sl"], S, dtype=np.float64)
        return torch.from_numpy(out["logp"]).permute(1, 0, 2).contiguous()  # [B,S,V]
    return fn


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        g = np.load(os.path.join(tl.GOLDEN_DIR, "tiny_greedy_g3.npz"))
        cfg = tl.CONFIGS["tiny"]
        sd = {k[2:]: g[k] for k in g.files if k.startswith("w:")}
        x = torch.from_numpy(g["x"])
        labels = torch.from_numpy(g["labels"]).long()
        S = g["logp_f64"].shape[0]
        res = parallel.sharded_eval(_oracle_forward(sd, cfg, labels, S), x, labels, S, LetterErrorRate)
        q.put((rank, res))
    finally:
        dist.destroy_process_group()


def test_shard_bounds_cover_batch():
    for n in (1, 3, 64, 65, 512):
        for world in (1, 2, 3, 8):
            spans = [parallel.shard_bounds(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            assert max(hi - lo for lo, hi in spans) <= -(-n // world)


def test_two_rank_eval_equals_single_process():
    g = np.load(os.path.join(tl.GOLDEN_DIR, "tiny_greedy_g3.npz"))
    cfg = tl.CONFIGS["tiny"]
    sd = {k[2:]: g[k] for k in g.files if k.startswith("w:")}
    x = torch.from_numpy(g["x"])
    labels = torch.from_numpy(g["labels"]).long()
    S = g["logp_f64"].shape[0]
    single = parallel.sharded_eval(_oracle_forward(sd, cfg, labels, S), x, labels, S, LetterErrorRate)
    # reference value straight from the frozen reference output
    logp_ref = torch.from_numpy(g["logp_f64"]).permute(1, 0, 2)
    assert abs(single["loss"] - O.nll_loss_ignore0(logp_ref.numpy(), labels.numpy())) < 1e-9

    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    results = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for _, res in results:
        assert res["n"] == x.size(0)
        assert abs(res["loss"] - single["loss"]) < 1e-9
        assert abs(res["ler"] - single["ler"]) < 1e-12
