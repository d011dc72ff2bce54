"""Multi-GPU plumbing: one process per GPU, utterances sharded on dim 0, no data-path collective.

The reference's only multi-GPU mechanism is single-process nn.DataParallel (train.py:76-78: scatter the batch on dim 0,
replicate the module, gather).  Utterances are independent in this model, so the B200-native equivalent shards the
batch across ranks and reduces nothing but scalars ({loss numerator, token count, LER sum, utterance count}) with one
all-reduce -- NCCL on GPUs, gloo in the CPU tests.
"""
from __future__ import annotations

import torch
import torch.distributed as dist


def shard_bounds(n, rank, world):
    """Contiguous shard [lo, hi) of `n` utterances for `rank`, chunked like torch.chunk / DataParallel's scatter."""
    per = -(-n // world)
    lo = min(rank * per, n)
    return lo, min(lo + per, n)


def reduce_sums(values, device=None, group=None):
    """All-reduce (SUM) a short list of python floats; returns a list of floats (identity when not distributed)."""
    t = torch.tensor(list(values), dtype=torch.float64, device=device)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return t.tolist()


def sharded_eval(forward_fn, batch_data, batch_label_idx, max_label_len, ler_fn, device=None, group=None):
    """Evaluate one global batch with every rank decoding its own shard.

    forward_fn(x_shard) -> log-probs [B_shard, S, V] (torch tensor); batch_label_idx int [B, S_lab].
    Returns dict(loss=NLL(ignore_index=0) over the GLOBAL batch, ler=mean LER, n=utterances)."""
    world = dist.get_world_size(group) if (dist.is_available() and dist.is_initialized()) else 1
    rank = dist.get_rank(group) if world > 1 else 0
    lo, hi = shard_bounds(batch_data.size(0), rank, world)
    nll = cnt = ler_sum = 0.0
    if hi > lo:
        logp = forward_fn(batch_data[lo:hi])
        L = min(logp.size(1), batch_label_idx.size(1), max_label_len)
        lab = batch_label_idx[lo:hi, :L].to(logp.device).long()
        picked = torch.gather(logp[:, :L, :], 2, lab.unsqueeze(-1)).squeeze(-1)
        keep = lab != 0
        nll = float(-(picked * keep).sum())
        cnt = float(keep.sum())
        ler_sum = float(sum(ler_fn(logp[:, :L, :].argmax(-1).cpu().numpy(), lab.cpu().numpy())))
    tot = reduce_sums([nll, cnt, ler_sum, float(hi - lo)], device=device, group=group)
    return {"loss": tot[0] / max(tot[1], 1.0), "ler": tot[2] / max(tot[3], 1.0), "n": int(tot[3])}
