"""Evaluation-side mirror of the reference's solver/solver.py ("next" row f1 of SURVEY.md section 8).

`batch_iterator` keeps the reference's signature and return value for the forward/evaluation path
(solver/solver.py:48-101 with is_training=False).  The training branch (:94-97: backward, clip_grad_norm_, optimizer
step) is outside this build's scope -- the B200 path is forward only -- and raises.

The NLL numerator / denominator are reduced on the device by `las_nll_sums` (include/las_b200.h); the letter error rate
needs a host-side edit distance exactly like the reference (which calls the `editdistance` package on CPU lists).
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _cabi


def _levenshtein(a, b):
    prev = list(range(len(b) + 1))
    for i, ca in enumerate(a, 1):
        cur = [i]
        for j, cb in enumerate(b, 1):
            cur.append(min(prev[j] + 1, cur[j - 1] + 1, prev[j - 1] + (ca != cb)))
        prev = cur
    return prev[-1]


def LetterErrorRate(pred_y, true_y):
    """solver/solver.py:11-24: prediction cut at the first <eos>=1, zeros dropped; truth drops 0 and 1."""
    out = []
    for p, t in zip(pred_y, true_y):
        ct = [int(w) for w in t if (w != 1 and w != 0)]
        cp = []
        for w in p:
            if w == 0:
                continue
            if w == 1:
                break
            cp.append(int(w))
        out.append(_levenshtein(cp, ct) / len(ct))
    return out


def label_smoothing_loss(pred_y, true_y, label_smoothing=0.1):
    """solver/solver.py:33-45 on the device: pred_y log-probs [B,S,V] (or the decoder's own [S,B,V] buffer, see `layout`), true_y
    one-hot [B,S,V] whose padding rows are all zero.  Per labelled step the smoothed target ((1-ls) one-hot + ls/V) dotted with
    the log-probs is (1-ls) logp[label] + (ls/V) sum_v logp[v]; `las_label_smoothing_terms` reduces that per utterance and
    divides by the utterance's label count, the scalar is minus the batch mean."""
    assert pred_y.size() == true_y.size()
    if not pred_y.is_cuda:
        raise RuntimeError("label_smoothing_loss runs on the device (las_label_smoothing_terms); there is no CPU fallback")
    rows = true_y.sum(dim=-1)
    labels = torch.where(rows > 0, true_y.argmax(dim=-1), torch.full_like(rows, -1, dtype=torch.long)).to(torch.int32).contiguous()
    return label_smoothing_loss_from_indices(pred_y.permute(1, 0, 2).contiguous(), labels, label_smoothing, pred_y.size(1))


def label_smoothing_loss_from_indices(logp_sbv, labels, label_smoothing, max_label_len):
    """Same loss straight from the decoder's [S,B,V] log-prob buffer (Speller.last_logp) and int32 labels [B,S'] (-1 = padding row)."""
    lib = _cabi.load_library()
    S, B, V = logp_sbv.shape
    per_utt = torch.empty(B, dtype=torch.float32, device=logp_sbv.device)
    with torch.cuda.device(logp_sbv.device):
        _cabi.check(lib.las_label_smoothing_terms(_cabi.ptr(logp_sbv), _cabi.ptr(labels), S, labels.size(1), B, V, int(max_label_len),
                                                  float(label_smoothing), _cabi.ptr(per_utt), _cabi.current_stream_ptr(logp_sbv.device)))
    return -per_utt.mean()


def nll_sums(logp_sbv, label_idx, max_label_len):
    """Device reduction of NLLLoss(ignore_index=0): returns a [2] tensor (sum of -logp[label], number of kept labels).
    logp_sbv is the decoder's [S,B,V] buffer; label_idx int32 [B,S_lab]."""
    lib = _cabi.load_library()
    S, B, V = logp_sbv.shape
    out = torch.zeros(2, dtype=torch.float32, device=logp_sbv.device)
    with torch.cuda.device(logp_sbv.device):
        _cabi.check(lib.las_nll_sums(_cabi.ptr(logp_sbv), _cabi.ptr(label_idx), S, label_idx.size(1), B, V, int(max_label_len),
                                     _cabi.ptr(out), _cabi.current_stream_ptr(logp_sbv.device)))
    return out


def batch_iterator(batch_data, batch_label, las_model, optimizer, tf_rate, is_training, max_label_len, label_smoothing,
                   use_gpu=True, vocab_dict=None):
    """Forward + loss + LER for one batch; same return value as the reference: (loss ndarray, [LER per utterance]).
    `las_model` may be wrapped in nn.DataParallel as train.py:76-78 does: the wrapper's replicas cannot hand back attributes, so
    the loss is then reduced from the gathered log-probabilities (`las_nll_sums`) instead of the decoder's fused terms."""
    if is_training:
        raise NotImplementedError("the B200 path is forward-only; training (backward / optimizer step, solver/solver.py:94-97) is out of scope")
    max_label_len = min([batch_label.size()[1], max_label_len])
    true_idx = torch.max(batch_label, dim=2)[1][:, :max_label_len].contiguous()
    wrapped = hasattr(las_model, "module") and not hasattr(las_model, "speller")
    # is_training is False here, so the reference takes the NLLLoss(ignore_index=0) branch (solver.py:70-77); the decoder emits its
    # terms -logp[s,b,label] itself (fused epilogue), so the loss needs no pass over the [B,S,V] log-probabilities
    if wrapped:
        raw_pred_seq, _ = las_model(batch_data=batch_data, batch_label=batch_label, teacher_force_rate=tf_rate, is_training=is_training)
    else:
        raw_pred_seq, _ = las_model(batch_data=batch_data, batch_label=batch_label, teacher_force_rate=tf_rate, is_training=is_training,
                                    nll_labels=true_idx)
    if len(raw_pred_seq) < max_label_len:
        # the reference fails here too (NLLLoss gets [B,V,steps] against [B,max_label_len] targets, solver.py:68-72)
        raise RuntimeError(f"the decoder ran {len(raw_pred_seq)} steps but the loss covers {max_label_len} labels: "
                           "speller.max_label_len is smaller than the label length")
    pred_y = torch.stack(raw_pred_seq, dim=1)[:, :max_label_len, :].contiguous()  # [B,S,V], as solver.py:68 (for the LER below)
    if wrapped:
        sums = nll_sums(pred_y.permute(1, 0, 2).contiguous(), true_idx.to(device=pred_y.device, dtype=torch.int32).contiguous(), max_label_len)
        loss = sums[0] / sums[1]
    else:
        terms = las_model.speller.last_nll_terms[:max_label_len]
        loss = terms.sum() / (true_idx != 0).sum().to(terms.device)
    batch_ler = LetterErrorRate(torch.max(pred_y, dim=2)[1].cpu().numpy(), true_idx.cpu().numpy())
    return loss.cpu().numpy(), batch_ler
