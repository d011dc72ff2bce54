"""torch-CPU restatement of the reference's op sequence (TEST INFRASTRUCTURE -- see oracle/__init__.py).

Purpose: the CPU baseline that `bench.py` times on the GPU box's host cores (`cpu_baseline.kind == "port"`,
and `--impl reference`).  /root/reference cannot travel to the GPU box and its sources must not be copied,
so this file re-issues, in the same order, the same torch library calls the reference makes on its CPU path
-- including the work the reference wastes:

  * one bidirectional `nn.LSTM(batch_first=True)` call per pyramid layer        model/las_model.py:72-79,90
  * a seq-len-1 multi-layer `nn.LSTM` call per decode step                      model/las_model.py:164-166,179
  * psi(enc) recomputed every step through a flatten/linear/unflatten           model/las_model.py:279, utils/functions.py:72-77
  * bmm energy, softmax over all U, context via a materialised repeat()*enc     model/las_model.py:289-297
  * cat + Linear + LogSoftmax                                                   model/las_model.py:181-182
  * greedy feedback through topk(1) and a per-sample Python loop with int()     model/las_model.py:223-227
  * <sos> one-hot built with LongTensor.scatter_ then cast                      utils/functions.py:54-63

Because the same torch kernels run in the same order, its fp32 outputs are expected to be bit-identical to the
reference's on the same machine; tests/test_oracle_golden.py pins it against tests/golden/*.npz.
"""
from __future__ import annotations

import torch
import torch.nn as nn
import torch.nn.functional as F


class RefTorchLAS:
    """Holds torch modules shaped like the reference's and replays its forward op sequence on CPU."""

    def __init__(self, sd, listener_layers, speller_layers, dtype=torch.float32, device="cpu"):
        """device="cuda" replays the same op sequence the reference runs with use_gpu=True (torch -> cuDNN RNN / cuBLAS; the
        <sos> one-hot is still built on the host and copied, utils/functions.py:60-63; the greedy loop still does one int(i) host
        sync per sample and step, model/las_model.py:225-226): bench.py's `gpu_baseline`, SURVEY.md 2.1's "kernel to beat on the
        same box"."""
        sd = {k: torch.as_tensor(v).to(dtype) for k, v in sd.items()}
        self.dtype = dtype
        self.device = torch.device(device)
        self.blstm = []
        for l in range(listener_layers):
            pre = f"listener.pLSTM_layer{l}.BLSTM."
            w_ih = sd[pre + "weight_ih_l0"]
            m = nn.LSTM(w_ih.shape[1], w_ih.shape[0] // 4, 1, bidirectional=True, batch_first=True).to(dtype)
            m.load_state_dict({k[len(pre):]: v for k, v in sd.items() if k.startswith(pre)})
            self.blstm.append(m.eval().to(self.device))
        pre = "speller.rnn_layer."
        w_ih = sd[pre + "weight_ih_l0"]
        self.hs = w_ih.shape[0] // 4
        self.rnn = nn.LSTM(w_ih.shape[1], self.hs, num_layers=speller_layers, batch_first=True).to(dtype)
        self.rnn.load_state_dict({k[len(pre):]: v for k, v in sd.items() if k.startswith(pre)})
        self.rnn.eval().to(self.device)

        def lin(name):
            w, b = sd[name + ".weight"], sd[name + ".bias"]
            m = nn.Linear(w.shape[1], w.shape[0]).to(dtype)
            m.load_state_dict({"weight": w, "bias": b})
            return m.eval().to(self.device)

        self.phi = lin("speller.attention.phi")
        self.psi = lin("speller.attention.psi")
        self.cd = lin("speller.character_distribution")
        self.vocab = self.cd.out_features

    @torch.no_grad()
    def listener(self, x):
        out = x.to(self.dtype).to(self.device)
        for m in self.blstm:
            b, t, f = out.shape
            out, _ = m(out.contiguous().view(b, int(t / 2), f * 2))
        return out

    def _attend(self, state, enc):
        q = F.relu(self.phi(state))  # [B,1,D]
        b, u, e = enc.shape
        k = F.relu(self.psi(enc.contiguous().view(-1, e)).view(b, u, -1))  # recomputed each step
        energy = torch.bmm(q, k.transpose(1, 2)).squeeze(dim=1)
        score = F.softmax(energy, dim=-1)
        ctx = torch.sum(enc * score.unsqueeze(2).repeat(1, 1, e), dim=1)
        return score, ctx

    @torch.no_grad()
    def speller(self, enc, steps, ground_truth=None, decode_mode=1):
        """ground_truth: one-hot int64 [B,S,V] (teacher forcing) or None (free running)."""
        b = enc.size(0)
        idx = torch.zeros(b, 1).unsqueeze(2).type(torch.LongTensor)
        word = torch.LongTensor(b, 1, self.vocab).zero_().scatter_(-1, idx, 1).to(self.dtype).to(self.device)
        rnn_in = torch.cat([word, enc[:, 0:1, :]], dim=-1)
        hidden = None
        logps, attns = [], []
        for step in range(steps):
            rnn_out, hidden = self.rnn(rnn_in, hidden)
            score, ctx = self._attend(rnn_out, enc)
            logp = F.log_softmax(self.cd(torch.cat([rnn_out.squeeze(dim=1), ctx], dim=-1)), dim=-1)
            logps.append(logp)
            attns.append(score)
            if ground_truth is not None:
                word = ground_truth[:, step:step + 1, :].to(self.dtype)
            elif decode_mode == 0:
                word = logp.unsqueeze(1)
            else:
                word = torch.zeros_like(logp)
                for row, i in enumerate(logp.topk(1)[1]):
                    word[row, int(i)] = 1
                word = word.unsqueeze(1)
            rnn_in = torch.cat([word, ctx.unsqueeze(1)], dim=-1)
        return torch.stack(logps), torch.stack(attns)

    @torch.no_grad()
    def forward(self, x, steps, ground_truth=None, decode_mode=1):
        enc = self.listener(x)
        logp, attn = self.speller(enc, steps, ground_truth, decode_mode)
        return enc, logp, attn
