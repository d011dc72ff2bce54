"""Parameter containers with the reference's state_dict layout.

The reference holds its weights in torch `nn.LSTM` / `nn.Linear` modules
(/root/reference/model/las_model.py:72-79,164-166,174,266-269).  These containers register tensors with
exactly the same names, shapes, order and default initialisation (same RNG consumption), so that

  * `load_state_dict(package["state_dict"])` of a reference checkpoint works with strict=True, and
  * constructing a model under `torch.manual_seed(s)` yields bit-identical weights to the reference
    constructed under the same seed (checked by tests/golden/make_golden.py against the real reference).

They hold parameters only; there is no forward here (the arithmetic lives in csrc/).
"""
from __future__ import annotations

import math

import torch
import torch.nn as nn


class LSTMWeights(nn.Module):
    """Parameter layout of torch.nn.LSTM / nn.GRU / nn.RNN(input_size, hidden_size, num_layers, bidirectional).

    Names: weight_ih_l{k}[_reverse] [G*H,in], weight_hh_l{k}[_reverse] [G*H,H], bias_ih_l{k}[_reverse],
    bias_hh_l{k}[_reverse]; G = 4 (LSTM, gates i,f,g,o), 3 (GRU, gates r,z,n) or 1 (RNN, tanh); all
    U(-1/sqrt(H), 1/sqrt(H)) drawn in registration order, exactly as torch's RNNBase does.
    """

    GATES = {"LSTM": 4, "GRU": 3, "RNN": 1}

    def __init__(self, input_size, hidden_size, num_layers=1, bidirectional=False, cell="LSTM"):
        super().__init__()
        self.cell = cell
        gates = self.GATES[cell]
        self.input_size = input_size
        self.hidden_size = hidden_size
        self.num_layers = num_layers
        self.bidirectional = bidirectional
        dirs = 2 if bidirectional else 1
        for layer in range(num_layers):
            in_dim = input_size if layer == 0 else hidden_size * dirs
            for d in range(dirs):
                sfx = "_reverse" if d == 1 else ""
                self.register_parameter(f"weight_ih_l{layer}{sfx}", nn.Parameter(torch.empty(gates * hidden_size, in_dim)))
                self.register_parameter(f"weight_hh_l{layer}{sfx}", nn.Parameter(torch.empty(gates * hidden_size, hidden_size)))
                self.register_parameter(f"bias_ih_l{layer}{sfx}", nn.Parameter(torch.empty(gates * hidden_size)))
                self.register_parameter(f"bias_hh_l{layer}{sfx}", nn.Parameter(torch.empty(gates * hidden_size)))
        self.reset_parameters()

    def reset_parameters(self):
        stdv = 1.0 / math.sqrt(self.hidden_size) if self.hidden_size > 0 else 0
        for w in self.parameters():
            nn.init.uniform_(w, -stdv, stdv)

    def forward(self, *a, **k):  # pragma: no cover - containers are never called
        raise RuntimeError("LSTMWeights is a parameter container; the LSTM runs inside the CUDA extension")


class LinearWeights(nn.Module):
    """Parameter layout and default init of torch.nn.Linear(in_features, out_features)."""

    def __init__(self, in_features, out_features):
        super().__init__()
        self.in_features = in_features
        self.out_features = out_features
        self.weight = nn.Parameter(torch.empty(out_features, in_features))
        self.bias = nn.Parameter(torch.empty(out_features))
        self.reset_parameters()

    def reset_parameters(self):
        nn.init.kaiming_uniform_(self.weight, a=math.sqrt(5))
        bound = 1 / math.sqrt(self.in_features) if self.in_features > 0 else 0
        nn.init.uniform_(self.bias, -bound, bound)

    def forward(self, *a, **k):  # pragma: no cover
        raise RuntimeError("LinearWeights is a parameter container; the projection runs inside the CUDA extension")
