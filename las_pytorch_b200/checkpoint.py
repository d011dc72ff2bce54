"""Checkpoint packages in the reference's format ("next" row f3 of SURVEY.md section 8).

The reference saves `LAS.serialize(...)` dicts with torch.save as `{save_folder}/{name}-epoch{N}.pth.tar`
(train.py:181-201) and resumes with `las.load_state_dict(package["state_dict"])` (train.py:83-90).  These helpers do the
same for our drop-in modules; a `module.` prefix (left by nn.DataParallel wrapping, train.py:76-78) is stripped.
"""
from __future__ import annotations

import torch

from .las_model import LAS, Listener, Speller


def strip_module_prefix(state_dict):
    return {(k[len("module."):] if k.startswith("module.") else k): v for k, v in state_dict.items()}


def _unit_name(etype):
    """`etype` in a reference package is the speller's rnn CLASS (model/las_model.py:54 overwrites the listener's string written at
    :49), e.g. torch.nn.LSTM; older / hand-made packages may hold the string."""
    name = etype if isinstance(etype, str) else getattr(etype, "__name__", "LSTM")
    return str(name).upper()


def load_package(package_or_path, las=None, map_location="cpu", precision=None, **speller_kwargs):
    """Loads a reference checkpoint package (`torch.save(las.serialize(...))`, train.py:181-201).  With `las=None` a model is built
    from the package's own hyper-parameters (einput/ehidden/elayer/etype/dvocab_size/dhidden/dlayer, model/las_model.py:42-63); the
    attention options the package does not record are read off the state_dict's shapes (phi -> mlp dim x heads, dim_reduce ->
    heads, no phi -> no MLP), the rest (`max_label_len`, `decode_mode`, `mlp_activate_in_attention`) come from `speller_kwargs`
    (defaults follow config/librispeech-config.yaml:27-34).  The package holds a pickled class (`etype`), so it is read with
    weights_only=False exactly as train.py:84 does: only load packages you trust."""
    pkg = package_or_path
    if not isinstance(pkg, dict):
        pkg = torch.load(package_or_path, map_location=map_location, weights_only=False)
    sd = strip_module_prefix(pkg["state_dict"])
    if las is None:
        unit = _unit_name(pkg.get("etype", "LSTM"))
        heads, use_mlp, mlp_dim = 1, "speller.attention.phi.weight" in sd, 0
        if "speller.attention.dim_reduce.weight" in sd:
            w = sd["speller.attention.dim_reduce.weight"]
            heads = w.shape[1] // w.shape[0]
        if use_mlp:
            mlp_dim = sd["speller.attention.psi.weight"].shape[0]
            assert sd["speller.attention.phi.weight"].shape[0] == mlp_dim * heads
        opts = dict(max_label_len=576, use_mlp_in_attention=use_mlp, mlp_dim_in_attention=mlp_dim, mlp_activate_in_attention="relu",
                    multi_head=heads, decode_mode=1)
        opts.update(speller_kwargs)
        extra = {} if precision is None else {"precision": precision}
        listener = Listener(input_feature_dim=pkg["einput"], hidden_size=pkg["ehidden"], num_layers=pkg["elayer"], rnn_unit=unit,
                            use_gpu=True, dropout_rate=pkg.get("edropout", 0.0), **extra)
        speller = Speller(vocab_size=pkg["dvocab_size"], hidden_size=pkg["dhidden"], rnn_unit=unit, num_layers=pkg["dlayer"],
                          listener_hidden_size=pkg["ehidden"], use_gpu=True, **opts, **extra)
        las = LAS(listener, speller)
    las.load_state_dict(sd, strict=True)
    return las, pkg


def save_package(las, path, optimizer=None, epoch=0, tr_loss=None, val_loss=None):
    torch.save(las.serialize(optimizer, epoch, tr_loss, val_loss), path)
