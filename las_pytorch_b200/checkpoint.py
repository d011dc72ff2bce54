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


def load_package(package_or_path, las=None, map_location="cpu", **speller_kwargs):
    """Loads a reference checkpoint package.  With `las=None` a model is built from the package's own hyper-parameters
    (einput/ehidden/elayer/dvocab_size/dhidden/dlayer, model/las_model.py:42-63); attention / decoding options that the
    package does not record come from `speller_kwargs` (defaults follow config/librispeech-config.yaml:27-34)."""
    pkg = package_or_path
    if not isinstance(pkg, dict):
        pkg = torch.load(package_or_path, map_location=map_location, weights_only=False)
    sd = strip_module_prefix(pkg["state_dict"])
    if las is None:
        opts = dict(max_label_len=576, use_mlp_in_attention=True, mlp_dim_in_attention=sd["speller.attention.phi.weight"].shape[0],
                    mlp_activate_in_attention="relu", multi_head=1, decode_mode=1)
        opts.update(speller_kwargs)
        listener = Listener(input_feature_dim=pkg["einput"], hidden_size=pkg["ehidden"], num_layers=pkg["elayer"], rnn_unit="LSTM",
                            use_gpu=True, dropout_rate=pkg.get("edropout", 0.0))
        speller = Speller(vocab_size=pkg["dvocab_size"], hidden_size=pkg["dhidden"], rnn_unit="LSTM", num_layers=pkg["dlayer"],
                          listener_hidden_size=pkg["ehidden"], use_gpu=True, **opts)
        las = LAS(listener, speller)
    las.load_state_dict(sd, strict=True)
    return las, pkg


def save_package(las, path, optimizer=None, epoch=0, tr_loss=None, val_loss=None):
    torch.save(las.serialize(optimizer, epoch, tr_loss, val_loss), path)
