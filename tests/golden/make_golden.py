"""Generates tests/golden/*.npz by running the UNMODIFIED reference (/root/reference) on CPU.

Run in the build container only (the GPU box has no /root/reference):

    python tests/golden/make_golden.py

The reference is imported as-is; only three data-prep-only third-party modules that `utils/functions.py`
imports at module top level are stubbed -- none is reachable from any `forward`.
Each case stores inputs, (for the tiny configs) the full state_dict, and the reference's fp32 and fp64 outputs.
For the larger configs the weights are reproduced from the seed by `las_pytorch_b200`'s own parameter
containers; this script asserts that those are bit-identical to the reference's freshly constructed weights and
stores a fingerprint.
"""
from __future__ import annotations

import copy
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, ROOT)

REF_ROOT = os.environ.get("LAS_REFERENCE_ROOT", "/root/reference")


def import_reference():
    for name in ("pydub", "editdistance", "python_speech_features"):
        sys.modules.setdefault(name, types.ModuleType(name))
    sys.modules["pydub"].AudioSegment = object
    sys.modules["python_speech_features"].logfbank = None
    sys.path.insert(0, REF_ROOT)
    import model.las_model as ref  # noqa

    return ref


def run_reference(las, x, gt_onehot, teacher_forced, dtype):
    """Runs the reference LAS on CPU.  Returns enc, logp [S,B,V], attn [S,B,U]."""
    m = copy.deepcopy(las).to(dtype).eval()
    m.speller.float_type = torch.DoubleTensor if dtype == torch.float64 else torch.FloatTensor
    with torch.no_grad():
        enc = m.listener(x.to(dtype))
        if teacher_forced:
            np.random.seed(0)
            preds, attns = m.speller(enc, ground_truth=gt_onehot, teacher_force_rate=1.1)  # always teacher-forced
        else:
            preds, attns = m.speller(enc, ground_truth=None, teacher_force_rate=0)
    logp = torch.stack(preds)
    if len(attns[0]) == 1:
        attn = torch.stack([a[0] for a in attns])                 # [S,B,U]
    else:
        attn = torch.stack([torch.stack(list(a)) for a in attns])  # multi_head > 1: [S,heads,B,U]
    return enc.numpy(), logp.numpy(), attn.numpy()


def main():
    import las_testlib as tl
    import las_pytorch_b200 as ours

    ref = import_reference()
    cases = [
        # name, cfg, B, T, S, mode ("tf" | "greedy" | "raw"), gain, store_weights
        ("tiny_tf_g3", "tiny", 3, 32, 6, "tf", 3.0, True),
        ("tiny_greedy_g3", "tiny", 3, 32, 10, "greedy", 3.0, True),
        ("tiny_raw_g3", "tiny", 3, 32, 7, "raw", 3.0, True),
        ("tiny_tf_g1", "tiny", 2, 16, 5, "tf", 1.0, True),
        ("odd_tf_g3", "odd", 2, 48, 9, "tf", 3.0, True),
        ("odd_greedy_g3", "odd", 2, 48, 11, "greedy", 3.0, True),
        ("small_tf_g1", "small", 4, 400, 50, "tf", 1.0, False),   # BASELINE.json configs[0]
        ("small_tf_g3", "small", 4, 400, 50, "tf", 3.0, False),
        ("small_greedy_g3", "small", 4, 400, 40, "greedy", 3.0, False),
        ("small_tf_g6", "small", 2, 160, 30, "tf", 6.0, False),
        ("paper_tf_g3", "paper", 2, 160, 24, "tf", 3.0, False),
        ("paper_greedy_g3", "paper", 2, 160, 24, "greedy", 3.0, False),
        # Speller / Attention variants
        ("tinymh_tf_g3", "tiny_mh", 3, 32, 6, "tf", 3.0, True),
        ("tinymh_greedy_g3", "tiny_mh", 3, 32, 10, "greedy", 3.0, True),
        ("tinynomlp_tf_g3", "tiny_nomlp", 3, 32, 6, "tf", 3.0, True),
        ("tinynomlp_greedy_g3", "tiny_nomlp", 3, 32, 10, "greedy", 3.0, True),
        ("tinygru_tf_g3", "tiny_gru", 3, 32, 6, "tf", 3.0, True),
        ("tinygru_greedy_g3", "tiny_gru", 3, 32, 10, "greedy", 3.0, True),
        ("tinyrnn_tf_g3", "tiny_rnn", 3, 32, 6, "tf", 3.0, True),
        ("tinyrnn_greedy_g3", "tiny_rnn", 3, 32, 10, "greedy", 3.0, True),
    ]
    only = [a for a in sys.argv[1:] if not a.startswith("-")]
    if only:  # regenerate just the named cases (the others stay byte-identical in git)
        cases = [c for c in cases if c[0] in only]
    for name, cfg, B, T, S, mode, gain, store_w in cases:
        c = tl.CONFIGS[cfg]
        dm = 0 if mode == "raw" else 1
        ref_las = tl.build_model(cfg, max_label_len=S, decode_mode=dm, seed=17, gain=gain, module=ref)
        our_las = tl.build_model(cfg, max_label_len=S, decode_mode=dm, seed=17, gain=gain, module=ours)
        sd_ref, sd_our = tl.state_dict_numpy(ref_las), tl.state_dict_numpy(our_las)
        assert list(sd_ref) == list(sd_our), "state_dict key order differs from the reference"
        for k in sd_ref:
            assert sd_ref[k].shape == sd_our[k].shape and np.array_equal(sd_ref[k], sd_our[k]), f"seeded init differs at {k}"
        our_las.load_state_dict(ref_las.state_dict(), strict=True)  # the drop-in contract

        x, labels = tl.make_inputs(B, T, c["F"], S, c["V"], seed=17)
        gt = tl.onehot(labels, c["V"])
        out = {"x": x.numpy(), "labels": labels.numpy().astype(np.int32), "gain": gain, "seed": 17,
               "cfg": cfg, "mode": mode, "fingerprint": tl.weights_fingerprint(sd_ref)}
        for dt, tag in ((torch.float32, "f32"), (torch.float64, "f64")):
            enc, logp, attn = run_reference(ref_las, x, gt, mode == "tf", dt)
            out[f"enc_{tag}"], out[f"logp_{tag}"], out[f"attn_{tag}"] = enc, logp, attn
        if store_w:
            for k, v in sd_ref.items():
                out["w:" + k] = v
        path = os.path.join(HERE, name + ".npz")
        np.savez_compressed(path, **out)
        d32 = np.abs(out["logp_f32"] - out["logp_f64"]).max()
        agree = (out["logp_f32"].argmax(-1) == out["logp_f64"].argmax(-1)).mean()
        print(f"{name:18s} enc{out['enc_f32'].shape} logp{out['logp_f32'].shape} fp32-vs-fp64 logp {d32:.2e} "
              f"argmax agree {agree:.3f} distinct tokens {len(np.unique(out['logp_f64'].argmax(-1)))} "
              f"-> {os.path.getsize(path) / 1024:.0f} KiB")


if __name__ == "__main__":
    main()
