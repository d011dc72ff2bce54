"""Parity AT THE BENCHMARKED SHAPES (BASELINE.json configs 2, 3, 4 = bench.py --workload c2 / c3 / c4).

The CUDA path (through the C ABI, same seed-17 gain-3 weights and seed-17 inputs as bench.py) against the reference's own
torch op sequence on the host cores (oracle/las_ref_torch.py, pinned to the reference-generated goldens by
tests/test_oracle_golden.py; model/las_model.py:81-91,178-238,275-297).  Per workload and arithmetic mode:

  * listener output max-abs                                     <= 2e-5 (fp32) / 3e-2 (bf16)
  * teacher-forced log-probs max-abs over all S steps           <= 1e-4 (fp32) / 2e-2 (bf16)      (north_star tolerances)
  * teacher-forced argmax equal wherever the oracle's top-2 margin exceeds 2 x the tolerance; the kept fraction is reported
    and must cover most positions (the mask is margin-proportional, not a fixed 0.2)
  * free-running greedy token agreement with the reference      >= 0.99 in fp32 mode (measured: 1.000 at all three shapes).
    bf16 mode: 300-600 free-running steps at gain-3 weights are chaotic -- one near-tie flip (1 % of positions have a top-2
    margin below the 4e-3 log-prob error bf16 operands cause) and the utterance follows another trajectory.  That is a property
    of bf16 GEMM operands, not of these kernels: the REFERENCE's own op sequence with nothing but its 2-D weights rounded to bf16
    (fp32 arithmetic everywhere) agrees with itself on 0.86 of the characters at c3.  The test therefore computes that number on
    the spot (`reference_bf16_weights_agreement`) and requires ours to be no worse than it minus 0.05 (1.5/B at small batches), and >= 0.99 wherever the
    rounded reference reaches it; every step of the trajectory we DO follow is held to 2e-2 by the re-scoring check below.
  * re-scoring: the reference, teacher-forced on OUR greedy tokens, reproduces our greedy log-probs within the tolerance
    (checks every step of the trajectory we actually followed, including after a near-tie flip)

Measured values go to $LAS_PARITY_REPORT (one JSON line per case) when set; tools/parity_shapes_report.py formats them.
"""
import json
import os

import numpy as np
import pytest
import torch

import las_testlib as tl

pytestmark = pytest.mark.gpu

# name -> (config, B, T, S): exactly bench.py's WORKLOADS
SHAPES = {"c2": ("small", 32, 1600, 300), "c3": ("paper", 64, 1600, 300), "c4": ("paper", 16, 3000, 600)}
TOL = {"fp32": dict(enc=2e-5, logp=1e-4, greedy=0.99), "bf16": dict(enc=3e-2, logp=2e-2, greedy=0.99),
       "fp16": dict(enc=4e-3, logp=3e-3, greedy=0.99)}  # LAS_MODE_F16: the same kernels with IEEE fp16 operands
_REF = {}


def _precisions():
    from las_pytorch_b200 import _cabi

    lib = _cabi.load_library()
    return ["fp32"] + (["bf16"] if lib.las_mode_available(_cabi.MODE_BF16) else []) + (["fp16"] if lib.las_mode_available(_cabi.MODE_F16) else [])


def reference_run(wl):
    """The reference's op sequence on the host cores, once per workload: listener, teacher-forced pass, greedy pass."""
    if wl in _REF:
        return _REF[wl]
    from oracle.las_ref_torch import RefTorchLAS

    cfg, B, T, S = SHAPES[wl]
    c = tl.CONFIGS[cfg]
    torch.set_num_threads(os.cpu_count() or 1)
    las = tl.build_model(cfg, max_label_len=S, seed=17, gain=3.0)
    sd = tl.state_dict_numpy(las)
    x, labels = tl.make_inputs(B, T, c["F"], S, c["V"], seed=17)
    m = RefTorchLAS(sd, c["L"], c["sl"])
    enc = m.listener(x)
    logp_tf, _ = m.speller(enc, S, tl.onehot(labels, c["V"]), 1)
    logp_gr, attn_gr = m.speller(enc, S, None, 1)
    # calibration of the bf16 / fp16 greedy gates: the same op sequence with only the 2-D weights rounded to the operand format
    cal = {}
    for name, rdt in (("bf16", torch.bfloat16), ("fp16", torch.float16)):
        sd_b = {k: (torch.from_numpy(v).to(rdt).float().numpy() if v.ndim == 2 else v) for k, v in sd.items()}
        mb = RefTorchLAS(sd_b, c["L"], c["sl"])
        logp_b, _ = mb.speller(mb.listener(x), S, None, 1)
        cal[name] = float((logp_b.argmax(-1) == logp_gr.argmax(-1)).float().mean())
    _REF[wl] = dict(cal=cal, model=m, sd=las.state_dict(), x=x, labels=labels, enc=enc, logp_tf=logp_tf.numpy(), logp_gr=logp_gr.numpy(),
                    attn_gr=attn_gr.numpy(), c=c)
    return _REF[wl]


@pytest.mark.parametrize("precision", _precisions())
@pytest.mark.parametrize("wl", sorted(SHAPES))
def test_parity_at_benchmark_shape(wl, precision):
    cfg, B, T, S = SHAPES[wl]
    r = reference_run(wl)
    c, tol = r["c"], TOL[precision]
    las = tl.build_model(cfg, max_label_len=S, seed=17, gain=3.0, precision=precision)
    las.load_state_dict(r["sd"], strict=True)
    las = las.cuda()
    x = r["x"].cuda()

    enc = las.listener(x)
    enc_err = float((enc.cpu() - r["enc"]).abs().max())

    np.random.seed(0)
    preds, _ = las(x, r["labels"].cuda(), 1.1, is_training=True)  # teacher forced on the label indices
    logp_tf = torch.stack(preds).cpu().numpy()
    tf_err = float(np.abs(logp_tf - r["logp_tf"]).max())
    srt = np.sort(r["logp_tf"], axis=-1)
    margin = srt[..., -1] - srt[..., -2]
    safe = margin > 2 * tol["logp"]
    tf_argmax_all = float((logp_tf.argmax(-1) == r["logp_tf"].argmax(-1)).mean())
    tf_argmax_safe_ok = bool(np.array_equal(logp_tf.argmax(-1)[safe], r["logp_tf"].argmax(-1)[safe]))

    preds, attns = las(x, None, 0.0, is_training=False)            # free-running greedy
    logp_gr = torch.stack(preds).cpu().numpy()
    tok = las.speller.last_tokens.cpu().numpy()
    assert np.array_equal(tok, logp_gr.argmax(-1))                 # the token stream is the argmax of the returned log-probs
    ref_tok = r["logp_gr"].argmax(-1)
    agree = float((tok == ref_tok).mean())
    per_utt = (tok == ref_tok).mean(0)
    # first step at which each utterance's trajectory leaves the reference's (S = never)
    first_div = np.where((tok != ref_tok).any(0), (tok != ref_tok).argmax(0), S)
    # the reference re-scores the trajectory we followed
    rescored, _ = r["model"].speller(r["enc"], S, tl.onehot(torch.from_numpy(tok.T.astype(np.int64)), c["V"]), 1)
    rescore_err = float(np.abs(logp_gr - rescored.numpy()).max())
    attn = torch.stack([a[0] for a in attns]).cpu().numpy()

    rep = dict(workload=wl, precision=precision, B=B, T=T, S=S, listener_max_abs=enc_err, tf_logp_max_abs=tf_err,
               tf_argmax_agreement=tf_argmax_all, tf_argmax_checked_fraction=float(safe.mean()), greedy_token_agreement=agree,
               greedy_utterances_identical=int((per_utt == 1.0).sum()), greedy_first_divergence_median=float(np.median(first_div)),
               greedy_rescored_logp_max_abs=rescore_err, distinct_tokens=int(len(np.unique(ref_tok))),
               reference_bf16_weights_agreement=r["cal"].get(precision, r["cal"]["bf16"]))
    print("PARITY " + json.dumps(rep))
    if os.environ.get("LAS_PARITY_REPORT"):
        with open(os.environ["LAS_PARITY_REPORT"], "a") as f:
            f.write(json.dumps(rep) + "\n")

    assert enc_err <= tol["enc"], rep
    assert tf_err <= tol["logp"], rep
    assert safe.mean() >= {"fp32": 0.9, "fp16": 0.8, "bf16": 0.4}[precision] and tf_argmax_safe_ok, rep
    assert rescore_err <= tol["logp"], rep
    assert np.abs(np.exp(logp_gr).sum(-1) - 1).max() < 1e-4 and np.abs(attn.sum(-1) - 1).max() < 1e-4
    # (slack: 0.05, or 1.5 utterances' worth at small batches -- one utterance that leaves the reference's trajectory early moves the
    # agreement by up to 1/B)
    floor = tol["greedy"] if precision == "fp32" else min(tol["greedy"], r["cal"][precision] - max(0.05, 1.5 / B))
    assert agree >= floor, rep
    # up to its first divergence every utterance IS the reference's trajectory; utterances that never diverge are identical
    assert (per_utt[first_div == S] == 1.0).all()
