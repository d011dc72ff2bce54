"""Context UMMA issue loop: descriptors advanced by constants (flags 0) vs rebuilt per instruction (flags 512).  Same instructions in
the same order, so the decoder's outputs must be bit-identical; then the timing of both on this box."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import las_testlib as tl  # noqa: E402
from las_pytorch_b200 import _cabi  # noqa: E402

lib = _cabi.load_library()
c = tl.CONFIGS["paper"]
ok = True
for B, T, S in ((64, 1600, 300), (5, 3000, 20), (3, 64, 6)):
    las = tl.build_model("paper", max_label_len=S, seed=17, gain=3.0, precision="bf16").cuda()
    x, _ = tl.make_inputs(B, T, c["F"], S, c["V"], seed=17)
    enc = las.listener(x.cuda())
    outs = []
    for f in (512, 0):
        lib.las_debug_set_option(5, f)
        pred, att = las.speller(enc, None, 0.0)
        outs.append((torch.stack(pred).clone(), torch.stack([a[0] for a in att]).clone(), las.speller.last_tokens.clone()))
    lib.las_debug_set_option(5, 0)
    same = all(torch.equal(a, b) for a, b in zip(*outs))
    ok &= same
    print(f"B={B} T={T} S={S}: bit-identical={same} finite={bool(torch.isfinite(outs[1][0]).all())}", flush=True)
    if B == 64:
        res = {}
        for rep in range(3):
            for f in (512, 0):
                lib.las_debug_set_option(5, f)
                las.speller(enc, None, 0.0)
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(5):
                    las.speller(enc, None, 0.0)
                e1.record()
                torch.cuda.synchronize()
                res.setdefault(f, []).append(e0.elapsed_time(e1) / 5 / S * 1e3)
        lib.las_debug_set_option(5, 0)
        for f in (512, 0):
            print(f"flags={f}: us/step min {min(res[f]):.3f} max {max(res[f]):.3f}", flush=True)
print("ALL BIT-IDENTICAL" if ok else "MISMATCH")
