"""Cost of one persistent-decoder launch besides its steps (prologue: weights -> shared / tensor memory, enc^T -> tensor memory, psi):
time(S steps) for small S, extrapolated to S = 0."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import las_testlib as tl  # noqa: E402

cfgname, B, T = "paper", 64, 1600
c = tl.CONFIGS[cfgname]
x, _ = tl.make_inputs(B, T, c["F"], 4, c["V"], seed=17)
res = {}
for S in (2, 4, 8, 300):
    las = tl.build_model(cfgname, max_label_len=S, seed=17, gain=3.0, precision="bf16").cuda()
    enc = las.listener(x.cuda())
    for _ in range(3):
        las.speller(enc, None, 0.0)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        las.speller(enc, None, 0.0)
    e1.record()
    torch.cuda.synchronize()
    res[S] = e0.elapsed_time(e1) / 10 * 1e3
    print(f"S={S}: {res[S]:.1f} us per decode call")
per = (res[8] - res[2]) / 6
print(f"per step {per:.2f} us -> fixed cost per call (prepare + psi + launch + prologue) {res[2] - 2 * per:.1f} us")
