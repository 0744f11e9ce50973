"""Per-role, per-segment cycle breakdown of the fused train kernel (diagnostics; needs a library built with
-DMR_FUSED_TIMING: tools/build_variant.sh timing -DMR_FUSED_TIMING, then
MR_LIB_PATH=variants/timing/libmovierec_b200.so python tools/fused_timing.py)."""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "movierecommender-tf-trt_b200"))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from movierec import _engine, _native as nat  # noqa: E402

SEG = {
    "producers": ["ids/prefetch", "issue 2 H1 tasks", "wait h1_free", "-", "H1 convert/store", "GMF part 1",
                  "wait dz2_full", "GMF part 2", "cp.async wait + bar"],
    "epilogue 1 (+ MMA issue)": ["issue fwd + wait fwd_done", "E1 ld + zdot", "wait gmf_ready", "E1 rest",
                                 "issue wgrad + bwd"],
    "epilogue 2": ["wait bwd_done", "E2"],
}


def main():
    nu, ni, L, f, negs = 138493, 26744, [256, 128, 64], 64, 4
    groups = 262144
    rng = np.random.default_rng(0)
    eng = _engine.NeuMFEngine(nu, ni, L, [0, 0, 0], mf_dim=f, seed=1)
    users = np.repeat(rng.integers(0, nu, groups), negs + 1)
    items = rng.integers(0, ni, groups * (negs + 1))
    y = np.tile([0] * negs + [1], groups).astype(np.float32)
    for _ in range(3):
        eng.train_step(users, items, y, group=negs + 1, k=10, grouped=True)
    torch.cuda.synchronize()
    out = (C.c_longlong * (256 * 48))()
    assert nat.lib.mr_fused_timing_read(out) == 0
    t = np.array(out[:]).reshape(256, 3, 16)[:148]
    tiles = (groups * 5 + 119) // 120 / 148.0
    for r, role in enumerate(SEG):
        tot = t[:, r, :].sum(axis=1).mean()
        print("{} (mean over CTAs; {:.0f} cycles per tile)".format(role, tot / tiles))
        for s, name in enumerate(SEG[role]):
            print("   {:18s} {:8.0f} cycles/tile  {:5.1f} %".format(name, t[:, r, s].mean() / tiles, 100 * t[:, r, s].mean() / tot))


if __name__ == "__main__":
    main()
