"""Descriptor explorer (GPU only): prints which shared-memory byte the tensor core fetches for logical
element (m, k) of the A operand under a given (major, LBO, SBO).  This is how the MN-major no-swizzle
semantics used by csrc/tc_common.cuh were pinned.  Run: gpurun -- python tools/tc_probe.py"""
import ctypes as C
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "movierecommender-tf-trt_b200"))
from movierec import _diag as nat  # noqa: E402

NW = 8192


def probe(start, lbo, sbo, a_mn):
    w = np.arange(NW)
    outs = []
    for vals in (w % 1024 + 1, w // 1024 + 1):
        raw = torch.from_numpy(vals.astype(np.float32)).cuda()
        D = torch.zeros(128 * 16, dtype=torch.float32, device="cuda")
        st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
        nat.check(nat.lib.mr_tc_probe(C.c_void_p(raw.data_ptr()), NW, start, lbo, sbo, a_mn, C.c_void_p(D.data_ptr()), st), "probe")
        torch.cuda.synchronize()
        outs.append(D.cpu().numpy().reshape(128, 16)[:, :8])
    ok = (outs[0] > 0) & (outs[1] > 0)
    res = ((outs[0] - 1) + 1024 * (outs[1] - 1)).astype(np.int64)  # word index fetched for (m, k)
    res[~ok] = -1
    return res


if __name__ == "__main__":
    for (lbo, sbo, mn, lt) in ((128, 1024, 0, 0), (4096, 128, 1, 0), (128, 4096, 1, 0), (4096, 1024, 1, 2), (1024, 4096, 1, 2),
                               (4096, 1024, 1, 1), (1024, 4096, 1, 1), (4096, 256, 1, 6), (256, 4096, 1, 6), (4096, 512, 1, 4)):
        W = probe(0, lbo | (lt << 24), sbo, mn)
        print("a_mn", mn, "lbo", lbo, "sbo", sbo, "layout_type", lt)
        print("  byte offsets m=0, k=0..7:", (W[0] * 4).tolist())
        print("  byte offsets k=0, m=0..11:", (W[:12, 0] * 4).tolist())
        print("  byte offsets k=0, m=32,64,96,127:", (W[[32, 64, 96, 127], 0] * 4).tolist())
        print("  byte offsets k=1, m=0..7:", (W[:8, 1] * 4).tolist())
