"""Sustained tensor-pipe rate of the 3xTF32 stage pattern (GPU only, diagnostics):
cycles per 128 x N x 8 TF32 MMA with operands in shared memory (SS) or A in TMEM (TS), K-major or MN-major,
rotating over nbuf stage buffers, with and without concurrent st.shared traffic from `writers` warps.
Run: gpurun -- python tools/tc_rate.py"""
import ctypes as C
import os
import sys

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "movierecommender-tf-trt_b200"))
from movierec import _diag as nat  # noqa: E402


def rate(N, iters=2000, nbuf=3, flags=0, writers=0, write_iters=0, grid=148):
    out = torch.zeros(grid, dtype=torch.int64, device="cuda")
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    for _ in range(2):
        nat.check(nat.lib.mr_tc_rate(N, iters, nbuf, flags, writers, write_iters, C.c_void_p(out.data_ptr()), grid, st), "tc_rate")
    torch.cuda.synchronize()
    cyc = out.cpu().numpy()
    return float(cyc.mean()) / (iters * 12), float(cyc.max()) / (iters * 12)


if __name__ == "__main__":
    print("cycles per MMA (mean over CTAs, max); ideal = 128*N/256")
    for N in (64, 128, 256):
        for flags, name in ((0, "SS K-major"), (2, "SS MN-major"), (1, "TS (A in TMEM) K-major B")):
            for nbuf in (1, 3):
                if (2 * 128 * 32 * 4 + 2 * N * 32 * 4) * nbuf > 200 * 1024:
                    continue
                m, mx = rate(N, nbuf=nbuf, flags=flags)
                print("N={:3d} {:28s} nbuf={} : {:6.1f} {:6.1f}   (ideal {:.0f})".format(N, name, nbuf, m, mx, 128 * N / 256))
    # producer-like shared-memory store traffic next to the MMAs: each writer warp issues 4 x 512 B per iteration
    for N in (128,):
        for flags, name in ((0, "SS"), (1, "TS")):
            for writers in (2, 4, 8):
                # per stage the real producers store 32 KB (A hi/lo): 64 warp-instructions of 512 B
                wi = 2000 * 64 // (4 * writers)
                m, mx = rate(N, nbuf=3, flags=flags, writers=writers, write_iters=wi)
                print("N={} {} + {} writer warps storing 32 KB per stage: {:6.1f} {:6.1f}".format(N, name, writers, m, mx))
    # a tcgen05.commit after every stage of 12 MMAs (flags bit 2), as the pipelined kernels issue them
    for flags, name in ((4, "SS + commit per stage"), (5, "TS + commit per stage"), (6, "SS MN-major + commit"), (7, "TS, MN-major B + commit")):
        m, mx = rate(128, nbuf=3, flags=flags)
        print("N=128 {:28s}: {:6.1f} {:6.1f}".format(name, m, mx))
    for flags, name in ((7 | 8, "TS MN-B commit, A at column 128"), (7 | 16, "TS MN-B commit, A ring of 4 stages"), (7 | 8 | 16, "TS MN-B commit, col 128 + ring")):
        m, mx = rate(128, nbuf=3, flags=flags)
        print("N=128 {:36s}: {:6.1f} {:6.1f}".format(name, m, mx))
    m, mx = rate(128, nbuf=3, flags=0, grid=1)
    print("single CTA, N=128 SS: {:.1f}".format(m))
