"""Device split + per-user item lists on the ML-20M shape, next to the NumPy statement of the same
rule on the host cores.  usage (GPU box): python tools/dataset_prep_bench.py > gpurun_out/dataset_prep.json"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "movierecommender-tf-trt_b200"))
import torch  # noqa: E402
from movierec import _engine, data_pipeline  # noqa: E402

nu, ni, n = 138493, 26744, 20000263
rng = np.random.default_rng(0)
users = np.concatenate([np.repeat(np.arange(nu), 20), rng.integers(0, nu, n - 20 * nu)]).astype(np.int32)
rng.shuffle(users)
items = np.minimum(rng.zipf(1.1, n) - 1, ni - 1).astype(np.int32)
du, di = torch.from_numpy(users).cuda(), torch.from_numpy(items).cuda()


def timed(fn, reps=5):
    fn()
    torch.cuda.synchronize()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(reps)]
    for a, b in ev:
        a.record()
        fn()
        b.record()
    torch.cuda.synchronize()
    return float(np.median([a.elapsed_time(b) for a, b in ev]))


ms_split = timed(lambda: _engine.split_last_two(du, nu))
ms_csr = timed(lambda: _engine.build_user_csr(du, di, nu, ni))
rowptr, csr = _engine.build_user_csr(du, di, nu, ni)
t0 = time.perf_counter()
order = np.argsort(users, kind="stable")
su = users[order]
last = np.r_[su[1:] != su[:-1], True]
t_split_cpu = (time.perf_counter() - t0) * 1e3
t0 = time.perf_counter()
h_rowptr, h_csr = data_pipeline.build_user_csr(users, items)
t_csr_cpu = (time.perf_counter() - t0) * 1e3
assert np.array_equal(rowptr.cpu().numpy(), h_rowptr) and np.array_equal(csr.cpu().numpy(), h_csr)
out = {"workload": "ml-20m shape: {} ratings, {} users, {} items (Zipf items, >= 20 ratings per user)".format(n, nu, ni),
       "split_last_two": {"device_ms": ms_split, "ratings_per_s": n / ms_split * 1e3,
                          "algorithmic_bytes": 12 * n, "algorithmic_gbs": 12 * n / ms_split / 1e6,
                          "numpy_ms_host": t_split_cpu},
       "build_user_csr": {"device_ms": ms_csr, "ratings_per_s": n / ms_csr * 1e3, "distinct_pairs": int(csr.numel()),
                          "algorithmic_bytes": 8 * n + 4 * int(csr.numel()) + 8 * (nu + 1),
                          "algorithmic_gbs": (8 * n + 4 * int(csr.numel()) + 8 * (nu + 1)) / ms_csr / 1e6,
                          "numpy_ms_host": t_csr_cpu, "equal_to_numpy_table": True},
       "host_cores": os.cpu_count(),
       "note": "device times include the workspace allocation of the Python wrapper; bytes = ids in + results out "
               "(the radix sorts' own passes are implementation traffic)"}
print(json.dumps(out))
