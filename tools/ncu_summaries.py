"""Per-launch summary of an `ncu --set full` capture exported with `ncu -i rep --page raw --csv`.
usage: python tools/ncu_summaries.py raw.csv out.json "<command that produced the capture>"
Also prints the DRAM bytes per bench.py phase (what profiles/roofline_traffic.json holds)."""
import csv
import json
import re
import sys

rows = list(csv.reader(open(sys.argv[1])))
H, U = rows[0], rows[1]


def col(name):
    return H.index(name) if name in H else None


def num(r, name, scale=1.0):
    i = col(name)
    if i is None or r[i] in ("", "n/a"):
        return None
    v = float(r[i].replace(",", ""))
    u = U[i]
    if u.startswith("M"):
        v *= 1e6
    elif u.startswith("G"):
        v *= 1e9
    elif u.startswith("K") or u.startswith("k"):
        v *= 1e3
    return v * scale


def dur_us(r):
    i = col("gpu__time_duration.sum")
    v = float(r[i].replace(",", ""))
    u = U[i]
    return v / 1e3 if u.startswith("n") else v * (1e3 if u.startswith("m") else 1.0)


PHASE = [("neumf_fused_train", "fused_tile"), ("sample_negatives", "sampler"), ("rank_", "rank"), ("tc_wgrad_kernel", "tc_wgrad"), ("tc_dense_kernel<1, 1>", "tc_dense_bwd"), ("tc_dense_kernel<1, 2>", "tc_dense_bwd"),
         ("tc_dense_kernel", "tc_dense_fwd"), ("head_", "head"), ("h1_from", "h1_gather"), ("segreduce", "segreduce"),
         ("group_sum", "misc"), ("optimizer", "optimizer")]
out, phase_bytes = [], {}
for r in rows[2:]:
    name = re.sub(r"\(.*", "", r[col("Kernel Name")]).replace("void ", "").strip()
    rd, wr = num(r, "dram__bytes_read.sum") or 0.0, num(r, "dram__bytes_write.sum") or 0.0
    us = dur_us(r)
    e = {"kernel": name, "grid": r[col("Grid Size")], "duration_us": round(us, 1),
         "dram_read_MB": round(rd / 1e6, 1), "dram_write_MB": round(wr / 1e6, 1),
         "dram_GBs": round((rd + wr) / us / 1e3), "registers": r[col("launch__registers_per_thread")]}
    for key, metric in (("tensor_pipe_pct", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"),
                        ("lsu_data_pipe_pct", "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed"),
                        ("l2_pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed"),
                        ("warps_active_pct", "sm__warps_active.avg.pct_of_peak_sustained_active"),
                        ("issue_active_pct", "sm__inst_executed.avg.per_cycle_active")):
        v = num(r, metric)
        if v is not None:
            e[key] = round(v, 2)
    out.append(e)
    # B1 / B1u (plain dense over the item / user table into the gradient tables) belong to the backward phase
    ph = next((p for k, p in PHASE if k in name), "misc")
    if ph == "tc_dense_fwd" and out and len([x for x in out if x["kernel"].startswith("tc_dense_kernel<1, 0>")]) > 4:
        ph = "tc_dense_bwd"
    phase_bytes[ph] = phase_bytes.get(ph, 0.0) + rd + wr
json.dump({"command": sys.argv[3] if len(sys.argv) > 3 else "", "kernels": out}, open(sys.argv[2], "w"), indent=1)
for k, v in sorted(phase_bytes.items()):
    print("%-14s %8.1f MB of DRAM traffic per step" % (k, v / 1e6))
