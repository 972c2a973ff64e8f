"""Turns the ncu outputs of tools/capture_profiles.sh into the small files under profiles/ that bench.py and the docs quote.

    python tools/summarize_profiles.py gpurun_out profiles

* <src>/r2_ncu_launches_deepfm.csv  -> profiles/roofline_traffic.json["deepfm"]: DRAM bytes (read + write) per C-ABI call,
  median per launch summed over the call's kernels (ncu: cold cache, serialised) — what bench.py copies into
  roofline.traffic for the dominant call; plus profiles/r2_launch_summary.json (median duration / bytes per kernel).
* <src>/r2_ncu_launches_topk.csv    -> roofline_traffic.json["retrieval"].
* <src>/r2_ncu_full_*.raw.csv       -> profiles/r2_ncu_full_summary.json: the counters the docs quote.
"""
import csv
import json
import os
import statistics
import sys

API_OF_KERNEL = [  # substring of the kernel name -> C-ABI call (bench.py's per-call names)
    ("embed_pool_fwd_kernel", "nrx_embed_pool_fwd_img"), ("tower_fwd3_kernel", "nrx_tower_fwd_head"),
    ("tower_bwd_dx3_kernel", "nrx_tower_bwd_dx"), ("tower_bwd_dw_kernel", "nrx_tower_bwd_dw"),
    ("tower_bwd_reduce_entry", "nrx_tower_bwd_dw"), ("segment_reduce_kernel", "nrx_embed_bwd_apply"),
    ("segment_fixup_kernel", "nrx_embed_bwd_apply"), ("plan_chunk_sort_kernel", "nrx_embed_bwd_plan"),
    ("plan_merge_kernel", "nrx_embed_bwd_plan"), ("adamw_kernel", "nrx_adamw_dense_dev"),
    ("field_logit_bwd", "nrx_field_logit_bwd"), ("field_logit_fwd", "nrx_field_logit_fwd"),
    ("tower_pack_kernel", "nrx_tower_pack"), ("hparams_step_kernel", "nrx_hparams_step"), ("reduce2_kernel", "nrx_reduce2_f32"),
    ("topk_scan_kernel", "nrx_topk_search"), ("topk_theta_kernel", "nrx_topk_search"), ("topk_final_kernel", "nrx_topk_search"),
    ("topk_qpack_kernel", "nrx_topk_search"), ("topk_exact", "nrx_topk_search"),
]
COUNTERS = ["gpu__time_duration.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
            "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
            "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__registers_per_thread", "launch__grid_size",
            "launch__block_size", "lts__t_sector_hit_rate.pct", "smsp__inst_executed.sum",
            "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
            "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
            "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
            "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
            "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
            "sm__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"]
UNIT_SCALE = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "nsecond": 1e-3, "ns": 1e-3, "usecond": 1.0, "us": 1.0,
              "msecond": 1e3, "ms": 1e3, "second": 1e6}


def launch_list(path):
    """-> {kernel base name: {"n": launches, "us": median duration, "bytes": median DRAM read + write}}"""
    rows = list(csv.reader(open(path)))
    hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"]
    if not hi:
        return {}
    hdr, data = rows[hi[0]], rows[hi[0] + 1:]
    iK, iM, iU, iV, iID = (hdr.index(c) for c in ("Kernel Name", "Metric Name", "Metric Unit", "Metric Value", "ID"))
    per = {}
    for r in data:
        if len(r) <= iV:
            continue
        d = per.setdefault(r[iID], {"k": r[iK]})
        d[r[iM]] = float(r[iV].replace(",", "")) * UNIT_SCALE.get(r[iU], 1.0)
    out = {}
    for d in per.values():
        name = d["k"].split("(")[0].replace("void ", "").replace("nrx::", "")
        e = out.setdefault(name, {"us": [], "bytes": []})
        e["us"].append(d.get("gpu__time_duration.sum", 0.0))
        e["bytes"].append(d.get("dram__bytes_read.sum", 0.0) + d.get("dram__bytes_write.sum", 0.0))
    return {k: {"n": len(v["us"]), "us": round(statistics.median(v["us"]), 2), "bytes": statistics.median(v["bytes"])}
            for k, v in out.items()}


def per_api(kernels):
    out = {}
    for name, v in kernels.items():
        for sub, api in API_OF_KERNEL:
            if sub in name:
                out[api] = out.get(api, 0.0) + v["bytes"]
                break
    return out


def full_capture(path):
    rows = list(csv.reader(open(path)))
    if len(rows) < 3:
        return None
    hdr, units, vals = rows[0], rows[1], rows[2]
    d = {"kernel": vals[hdr.index("Kernel Name")].split("(")[0]}
    for c in COUNTERS:
        if c in hdr:
            i = hdr.index(c)
            d[c] = {"value": float(vals[i].replace(",", "")) if vals[i] not in ("", "n/a") else None, "unit": units[i]}
    return d


def main():
    src, dst = (sys.argv + ["gpurun_out", "profiles"])[1:3]
    tpath = os.path.join(dst, "roofline_traffic.json")
    traffic = json.load(open(tpath)) if os.path.exists(tpath) else {}
    summary = {}
    for wl, f in (("deepfm", "r2_ncu_launches_deepfm.csv"), ("retrieval", "r2_ncu_launches_topk.csv")):
        p = os.path.join(src, f)
        if not os.path.exists(p):
            continue
        k = launch_list(p)
        summary[wl] = k
        traffic[wl] = per_api(k)
    traffic["_source"] = ("profiles/r2_launch_summary.json <- r2_ncu_launches_{deepfm,topk}.csv (ncu dram__bytes_read.sum + "
                          "dram__bytes_write.sum, median per launch, summed over the kernels of the call; cold cache, serialised)")
    json.dump(traffic, open(tpath, "w"), indent=1)
    json.dump(summary, open(os.path.join(dst, "r2_launch_summary.json"), "w"), indent=1)
    full = {}
    for f in sorted(os.listdir(src)):
        if f.startswith("r2_ncu_full_") and f.endswith(".raw.csv"):
            d = full_capture(os.path.join(src, f))
            if d:
                full[f[len("r2_ncu_full_"):-len(".raw.csv")]] = d
    if full:
        json.dump(full, open(os.path.join(dst, "r2_ncu_full_summary.json"), "w"), indent=1)
    print(json.dumps({"traffic": {k: v for k, v in traffic.items() if k != "_source"}, "full": list(full)}, indent=1)[:3000])


if __name__ == "__main__":
    main()
