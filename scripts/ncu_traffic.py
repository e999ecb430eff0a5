"""Build profiles/ncu_traffic.json (DRAM bytes per launch of every conv instantiation) and a per-launch table from the
raw CSV of one `ncu --set full` capture of a predict step (scripts/gpu_profile.sh).

    python scripts/ncu_traffic.py gpurun_out/full_step_raw.csv gpurun_out/launch_order.txt profiles/ncu_traffic.json profiles/r1_ncu_step_v10.md
"""
import csv
import json
import re
import sys


def norm(name):
    m = re.search(r"(conv_tc_kernel|conv_row_kernel)<([^>]*)>", name)
    if not m:
        return re.sub(r"\(.*", "", name).replace("void ", "").replace("rsb::", "")
    args = [a.strip() for a in m.group(2).split(",")]
    conv = {"true": "1", "false": "0"}
    args = [conv.get(a, a) for a in args]
    keep = 5 if m.group(1) == "conv_tc_kernel" else 3
    return "%s<%s>" % (m.group(1), ",".join(args[:keep]))


def main():
    raw, order, out_json, out_md = sys.argv[1:5]
    rows = list(csv.reader(open(raw)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    idx = {h: i for i, h in enumerate(hdr)}
    # ncu scales the unit of every column to the capture (us / ms, Mbyte / Gbyte ...): bring everything to us and MB (GB for TMA)
    SCALE = {"ns": 1e-3, "us": 1.0, "usecond": 1.0, "ms": 1e3, "msecond": 1e3, "s": 1e6, "second": 1e6,
             "byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3, "Tbyte": 1e6}

    def col(r, key):
        for h in ([key] if key in idx else []) + [h for h in hdr if h.endswith("." + key)]:
            try:
                return float(r[idx[h]].replace(",", "")) * SCALE.get(units[idx[h]], 1.0)
            except ValueError:
                continue
        return 0.0

    names = open(order).read().split("LAUNCH_ORDER ")[-1].strip().split(",")
    agg = {}
    title = sys.argv[5] if len(sys.argv) > 5 else "one predict step"
    md = ["# %s under `ncu --set full --clock-control none` (batch 32 x 3x512x512), per launch" % title, "",
          "Cold-cache, serialised launches: use the shares and the per-launch counters, not the absolute times (bench.py times the real step).", "",
          "| # | layer | kernel | us | tensor pipe % | DRAM MB (r+w) | L2->SM GB (TMA) | DRAM % |", "|---|---|---|---|---|---|---|---|"]
    for i, r in enumerate(data):
        k = norm(r[idx["Kernel Name"]])
        us = col(r, "gpu__time_duration.sum")
        mb = col(r, "dram__bytes_read.sum") + col(r, "dram__bytes_write.sum")
        tens = col(r, "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed")
        l2 = col(r, "l1tex__m_xbar2l1tex_read_bytes_mem_global_op_tma_ld.sum") / 1e3  # MB -> GB
        dpct = col(r, "dram__bytes_read.sum.pct_of_peak_sustained_elapsed") + col(r, "dram__bytes_write.sum.pct_of_peak_sustained_elapsed")
        a = agg.setdefault(k, {"mb": 0.0, "launches": 0, "us_total": 0.0})
        a["mb"] += mb
        a["launches"] += 1
        a["us_total"] += us
        md.append("| %d | %s | `%s` | %.1f | %.1f | %.1f | %.2f | %.1f |" % (i, names[i] if i < len(names) else "?", k, us, tens, mb, l2, dpct))
    out = {"source": "%s (ncu --set full, one predict step, batch 32 x 3x512x512)" % out_md,
           "unit": "MB per launch (dram__bytes_read.sum + dram__bytes_write.sum, averaged over the instantiation's launches in the step)",
           "kernels": {k: {"mb_per_launch": round(v["mb"] / v["launches"], 2), "launches": v["launches"], "us_total": round(v["us_total"], 1)}
                       for k, v in agg.items() if k.startswith("conv_")}}
    json.dump(out, open(out_json, "w"), indent=1)
    open(out_md, "w").write("\n".join(md) + "\n")
    print(json.dumps(out["kernels"], indent=1))


if __name__ == "__main__":
    main()
