"""Summarise an `ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --csv` launch list of
`bench.py`: cut out ONE training step (first input kernel of a forward pass ... the kernel before the next forward pass), print per-kernel
time share and DRAM traffic, and write
  <out>.step.csv   the launches of that step (id, kernel, ns, dram read, dram write)
  <out>.json       per-family totals; bench.py reads "conv" -> roofline.traffic (DRAM bytes per conv launch)
usage: python tools/launch_summary.py gpurun_out/launches.csv profiles/r1_launches_final"""
import collections
import csv
import json
import sys


def short(name):
    return name.split("(")[0].replace("void ", "").replace("b2::", "")


def main():
    src, out = sys.argv[1], sys.argv[2]
    rows = list(csv.reader(open(src)))
    hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    hdr = rows[hi]
    launches = collections.OrderedDict()
    for r in rows[hi + 1:]:
        if len(r) < len(hdr):
            continue
        rec = dict(zip(hdr, r))
        e = launches.setdefault(int(rec["ID"]), {"kernel": short(rec["Kernel Name"]), "grid": rec.get("Grid Size", "")})
        e[rec["Metric Name"]] = float(rec["Metric Value"].replace(",", ""))
    seq = list(launches.values())
    starts = [i for i, e in enumerate(seq) if e["kernel"].startswith(("cast_input_kernel", "im2col_input_kernel"))]
    # a step runs from the first input kernel of a forward pass to the kernel before the next forward pass's (Adam runs in buckets on
    # a side stream beside backward, so "the first adam_kernel" is no longer the end of a step); input kernels that follow each
    # other directly (im2col view + plain cast of the same batch) open the same step
    heads = [s for j, s in enumerate(starts) if j == 0 or s != starts[j - 1] + 1]
    step = None
    if len(heads) >= 3:
        step = seq[heads[1]:heads[2]]            # the second step of the capture (the first one also pays one-time work)
    elif len(heads) == 2:
        step = seq[heads[0]:heads[1]]
    if step is None or not any(e["kernel"].startswith("adam_kernel") for e in step):
        raise SystemExit("no complete step (input kernel ... next input kernel, with an adam_kernel inside) in the capture")
    with open(out + ".step.csv", "w") as f:
        w = csv.writer(f)
        w.writerow(["i", "kernel", "grid", "gpu__time_duration.sum [ns]", "dram__bytes_read.sum", "dram__bytes_write.sum"])
        for i, e in enumerate(step):
            w.writerow([i, e["kernel"], e["grid"], int(e.get("gpu__time_duration.sum", 0)), int(e.get("dram__bytes_read.sum", 0)),
                        int(e.get("dram__bytes_write.sum", 0))])
    fam = collections.OrderedDict()
    for e in step:
        k = e["kernel"].split("<")[0]
        f = fam.setdefault(k, {"launches": 0, "ms": 0.0, "dram_read_bytes": 0.0, "dram_write_bytes": 0.0})
        f["launches"] += 1
        f["ms"] += e.get("gpu__time_duration.sum", 0) / 1e6
        f["dram_read_bytes"] += e.get("dram__bytes_read.sum", 0)
        f["dram_write_bytes"] += e.get("dram__bytes_write.sum", 0)
    total = sum(f["ms"] for f in fam.values())
    for k, f in fam.items():
        f["share_of_step"] = f["ms"] / total
    conv = [f for k, f in fam.items() if k.startswith("conv_")]
    wg = [f for k, f in fam.items() if k.startswith("wgrad")]

    def tot(fs):
        return {"launches": sum(f["launches"] for f in fs), "ms": sum(f["ms"] for f in fs),
                "dram_bytes": sum(f["dram_read_bytes"] + f["dram_write_bytes"] for f in fs),
                "share_of_step": sum(f["ms"] for f in fs) / total}
    summary = {"source": src, "kernels_in_step": len(step), "step_ms_under_ncu": total, "families": fam, "conv": tot(conv), "wgrad": tot(wg)}
    summary["conv"]["dram_bytes_per_launch"] = summary["conv"]["dram_bytes"] / max(1, summary["conv"]["launches"])
    summary["wgrad"]["dram_bytes_per_launch"] = summary["wgrad"]["dram_bytes"] / max(1, summary["wgrad"]["launches"])
    with open(out + ".json", "w") as f:
        json.dump(summary, f, indent=1)
    print(f"one step = {len(step)} kernels, {total:.3f} ms under ncu (cold caches, serialised)")
    for k, f in sorted(fam.items(), key=lambda kv: -kv[1]["ms"]):
        print(f"{k:<26} x{f['launches']:<3d} {f['ms']:7.3f} ms {100 * f['share_of_step']:5.1f}%  read {f['dram_read_bytes'] / 1e6:8.1f} MB"
              f"  write {f['dram_write_bytes'] / 1e6:8.1f} MB")


if __name__ == "__main__":
    main()
