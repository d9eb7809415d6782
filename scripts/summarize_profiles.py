"""Turn gpurun_out/*.ncu-rep and launch lists into small tracked summaries under profiles/.

    python scripts/summarize_profiles.py r01          # reads gpurun_out/, writes profiles/r01_*
"""
import csv
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KEEP = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_bytes.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "launch__grid_size", "launch__block_size",
    "launch__registers_per_thread", "sm__cycles_elapsed.max", "smsp__inst_executed.sum",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_tensor.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
]


def raw_metrics(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    if len(rows) < 3:
        return []
    hdr, units = rows[0], rows[1]
    kernels = []
    for vals in rows[2:]:
        d = {"kernel": vals[hdr.index("Kernel Name")]}
        stalls = {}
        for i, h in enumerate(hdr):
            if h in KEEP:
                d[h] = "%s %s" % (vals[i], units[i])
            if "issue_stalled" in h and h.endswith("_per_issue_active.ratio"):
                try:
                    stalls[h.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", "")] = float(vals[i])
                except ValueError:
                    pass
        d["top_stalls_warps_per_issue"] = dict(sorted(stalls.items(), key=lambda kv: -kv[1])[:6])
        kernels.append(d)
    return kernels


def main():
    tag = sys.argv[1] if len(sys.argv) > 1 else "r01"
    src = os.path.join(ROOT, "gpurun_out")
    dst = os.path.join(ROOT, "profiles")
    os.makedirs(dst, exist_ok=True)
    for name in sorted(os.listdir(src)):
        path = os.path.join(src, name)
        if name.endswith(".ncu-rep"):
            ks = raw_metrics(path)
            with open(os.path.join(dst, "%s_%s.json" % (tag, name[:-8])), "w") as fh:
                json.dump(ks, fh, indent=1)
            print(name, "->", len(ks), "kernel(s)")
        elif name.startswith("launches") and name.endswith(".csv"):
            with open(path) as fh, open(os.path.join(dst, "%s_%s" % (tag, name)), "w") as out:
                for line in fh:
                    if line.startswith('"') or line.startswith("ID"):
                        out.write(line)
            print(name, "copied")
            if name == "launches_warm.csv":
                # DRAM bytes per launch of the dominant kernel in steady state (caches NOT flushed between kernels)
                rows = [r for r in csv.reader(open(path)) if len(r) > 10]
                hdr = rows[0]
                ki, mi, vi, ii = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value"), hdr.index("ID")
                per = {}
                fused = False
                for r in rows[1:]:
                    if ("stft2048_kernel" in r[ki] or "stft2048_pair_kernel" in r[ki]) and r[mi].startswith("dram__bytes"):
                        # the one-kernel mel path: the pair kernel, or stft2048_kernel<OUT_MEL_FUSED>
                        fused = fused or "stft2048_pair_kernel" in r[ki] or "stft2048_kernel<3," in r[ki]
                        per.setdefault(r[ii], 0.0)
                        per[r[ii]] += float(r[vi].replace(",", ""))
                vals = list(per.values())[1:] or list(per.values())      # drop the first launch (follows the RNG fill)
                if vals:
                    with open(os.path.join(dst, "%s_%s_dram_bytes.json" % (tag, "melfused" if fused else "stft2048")), "w") as fh:
                        sys.path.insert(0, ROOT)
                        import build_native
                        json.dump({"dram_bytes_per_launch": sum(vals) / len(vals), "launches": len(vals), "source": name,
                                   "source_fingerprint": build_native._fingerprint()[:16],
                                   "note": "dram__bytes_read.sum + dram__bytes_write.sum per launch of the dominant kernel, ncu --cache-control "
                                           "none (steady state of bench.py, 20032 frames per launch)"}, fh)
        elif name.startswith("bench_") and name.endswith(".json"):
            with open(path) as fh, open(os.path.join(dst, "%s_%s" % (tag, name)), "w") as out:
                out.write(fh.read())


if __name__ == "__main__":
    main()
