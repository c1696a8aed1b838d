"""Summarise an `ncu --set full` report into the per-launch table kept under profiles/ (run where ncu is installed).

usage: python tools/ncu_summary.py REPORT.ncu-rep OUT.txt [--layers] [--batch B] [--traffic profiles/traffic.json] [--title "..."]

--layers   the report holds one model call (16 conv launches, in execution order): label rows with the layer names and
           add the sum line; with --traffic also writes the DRAM-traffic table bench.py reads for roofline.traffic.
"""
import csv
import io
import json
import os
import subprocess
import sys

LAYERS = ["inc.0", "down1.0.0", "down1.0.2", "down2.0.0", "down2.0.2", "down3.0.0", "down3.0.2", "up3", "conv3.0", "conv3.2",
          "up2", "conv2.0", "conv2.2", "up1", "conv1.0", "conv1.2"]
COLS = [("ms", "gpu__time_duration.sum", "ms"),
        ("tensor_pipe%", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", None),
        ("utchmma_f16%", "sm__ops_path_tensor_op_utchmma_src_fp16_dst_fp32_sparsity_off.avg.pct_of_peak_sustained_elapsed", None),
        ("dram_rd_GB", "dram__bytes_read.sum", "Gbyte"),
        ("dram_wr_GB", "dram__bytes_write.sum", "Gbyte"),
        ("dram%", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", None),
        ("lts%", "lts__throughput.avg.pct_of_peak_sustained_elapsed", None),
        ("sm%", "sm__throughput.avg.pct_of_peak_sustained_elapsed", None),
        ("issue%", "sm__issue_active.avg.pct_of_peak_sustained_elapsed", None),
        ("warps%", "sm__warps_active.avg.pct_of_peak_sustained_active", None),
        ("SM_GHz", "sm__cycles_elapsed.avg.per_second", "Ghz"),
        ("regs", "launch__registers_per_thread", None)]
SCALE = {"us": 1e-3, "ms": 1.0, "s": 1e3, "Kbyte": 1e-6, "Mbyte": 1e-3, "Gbyte": 1.0, "byte": 1e-9, "Ghz": 1.0, "Mhz": 1e-3}


def main():
    rep, out = sys.argv[1], sys.argv[2]
    layers = "--layers" in sys.argv
    traffic = sys.argv[sys.argv.index("--traffic") + 1] if "--traffic" in sys.argv else None
    title = sys.argv[sys.argv.index("--title") + 1] if "--title" in sys.argv else rep
    batch = int(sys.argv[sys.argv.index("--batch") + 1]) if "--batch" in sys.argv else 64
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, body = rows[0], rows[1], rows[2:]

    def col(name):
        hits = [i for i, h in enumerate(hdr) if h == name]
        return hits[0] if hits else None

    idx = {label: col(name) for label, name, _ in COLS}
    lines = [f"# {title}", "launch     kernel" + " " * 47 + "".join(f"{label:>13s}" for label, _, _ in COLS)]
    tot_ms, tj = 0.0, {"layers": {}}
    for n, r in enumerate(body):
        vals = []
        for label, name, want in COLS:
            i = idx[label]
            if i is None or r[i] == "":
                vals.append(float("nan"))
                continue
            v = float(r[i])
            if want is not None:
                v *= SCALE.get(units[i], 1.0)
            vals.append(v)
        tag = LAYERS[n] if layers and n < len(LAYERS) else str(n)
        lines.append(f"{tag:10s} {r[4][:52]:52s}" + "".join(f"{v:13.3f}" for v in vals))
        by = {label: v for (label, _, _), v in zip(COLS, vals)}
        tot_ms += by["ms"]
        if layers and n < len(LAYERS):
            tj["layers"][LAYERS[n]] = {"dram_read_GB": round(by["dram_rd_GB"], 6), "dram_write_GB": round(by["dram_wr_GB"], 6)}
    if layers:
        lines.append(f"# sum of the {len(body)} launches: {tot_ms:.3f} ms for {batch} patch(es) x 301.85 GFLOP = "
                     f"{batch * 301.851 / tot_ms:.0f} TFLOP/s (cold, serialised under the profiler)")
    open(out, "w").write("\n".join(lines) + "\n")
    print("\n".join(lines))
    if traffic and layers:
        rd = sum(v["dram_read_GB"] for v in tj["layers"].values())
        wr = sum(v["dram_write_GB"] for v in tj["layers"].values())
        sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
        import bench
        res = {"source": f"{out} (ncu --set full, one model call = 16 launches, batch 64)", "per": "model call of 64 patches",
               "csrc_fingerprint": bench.csrc_fingerprint(),     # bench.py quotes this capture only while the kernels are unchanged
               "git_head": os.environ.get("S1S2_GIT_HEAD"),
               "dram_read_GB": round(rd, 6), "dram_write_GB": round(wr, 6), "dram_total_GB": round(rd + wr, 6),
               "algorithmic_activation_GB": 18.624, "layers": tj["layers"]}
        json.dump(res, open(traffic, "w"), indent=1)


if __name__ == "__main__":
    main()
