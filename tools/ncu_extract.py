"""Reads `ncu --set full` reports (CPU box, no GPU needed) and writes (1) a compact per-kernel CSV of the metrics the
design discussion uses into profiles/, (2) profiles/ncu_traffic.json: DRAM bytes per launch for the kernels bench.py's
`roofline.traffic` names, tagged with the capture and the digest of the sources the library was built from.

    python tools/ncu_extract.py gpurun_out/foo.ncu-rep profiles/round2x_ncu_foo.csv [--traffic name=kernel_regex ...]
"""
import csv
import datetime
import io
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KEEP = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__data_pipe_lsu_wavefronts.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__t_bytes.sum",
    "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum",
    "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_tc_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "smsp__average_warp_latency_per_inst_issued.ratio",
]
STALLS = re.compile(r"smsp__average_warps_issue_stalled_(\w+)_per_issue_active\.ratio")


def read(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    return hdr, units, rows[2:]


def main():
    rep, dst = sys.argv[1], sys.argv[2]
    traffic = {}
    if "--traffic" in sys.argv:
        for spec in sys.argv[sys.argv.index("--traffic") + 1:]:
            name, rx = spec.split("=", 1)
            traffic[name] = re.compile(rx)
    hdr, units, rows = read(rep)
    col = {h: i for i, h in enumerate(hdr)}
    kname = col.get("Kernel Name")
    keep = [h for h in KEEP if h in col] + [h for h in hdr if STALLS.match(h)]
    os.makedirs(os.path.dirname(dst), exist_ok=True)
    with open(dst, "w", newline="") as fh:
        w = csv.writer(fh)
        w.writerow(["kernel", "metric", "value", "unit"])
        for r in rows:
            for h in keep:
                w.writerow([r[kname][:90], h, r[col[h]], units[col[h]]])
    print(f"wrote {dst}: {len(rows)} launch(es), {len(keep)} metrics each")
    if traffic:
        path = os.path.join(ROOT, "profiles", "ncu_traffic.json")
        try:
            cur = json.load(open(path))
        except (OSError, ValueError):
            cur = {"traffic_mb": {}, "sources": {}}

        def to_mb(v, u):
            v = float(v.replace(",", ""))
            return v * {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3}[u]
        for name, rx in traffic.items():
            hit = [r for r in rows if rx.search(r[kname])]
            if not hit:
                print(f"  traffic {name}: no launch matches {rx.pattern}")
                continue
            mb = [to_mb(r[col["dram__bytes_read.sum"]], units[col["dram__bytes_read.sum"]]) +
                  to_mb(r[col["dram__bytes_write.sum"]], units[col["dram__bytes_write.sum"]]) for r in hit]
            cur["traffic_mb"][name] = sum(mb) / len(mb)
            cur["sources"][name] = os.path.basename(rep)
            print(f"  traffic {name}: {cur['traffic_mb'][name]:.1f} MB per launch ({len(hit)} launch(es))")
        sys.path.insert(0, ROOT)
        from boostmvsnerfs_b200 import build as _b
        cur["csrc_digest"] = _b._digest()[:16]
        cur["workload"] = os.environ.get("WORKLOAD", "C2")          # what tools/frame_ab.py rendered under ncu
        cur["capture"] = sorted(set(cur["sources"].values()))
        cur["when"] = datetime.datetime.now(datetime.timezone.utc).strftime("%Y-%m-%dT%H:%MZ")
        json.dump(cur, open(path, "w"), indent=1)


if __name__ == "__main__":
    main()
