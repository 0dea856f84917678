"""Regenerate profiles/conv_traffic.json: DRAM bytes per frame of the conv-kernel family of one DenseNet-121 forward.

Run on the GPU box:   python tools/measure_traffic.py [frames]
It launches `ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum` on tools/ncu_target.py, keeps the
launch list (profiles/<tag>_launches.csv), sums the bytes of the conv kernels (everything bench.py's ProfScope counts as
kProfConvGemm: conv_gemm / conv1x1_ts / conv3x3_halo / stem / dense_layer_fused) and writes the JSON bench.py reads, stamped with
the hash of the kernel sources so that a stale file is detected instead of silently reused."""
import csv
import datetime
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402  (kernel_source_hash)

frames = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
out_csv = os.path.join(ROOT, "profiles", "r2_launches_current.csv")
cmd = ["ncu", "--metrics", "gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum", "--clock-control", "none",
       "--profile-from-start", "off", "--csv", "--log-file", out_csv, sys.executable, os.path.join(ROOT, "tools", "ncu_target.py"),
       str(frames)]
subprocess.run(cmd, check=True, cwd=ROOT)
CONV = re.compile(r"conv_gemm_kernel|conv1x1_ts_kernel|conv3x3_halo_kernel|stem_pool_kernel|stem_s2d_kernel|dense_layer_fused_kernel")
rows = list(csv.reader(open(out_csv)))
hdr = None
tot = {"conv_bytes": 0.0, "conv_ns": 0.0, "all_bytes": 0.0, "all_ns": 0.0, "launches": set(), "conv_launches": set()}
for r in rows:
    if "Kernel Name" in r:
        hdr = r
        continue
    if not hdr or len(r) != len(hdr):
        continue
    d = dict(zip(hdr, r))
    v = float(d["Metric Value"].replace(",", ""))
    u = d["Metric Unit"]
    is_conv = bool(CONV.search(d["Kernel Name"]))
    tot["launches"].add(d["ID"])
    if is_conv:
        tot["conv_launches"].add(d["ID"])
    if d["Metric Name"] == "gpu__time_duration.sum":
        ns = v * {"ns": 1, "nsecond": 1, "us": 1e3, "usecond": 1e3, "ms": 1e6, "msecond": 1e6}.get(u, 1)
        tot["all_ns"] += ns
        if is_conv:
            tot["conv_ns"] += ns
    else:
        b = v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)
        tot["all_bytes"] += b
        if is_conv:
            tot["conv_bytes"] += b
res = {
    "frames": frames,
    "conv_dram_bytes_per_frame": tot["conv_bytes"] / frames,
    "all_dram_bytes_per_frame": tot["all_bytes"] / frames,
    "conv_kernel_ms_ncu": tot["conv_ns"] / 1e6,
    "all_kernel_ms_ncu": tot["all_ns"] / 1e6,
    "conv_launches": len(tot["conv_launches"]),
    "launches": len(tot["launches"]),
    "kernel_source_hash": bench.kernel_source_hash(),
    "when": datetime.datetime.utcnow().strftime("%Y-%m-%dT%H:%M:%SZ"),
    "launch_list": os.path.relpath(out_csv, ROOT),
    "how": "ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none on tools/ncu_target.py",
}
with open(os.path.join(ROOT, "profiles", "conv_traffic.json"), "w") as f:
    json.dump(res, f, indent=1)
print(json.dumps(res))
