"""tools/traffic_from_profile.py -- profiles/traffic_<cfg>.json from an ncu summary (profiles/summarize.py output):
dram__bytes_read.sum + dram__bytes_write.sum of the k_render launch, tied to the SASS hash of the kernel that was
measured (profiles/sass_summary.json of the same build).  bench.py refuses the file when the hash no longer matches.

    python tools/traffic_from_profile.py cfg3 profiles/r02_a_render_cfg3.csv"""
import csv
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
cfg, path = sys.argv[1], sys.argv[2]
rows = list(csv.reader(open(path)))
hdr = rows[0]
row = next(r for r in rows[1:] if "k_render<0>" in r[0])
col = lambda key: float(row[next(i for i, h in enumerate(hdr) if h.startswith(key))])
unit = lambda key: next(h for h in hdr if h.startswith(key)).split("[")[1].rstrip("]")
scale = {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3, "byte": 1.0}
rd = col("dram__bytes_read.sum") * scale[unit("dram__bytes_read.sum")]
wr = col("dram__bytes_write.sum") * scale[unit("dram__bytes_write.sum")]
sass = json.load(open(os.path.join(ROOT, "profiles", "sass_summary.json")))
out = {"k_render_dram_bytes_per_launch": int(rd + wr), "dram_bytes_read": int(rd), "dram_bytes_write": int(wr),
       "kernel": "k_render<(bool)0>", "kernel_sass_sha256": sass["k_render<(bool)0>"]["sha256"],
       "kernel_us_under_ncu": col("gpu__time_duration.sum"),
       "source": "%s (ncu --set full --clock-control none, dram__bytes_read.sum + dram__bytes_write.sum, one launch = 30 frames)"
                 % os.path.relpath(path, ROOT)}
json.dump(out, open(os.path.join(ROOT, "profiles", "traffic_%s.json" % cfg), "w"), indent=1)
print(out)
