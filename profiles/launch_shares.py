"""Per-kernel totals of an ncu launch list (`--metrics gpu__time_duration.sum --csv`, profiles/capture.sh).

    python profiles/launch_shares.py gpurun_out/r01c_launches.csv > profiles/r01/launch_shares_r01.txt
"""
import collections
import csv
import re
import sys

rows = [r for r in csv.reader(line for line in open(sys.argv[1]) if line.startswith('"'))]
head, rows = rows[0], rows[1:]
k_name, k_val = head.index("Kernel Name"), head.index("Metric Value")
tot, cnt = collections.Counter(), collections.Counter()
for r in rows:
    name = re.sub(r"\(.*", "", r[k_name])[:52]
    tot[name] += float(r[k_val].replace(",", ""))
    cnt[name] += 1
all_ns = sum(tot.values())
print("# launch list of `python bench.py --steps 3 --warmup 3` under ncu (gpu__time_duration.sum, ns; cold-cache, serialised)")
print("# %-50s %8s %10s %9s %7s" % ("kernel", "launches", "total_ns", "mean_ns", "share"))
for name, ns in tot.most_common():
    print("%-52s %8d %10d %9d %7.3f" % (name, cnt[name], ns, ns / cnt[name], ns / all_ns))
