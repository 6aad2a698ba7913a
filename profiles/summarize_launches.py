"""Summarise an ncu launch list (--metrics gpu__time_duration.sum --csv) of bench.py: the launches of
the LAST training step (from the last first-forward kernel to the last zero_rows / adam launch) and
the per-kernel shares.   python profiles/summarize_launches.py LIST.csv "header comment" > LIST.txt"""
import collections
import csv
import re
import sys

path, note = sys.argv[1], (sys.argv[2] if len(sys.argv) > 2 else "")
rows = list(csv.reader(l for l in open(path) if not l.startswith("==")))
hdr = rows[0]
ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")


def short(k):
    k = re.sub(r"\(.*", "", k)
    k = re.sub(r"^void ", "", k)
    k = k.replace("lgcn::", "").replace("(int)", "").replace("(bool)", "")
    return k.strip()


launches = [(short(r[ki]), float(r[vi].replace(",", "")) / 1e6) for r in rows[1:]]
mine = [i for i, (k, _) in enumerate(launches) if k.startswith(("spmm_", "bpr_", "adam", "zero_rows", "ftc::", "fusion_"))]
# steps end with the last zero_rows (plain) or the last adam_kernel (fusion); walk back to the previous one
ends = [i for i in mine if launches[i][0].startswith("zero_rows") or launches[i][0].startswith("adam_kernel")]
last = ends[-1]
prev = max([i for i in ends if i < last and launches[i + 1][0] != launches[last][0]
            and not launches[i + 1][0].startswith("adam_kernel")] or [-1])
step = [launches[i] for i in mine if prev < i <= last]
print(f"# {note}")
print(f"# raw list: {path.split('/')[-1]}; per-launch times are cold-cache and serialised under ncu -- compare SHARES.")
print(f"# last training step of the run = {len(step)} launches:")
print("#   launch  ms      kernel")
for i, (k, ms) in enumerate(step):
    print(f"#   {i:4d}  {ms:7.3f}  {k}")
tot = sum(ms for _, ms in step)
print(f"# total {tot:.3f} ms")
agg = collections.OrderedDict()
for k, ms in step:
    a = agg.setdefault(k, [0, 0.0])
    a[0] += 1
    a[1] += ms
print("# shares:")
for k, (n, ms) in sorted(agg.items(), key=lambda x: -x[1][1]):
    print(f"#   {ms / tot * 100:5.1f} %  {ms:8.3f} ms  x{n:<3d} {k}")
