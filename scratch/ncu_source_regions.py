"""Source-level reading of an `ncu --set full --import-source on` capture.

    ncu -i capture.ncu-rep --page source --csv --print-source cuda,sass > src.csv
    python scratch/ncu_source_regions.py src.csv [top_n]

Prints (1) per SASS instruction, de-duplicated by address (the cuda,sass view lists an inlined instruction under every
line of its inline stack): the hot loop = the instructions with the largest execution count, its share of the kernel's
instructions and of its warp samples, and the stall reasons of its samples; (2) the source lines with the most samples
(a line's figures include the instructions inlined into it)."""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
top_n = int(sys.argv[2]) if len(sys.argv) > 2 else 25
hdr, cur_file = None, None
sass = {}                                    # address -> (text, samples, executed, stalls)
lines = collections.OrderedDict()            # (file, line) -> [samples, executed, thread-instructions, text]
for r in rows:
    if not r:
        continue
    if r[0] in ("File Path", "File Name"):
        cur_file = r[1].split("/")[-1]
        continue
    if r[0] == "Line No":
        hdr = r
        continue
    if hdr is None or len(r) < len(hdr):
        continue
    off = len(r) - len(hdr)                  # source text with unescaped quotes splits into extra columns
    col = lambda name: r[hdr.index(name) + off]
    try:
        s, i, t = int(col("# Samples") or 0), int(col("Instructions Executed") or 0), int(col("Thread Instructions Executed") or 0)
    except ValueError:
        continue
    if r[0] == "":
        st = {k: int(col(k) or 0) for k in hdr if k.startswith("stall_") and "Not Issued" not in k}
        sass[r[2]] = (r[3].strip(), s, i, st)
    else:
        try:
            a = lines.setdefault((cur_file, int(r[0])), [0, 0, 0, r[1].strip()])
        except ValueError:
            continue
        a[0] += s; a[1] += i; a[2] += t
tot_s = sum(v[1] for v in sass.values())
tot_i = sum(v[2] for v in sass.values())
print(f"{len(sass)} SASS instructions, {tot_i:.4g} warp instructions executed, {tot_s} warp samples")
cnt = collections.Counter(v[2] for v in sass.values())
hot = max((c for c in cnt if cnt[c] >= 32), default=None)
if hot:
    loop = [v for v in sass.values() if v[2] == hot]
    ls, li = sum(v[1] for v in loop), sum(v[2] for v in loop)
    agg = collections.Counter()
    for v in loop:
        agg.update(v[3])
    n = sum(agg.values()) or 1
    print(f"hot loop: {len(loop)} instructions x {hot} executions = {100 * li / tot_i:.1f} % of the instructions, "
          f"{100 * ls / tot_s:.1f} % of the samples")
    print("  its samples by stall reason:", ", ".join(f"{k[6:]} {100 * v / n:.0f} %" for k, v in agg.most_common(6)))
    rest_s, rest_i = tot_s - ls, tot_i - li
    print(f"everything else: {100 * rest_i / tot_i:.1f} % of the instructions in {100 * rest_s / tot_s:.1f} % of the samples")
print(f"\ntop {top_n} source lines by samples (share of samples / of instructions, lanes per instruction):")
ts = sum(v[0] for v in lines.values()) or 1
ti = sum(v[1] for v in lines.values()) or 1
for (f, ln), (s, i, t, src) in sorted(lines.items(), key=lambda kv: -kv[1][0])[:top_n]:
    print(f"  {f}:{ln:<4d} {100 * s / ts:5.2f} % {100 * i / ti:5.2f} % {t / max(i, 1):5.1f} | {src[:100]}")
