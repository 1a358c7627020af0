"""Summarise an ncu launch list (--metrics gpu__time_duration.sum --csv) of `bench.py --steps 1 --warmup 1`:
kernel time per kernel name over the LAST step (from the last-but-one make_records launch marker, i.e. the start of
the final recombination call, to the end).  usage: launch_summary.py launches.csv [marker-substring]"""
import collections, csv, re, sys

path = sys.argv[1]
marker = sys.argv[2] if len(sys.argv) > 2 else "make_records"
rows = [r for r in csv.reader(open(path, errors="replace")) if len(r) > 10 and r[0].isdigit()]
names = [r[4] for r in rows]
marks = [i for i, n in enumerate(names) if marker in n]
# bench.py runs: warm-up step(s), timed step(s) on device-resident inputs, then e2e step(s) from host buffers;
# take the launches between the last two markers' predecessors = one complete device-resident step
if len(marks) >= 3:
    lo, hi = marks[-3], marks[-2]
elif len(marks) == 2:
    lo, hi = marks[0], marks[1]
else:
    lo, hi = 0, len(rows)
sel = rows[lo:hi]
tot = collections.Counter(); cnt = collections.Counter()
for r in sel:
    n = re.sub(r"<.*", "", r[4].replace("<unnamed>::", "")).replace("void ", "").strip()
    n = re.sub(r"\(.*", "", n)
    tot[n] += float(r[-1]) / 1e3
    cnt[n] += 1
total = sum(tot.values())
print("# one step: %d launches, %.2f ms summed kernel time (cold-cache, serialised under ncu)" % (len(sel), total / 1e3))
print("%10s %7s %6s  kernel" % ("us", "share", "calls"))
for n, t in tot.most_common(28):
    print("%10.1f %6.1f%% %6d  %s" % (t, 100 * t / total, cnt[n], n[:90]))
