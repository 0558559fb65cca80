"""Per-kernel totals of an ncu launch list (--metrics gpu__time_duration.sum --csv), and -- for a PPO update -- the launches of
one Adam step by stream (a step ends with k_clip_adam).    python tools/launch_summary.py launches.csv [out.csv]"""
import collections
import csv
import re
import sys


def main(path, out=None):
    rows = list(csv.reader(open(path, errors="ignore")))
    hdr = None
    cnt, tot, seq = collections.Counter(), collections.Counter(), []
    for r in rows:
        if "Kernel Name" in r:
            hdr = r
            continue
        if hdr and len(r) == len(hdr):
            d = dict(zip(hdr, r))
            n = re.sub(r"<.*", "", re.sub(r"\(.*", "", d["Kernel Name"]).replace("void ", "").replace("enero::", ""))
            try:
                v = float(d["Metric Value"].replace(",", ""))
            except ValueError:
                continue
            if d.get("Metric Unit") == "ns":
                v /= 1000.0
            cnt[n] += 1
            tot[n] += v
            seq.append((n, v, d["Stream"]))
    total = sum(tot.values())
    lines = [("kernel", "launches", "total_us", "avg_us", "share_pct")]
    for n, c in sorted(cnt.items(), key=lambda kv: -tot[kv[0]]):
        lines.append((n, c, round(tot[n], 1), round(tot[n] / c, 2), round(100 * tot[n] / total, 1)))
    ends = [i for i, s in enumerate(seq) if s[0] == "k_clip_adam"]
    if len(ends) >= 2:
        a, b = ends[0] + 1, ends[1] + 1
        per = collections.defaultdict(lambda: [0, 0.0])
        for n, v, s in seq[a:b]:
            per[s][0] += 1
            per[s][1] += v
        lines.append(("-- one Adam step: %d launches, %.1f us summed" % (b - a, sum(v for _, v, _ in seq[a:b])), "", "", "", ""))
        for s, (c, v) in sorted(per.items(), key=lambda kv: -kv[1][1]):
            lines.append(("stream %s" % s, c, round(v, 1), "", ""))
        longest = max(per, key=lambda s: per[s][1])
        agg = collections.defaultdict(lambda: [0, 0.0])
        for n, v, s in seq[a:b]:
            if s == longest:
                agg[n][0] += 1
                agg[n][1] += v
        for n, (c, v) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            lines.append(("  stream %s: %s" % (longest, n), c, round(v, 1), "", ""))
    w = csv.writer(open(out, "w", newline="") if out else sys.stdout)
    w.writerows(lines)


if __name__ == "__main__":
    main(*sys.argv[1:3])
