"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel totals per encode+decode pass and
shares (the per-launch times are cold-cache and serialised: compare SHARES).  Also writes a reduced CSV (kernel, grid,
block, ns) next to the summary.   python tools/launch_summary.py gpurun_out/launches.csv profiles/r01_launches_v8"""
import collections, csv, re, sys
src, dst = sys.argv[1], sys.argv[2]
rows = []
with open(src, newline="") as fh:
    rd = csv.reader(l for l in fh if l.startswith('"'))
    hdr = next(rd)
    ix = {h: i for i, h in enumerate(hdr)}
    for r in rd:
        if r[0] == "ID" or len(r) != len(hdr):
            continue
        rows.append((r[ix["Kernel Name"]], r[ix["Grid Size"]], r[ix["Block Size"]], float(r[ix["Metric Value"]])))
short = lambda n: re.sub(r"\(.*", "", n)[:100]
passes = sum(1 for r in rows if "conv_rowlane_kernel<1, 16, 27>" in r[0]) or 1
tot = collections.defaultdict(lambda: [0.0, 0])
for n, g, b, ns in rows:
    tot[short(n)][0] += ns; tot[short(n)][1] += 1
total = sum(v[0] for v in tot.values())
with open(dst + ".csv", "w", newline="") as fh:
    w = csv.writer(fh); w.writerow(["kernel", "grid", "block", "gpu__time_duration.sum [ns]"])
    for n, g, b, ns in rows:
        w.writerow([short(n), g, b, int(ns)])
k3 = sum(v[0] for n, v in tot.items() if "conv_k3" in n)
probe = [ns for n, g, b, ns in rows if "conv_k3_octet_h2_kernel<16, 16" in n and ns > 150e3]
with open(dst + ".summary.txt", "w") as fh:
    fh.write(f"ncu --metrics gpu__time_duration.sum --clock-control none : python bench.py --steps 1 --warmup 3 --no-cpu-baseline "
             f"({passes} encode+decode passes captured, {len(rows)} launches; cold-cache, serialised -- compare SHARES)\n")
    fh.write(f"total {total / 1e6:.1f} ms over {passes} passes = {total / passes / 1e6:.2f} ms/pass\n")
    fh.write(f"k=3 convolution kernels: {k3 / passes / 1e6:.2f} ms/pass = {100 * k3 / total:.1f}% of kernel time; decoder.conv2 launch alone "
             f"(conv_k3_octet_h2_kernel<16,16> @1.69M rows): {sum(probe) / max(len(probe), 1) / 1e6:.3f} ms = "
             f"{100 * sum(probe) / total:.1f}%\n")
    for n, (ns, c) in sorted(tot.items(), key=lambda kv: -kv[1][0])[:70]:
        fh.write(f"{ns / passes / 1e6:8.3f} ms/pass {100 * ns / total:5.1f}% x{c / passes:6.1f}  {n}\n")
print(open(dst + ".summary.txt").read()[:3000])
