"""Summarise an ncu launch list (`ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --csv`) of
`bench.py --steps 2 --warmup 1 --no-graph --no-also --no-cpu-baseline`: per-kernel totals of the LAST full train step (time, share,
DRAM bytes, GB/s) and the per-launch list of the igemm kernel with its template variant and grid.

    python tools/launch_table.py gpurun_out/launches.csv [launches_per_step]
"""
import collections
import csv
import re
import sys


def load(path):
    rows = []
    with open(path) as f:
        for line in f:
            if line.startswith('"ID"'):
                hdr = next(csv.reader([line]))
                break
        rd = csv.reader(f)
        for r in rd:
            if len(r) == len(hdr):
                rows.append(dict(zip(hdr, r)))
    launches = collections.OrderedDict()
    for r in rows:
        d = launches.setdefault(int(r["ID"]), {"name": r["Kernel Name"], "grid": r["Grid Size"], "block": r["Block Size"]})
        v = float(r["Metric Value"].replace(",", ""))
        unit = r["Metric Unit"]
        if r["Metric Name"] == "gpu__time_duration.sum":
            d["us"] = v * {"ns": 1e-3, "us": 1.0, "usecond": 1.0, "nsecond": 1e-3, "ms": 1e3, "msecond": 1e3}.get(unit, 1e-3)
        else:
            d[r["Metric Name"]] = v * {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1.0)
    return list(launches.values())


def short(name):
    n = name.replace("void ", "").replace("b200::", "").replace("(bool)", "")
    m = re.match(r"igemm_kernel<Epi<([^>]*)>, *(\d)(?:, *(\d))?>", n)
    if m:
        e = [x.strip() for x in m.group(1).split(",")]
        tag = "".join(("S" if e[0] == "1" else "-", "B" if e[1] == "1" else "-", {"0": "-", "1": "L", "2": "R"}[e[2].strip()], "G" if e[3] == "1" else "-",
                       "N" if e[4] == "1" else "-"))
        return f"igemm[{tag}{' pair' if m.group(2) == '1' else ''}{' win' if m.group(3) == '1' else (' win4' if m.group(3) == '4' else '')}]"
    return n.split("<")[0].split("(")[0]


def main():
    L = load(sys.argv[1])
    ours = [x for x in L if "b200::" in x["name"] or "igemm" in x["name"]]
    # the last COMPLETE step: between the last two mask_prepare launches (the capture may end inside a step)
    starts = [i for i, x in enumerate(ours) if "mask_prepare" in x["name"]]
    step = ours[starts[-2]:starts[-1]] if len(starts) >= 2 else ours
    tot = sum(x.get("us", 0) for x in step)
    agg = collections.OrderedDict()
    for x in step:
        k = short(x["name"]).split("[")[0]
        a = agg.setdefault(k, [0, 0.0, 0.0])
        a[0] += 1
        a[1] += x.get("us", 0)
        a[2] += x.get("dram__bytes_read.sum", 0) + x.get("dram__bytes_write.sum", 0)
    print(f"last step: {len(step)} launches, {tot:.0f} us of kernel time, {sum(a[2] for a in agg.values()) / 1e9:.2f} GB DRAM traffic")
    for k, (n, us, by) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"  {k:28s} n={n:3d} {us:8.1f} us {100 * us / tot:5.1f}%  {by / 1e6:8.1f} MB  {by / max(us, 1e-9) / 1e3:7.1f} GB/s")
    print("igemm launches of the step (launch order):")
    i = 0
    for x in step:
        if "igemm" in x["name"]:
            by = x.get("dram__bytes_read.sum", 0) + x.get("dram__bytes_write.sum", 0)
            print(f"  {i:2d} {short(x['name']):26s} grid {x['grid']:12s} {x.get('us', 0):7.1f} us  {by / 1e6:7.1f} MB")
            i += 1


if __name__ == "__main__":
    main()
