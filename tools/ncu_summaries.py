"""Turns the raw ncu CSV logs of tools/gpu_calls/gpu_round23.sh into the summaries committed under profiles/:
  launches: per-kernel share of ONE eval forward (the last forward in the log), from gpu__time_duration.sum
  conv traffic: DRAM bytes / duration / tensor-pipe activity summed over the conv_gemm_kernel launches of one forward
usage: python tools/ncu_summaries.py gpurun_out/launches_eval_step.csv gpurun_out/conv_traffic.csv profiles/r01"""
import csv
import io
import json
import re
import sys
from collections import defaultdict


def rows_of(path):
    text = open(path).read()
    start = text.index('"ID"')
    return list(csv.DictReader(io.StringIO(text[start:])))


def short(name):
    name = re.sub(r"\(anonymous namespace\)::|dpft::|void |<unnamed>::", "", name)
    return re.sub(r"\(.*$", "", name)[:110]


def launches(path, out_prefix):
    rows = rows_of(path)
    per_id = {}
    for r in rows:
        if r["Metric Name"] == "gpu__time_duration.sum":
            per_id[int(r["ID"])] = (short(r["Kernel Name"]), float(r["Metric Value"]) / 1e3, r["Grid Size"], r["Block Size"])
    ids = sorted(per_id)
    # one forward = from the last stem kernel of the camera view (first launch of a forward) to the end of the log
    firsts = [i for i in ids if per_id[i][0].startswith("stem_tc_kernel<3")]
    begin = firsts[-1] if firsts else ids[0]
    step = [per_id[i] for i in ids if i >= begin]
    agg = defaultdict(lambda: [0.0, 0])
    for name, us, _, _ in step:
        agg[name][0] += us
        agg[name][1] += 1
    total = sum(v[0] for v in agg.values())
    with open(out_prefix + "_ncu_launches_eval_step.txt", "w") as f:
        f.write("ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off python tools/one_forward.py 2   (eager, serial streams; "
                "cold-cache serialised launches: shares, not absolute times)\n")
        f.write(f"one eval forward of the bench workload: {len(step)} launches, {total / 1e3:.3f} ms summed\n\n")
        f.write(f"{'us':>10} {'share':>7} {'launches':>8}  kernel\n")
        for name, (us, n) in sorted(agg.items(), key=lambda kv: -kv[1][0]):
            f.write(f"{us:10.1f} {100 * us / total:6.1f}% {n:8d}  {name}\n")
        f.write("\nlaunch list (in order):\n")
        for name, us, grid, block in step:
            f.write(f"{us:9.2f} us  grid {grid:>16} block {block:>14}  {name}\n")
    return total, len(step)


def conv_traffic(path, out_prefix):
    rows = rows_of(path)
    per_id = defaultdict(dict)
    for r in rows:
        per_id[int(r["ID"])][r["Metric Name"]] = (float(r["Metric Value"].replace(",", "")), r["Metric Unit"])
        per_id[int(r["ID"])]["name"] = short(r["Kernel Name"])
        per_id[int(r["ID"])]["grid"] = r["Grid Size"]
    ids = sorted(per_id)
    n = len(ids) // 2 if len(ids) > 300 else 0    # two forwards profiled: keep the second; one (cudaProfilerStart range): keep all
    sel = [per_id[i] for i in ids[n:]]

    def to_bytes(v):
        val, unit = v
        return val * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[unit]

    def to_us(v):
        val, unit = v
        return val * {"ns": 1e-3, "us": 1, "ms": 1e3, "usecond": 1, "nsecond": 1e-3, "msecond": 1e3}[unit]

    rd = sum(to_bytes(s["dram__bytes_read.sum"]) for s in sel)
    wr = sum(to_bytes(s["dram__bytes_write.sum"]) for s in sel)
    us = sum(to_us(s["gpu__time_duration.sum"]) for s in sel)
    key = "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"
    tw = sum(s[key][0] * to_us(s["gpu__time_duration.sum"]) for s in sel) / us
    summary = {"source": "ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum," + key +
                         " --clock-control none --profile-from-start off -k 'regex:conv_gemm|conv3x3_halo|conv_expand' python tools/one_forward.py (one eager forward, serial streams)",
               "launches": len(sel), "dram_read_bytes": rd, "dram_write_bytes": wr, "dram_bytes": rd + wr,
               "sum_duration_us_under_ncu": us, "tensor_pipe_active_pct_time_weighted": tw}
    with open(out_prefix + "_ncu_conv_traffic.json", "w") as f:
        json.dump(summary, f, indent=1)
    with open(out_prefix + "_ncu_conv_launches.txt", "w") as f:
        f.write(json.dumps(summary) + "\n\n")
        f.write(f"{'us':>9} {'read MB':>9} {'write MB':>9} {'tensor%':>8}  grid  kernel\n")
        for s in sel:
            f.write(f"{to_us(s['gpu__time_duration.sum']):9.2f} {to_bytes(s['dram__bytes_read.sum']) / 1e6:9.2f} "
                    f"{to_bytes(s['dram__bytes_write.sum']) / 1e6:9.2f} {s[key][0]:8.2f}  {s['grid']}  {s['name']}\n")
    return summary


if __name__ == "__main__":
    l, c, prefix = sys.argv[1:4]
    print(launches(l, prefix))
    print(conv_traffic(c, prefix))
