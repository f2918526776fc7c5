"""Summarise an .ncu-rep here (no GPU): key raw metrics per launch + hottest SASS lines per kernel.
usage: python tools/ncu_summary.py gpurun_out/prof.ncu-rep [launch_index_for_source]"""
import collections, csv, io, subprocess, sys

rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
idx = {h: i for i, h in enumerate(hdr)}
want = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__m_xbar2l1tex_read_bytes.sum", "smsp__inst_executed.sum",
        "sm__cycles_elapsed.max", "launch__grid_size", "launch__registers_per_thread"]
want = [w for w in want if w in idx]
print(",".join(want))
print(",".join(units[idx[w]] for w in want))
for r in rows[2:]:
    print(",".join(r[idx[w]] for w in want))
if len(sys.argv) > 2:
    k = int(sys.argv[2])
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--launch-skip", str(k), "--launch-count", "1"],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(src)))
    print("#", rows[0][:2])
    hdr = rows[1]; idx = {h: i for i, h in enumerate(hdr)}
    seen, data = set(), []
    for r in rows[2:]:
        if len(r) <= idx["# Samples"] + 3 or r[idx["Address"]] in seen:
            continue
        seen.add(r[idx["Address"]]); data.append(r)
    def I(r, k):
        try: return int(float(r[idx[k]] or 0))
        except ValueError: return 0
    st = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    print("# total samples", sum(I(r, "# Samples") for r in data), "warp-instructions", sum(I(r, "Instructions Executed") for r in data))
    cls = collections.defaultdict(lambda: [0, 0])
    for r in data:
        c = cls[I(r, "Instructions Executed")]; c[0] += I(r, "# Samples"); c[1] += 1
    for kk, v in sorted(cls.items(), key=lambda kv: -kv[1][0])[:10]:
        print("# exec-count", kk, "samples", v[0], "sass-instrs", v[1])
    for r in sorted(data, key=lambda r: -I(r, "# Samples"))[:int(sys.argv[3]) if len(sys.argv) > 3 else 24]:
        ss = {h.replace("stall_", ""): I(r, h) for h in st if I(r, h) > 20}
        print(r[idx["Address"]][-5:], I(r, "Instructions Executed"), I(r, "# Samples"), r[idx["Source"]][:60], ss)
