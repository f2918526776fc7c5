"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per kernel launches, total / average time and
share of the listed time.  usage: python tools/launch_list_summary.py gpurun_out/launches.csv [only-this-repo]"""
import collections, csv, re, sys

rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 14 and r[12] == "gpu__time_duration.sum"]
ours = re.compile(r"^(nerf::)?(mlp_fused_kernel|wgrad_pair_kernel|wgrad_kernel|head_grad_kernel|pack_|adam_|composite_|mse_loss|sample_z|raygen_|draw_pixels|train_batch|encode_kernel)")
agg = collections.OrderedDict()
prev = ""
for r in rows:
    name = re.sub(r"\(.*", "", r[4])
    name = re.sub(r"^void ", "", name).replace("nerf::", "")
    if len(sys.argv) > 2 and not ours.match(name):
        continue
    name = {"mlp_fused_kernel<0, 1>": "mlp_fused_kernel<FWD, train> [train step]", "mlp_fused_kernel<0, 0>": "mlp_fused_kernel<FWD, inference> [render frames]",
            "mlp_fused_kernel<1, 0>": "mlp_fused_kernel<BWD> [dgrad]"}.get(name, name)
    prev = re.sub(r" \[.*", "", name)
    a = agg.setdefault(name, [0, 0.0])
    a[0] += 1
    a[1] += float(r[14].replace(",", "")) / 1e3
total = sum(a[1] for a in agg.values())
print("kernel,launches,total_us,avg_us,share_of_listed_time")
for k, (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{k},{n},{us:.1f},{us / n:.2f},{us / total:.4f}")
