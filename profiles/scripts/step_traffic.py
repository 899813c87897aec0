"""Post-processes an ncu CSV of one GAN training step (every convolution kernel launch) into
profiles/r2_igemm_step_traffic.{txt,json}: DRAM read+write bytes, duration and tensor-pipe activity per launch.

capture (GPU box):
  ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active \
      --clock-control none -k regex:'igemm|thin_' -s <launches of 2 warm-up steps> -c <launches of one step> --csv \
      --log-file gpurun_out/r2_step_traffic.csv python profiles/scripts/prof_step.py
then here:  python profiles/scripts/step_traffic.py gpurun_out/r2_step_traffic.csv"""
import csv, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
rows = list(csv.reader(l for l in open(sys.argv[1]) if l.startswith('"')))
hdr = rows[0]; idx = {h: i for i, h in enumerate(hdr)}
per = {}
for r in rows[1:]:
    if len(r) < len(hdr): continue
    key = r[idx["ID"]]
    d = per.setdefault(key, {"name": r[idx["Kernel Name"]]})
    val = float(r[idx["Metric Value"]].replace(",", ""))
    unit = r[idx["Metric Unit"]]
    name = r[idx["Metric Name"]]
    scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "usecond": 1.0, "msecond": 1e3, "nsecond": 1e-3, "ns": 1e-3, "us": 1.0, "ms": 1e3, "second": 1e6, "%": 1.0}.get(unit, 1.0)
    d[name] = val * scale
launches = [d for _, d in sorted(per.items(), key=lambda kv: int(kv[0]))]
tot_b = sum(d.get("dram__bytes_read.sum", 0) + d.get("dram__bytes_write.sum", 0) for d in launches)
tot_t = sum(d.get("gpu__time_duration.sum", 0) for d in launches)
tp = sum(d.get("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", 0) * d.get("gpu__time_duration.sum", 0) for d in launches) / max(tot_t, 1e-9)
out_txt = os.path.join(ROOT, "profiles", "r2_igemm_step_traffic.txt")
with open(out_txt, "w") as f:
    f.write(f"# {len(launches)} convolution launches of one GAN step (batch 64): DRAM {tot_b / 1e9:.3f} GB, summed duration {tot_t / 1e3:.3f} ms (ncu, serialised, cold cache), "
            f"time-weighted tensor-pipe activity {tp:.1f} %\n# id  us  DRAM-read MB  DRAM-write MB  tensor%  kernel\n")
    for i, d in enumerate(launches):
        f.write(f"{i:3d} {d.get('gpu__time_duration.sum', 0):8.1f} {d.get('dram__bytes_read.sum', 0) / 1e6:9.1f} {d.get('dram__bytes_write.sum', 0) / 1e6:9.1f} "
                f"{d.get('sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 0):6.1f}  {d['name'][:70]}\n")
json.dump({"launches": len(launches), "dram_bytes_per_step": tot_b, "dram_bytes_per_launch": tot_b / max(len(launches), 1),
           "ncu_time_us_per_step": tot_t, "tensor_pipe_active_pct_time_weighted": tp,
           "source": "ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active... -k regex:igemm|thin_ over one step of profiles/scripts/prof_step.py"},
          open(os.path.join(ROOT, "profiles", "r2_igemm_step_traffic.json"), "w"), indent=1)
print(open(out_txt).readline())
