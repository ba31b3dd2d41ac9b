mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r1z_small_launches.csv python tools/ncu_small.py > gpurun_out/r1z_small.log 2>&1
python - <<PY
import csv
rows = [r for r in csv.reader(open("gpurun_out/r1z_small_launches.csv")) if len(r) > 10]
hdr = rows[0]
ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
for r in rows[1:][-40:]:
    print("%-70s %s" % (r[ki][:70], r[vi]))
PY
