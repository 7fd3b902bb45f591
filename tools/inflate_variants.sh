#!/bin/bash
# time k_inflate variants on the GPU box: tools/inflate_variants.sh "<flags A>" "<flags B>" ...
for V in "$@"; do
  ORB_NVCC_EXTRA="$V" python visual-odometry-gpu_b200/build.py --force > /dev/null 2>&1 || { echo "build failed: $V"; continue; }
  python tools/inflate_debug.py | grep -c "first mismatch" | sed "s/^/$V mismatching streams: /"
  ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"k_inflate" -c 2 --csv --log-file gpurun_out/iv.csv python tools/ingest_probe.py 256 real 256 > /dev/null 2>&1
  python - <<XX
import csv
rows=list(csv.reader(open("gpurun_out/iv.csv")))
h=[i for i,r in enumerate(rows) if "Kernel Name" in r][0]
v=rows[h].index("Metric Value")
print("$V", "k_inflate 256 KITTI streams:", [r[v] for r in rows[h+1:h+3]])
XX
done
python visual-odometry-gpu_b200/build.py --force > /dev/null 2>&1
