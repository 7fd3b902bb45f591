#!/bin/bash
# k_inflate variants timed on real (Huffman-only) and zlib-6 streams + whole-pipeline ingest rate
for V in "$@"; do
  ORB_NVCC_EXTRA="$V" python visual-odometry-gpu_b200/build.py --force > /dev/null 2>&1 || { echo "build failed: $V"; continue; }
  python tools/inflate_debug.py | grep -c "first mismatch" | sed "s/^/$V mismatching streams: /"
  for KIND in real synth; do
    ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"k_inflate" -c 2 --csv --log-file gpurun_out/iv.csv python tools/ingest_probe.py 256 $KIND 256 > /dev/null 2>&1
    python - <<XX
import csv
rows=list(csv.reader(open("gpurun_out/iv.csv")))
h=[i for i,r in enumerate(rows) if "Kernel Name" in r][0]
v=rows[h].index("Metric Value")
print("$V", "$KIND", "k_inflate 256 streams ns:", [r[v] for r in rows[h+1:h+3]])
XX
  done
  python tools/ingest_probe.py 1024 real 0 | sed "s/^/$V /"; python tools/ingest_probe.py 1024 synth 0 | sed "s/^/$V /"
done
python visual-odometry-gpu_b200/build.py --force > /dev/null 2>&1
