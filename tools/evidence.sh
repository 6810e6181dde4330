#!/bin/bash
# Round evidence on one B200: bench lines of every configuration, the reference arm, the ncu
# launch list of the bench command and one --set full capture of each kernel per configuration.
#   bash tools/evidence.sh r02_v10
tag=${1:-rXX}
out=gpurun_out
mkdir -p $out
for c in c3 c2 c4 c5; do
  timeout 900 python bench.py --config $c --steps 5 --warmup 3 > $out/${tag}_bench_$c.json 2> $out/${tag}_bench_$c.err
done
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $out/${tag}_bench_c3_reference_arm.json 2>> $out/${tag}_bench_c3.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $out/${tag}_launches_c3.csv \
  python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > $out/${tag}_launches_c3.log 2>&1
for c in c3 c2 c4 c5; do
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:k3d_s -s 2 -c 2 -o $out/${tag}_$c -f \
    python tools/kbench.py $c 1 > $out/${tag}_ncu_$c.log 2>&1
done
ls -la $out | tail -20
