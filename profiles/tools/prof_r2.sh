#!/bin/bash
set -x
B="python bench.py --steps 1 --warmup 1 --sim-steps 8 --no-cpu-baseline --no-secondary --no-sdf-query --strong-total 0"
$B > gpurun_out/plain_r2.json 2> gpurun_out/plain_r2.err || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/launches_r2.csv $B > gpurun_out/ncu_l_r2.log 2>&1
for K in contacts_kernel dyn_forward_kernel dyn_backward_kernel contact_geometry_bwd_kernel; do
  S=$(python profiles/tools/pick_launch.py gpurun_out/launches_r2.csv $K)
  echo "$K heaviest launch index $S" >> gpurun_out/picked_r2.txt
  ncu --set full --clock-control none --import-source on -k regex:$K -s $S -c 1 -o gpurun_out/prof_r2_$K -f $B > gpurun_out/ncu_${K}_r2.log 2>&1
done
python profiles/tools/c4prof.py 256 8 > gpurun_out/c4_plain.log 2>&1 || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/launches_r2_c4.csv python profiles/tools/c4prof.py 256 8 > gpurun_out/ncu_l_c4.log 2>&1
S=$(python profiles/tools/pick_launch.py gpurun_out/launches_r2_c4.csv contacts_kernel)
echo "c4 contacts_kernel heaviest launch index $S" >> gpurun_out/picked_r2.txt
ncu --set full --clock-control none --import-source on -k regex:contacts_kernel -s $S -c 1 -o gpurun_out/prof_r2_contacts_c4 -f python profiles/tools/c4prof.py 256 8 > gpurun_out/ncu_c4_r2.log 2>&1
tail -n 12 gpurun_out/c4_plain.log
cat gpurun_out/picked_r2.txt
ls -la gpurun_out/*.ncu-rep
