# usage: bash tools/profile_round.sh <tag>   -- bench line, ncu launch list and one full capture of the scan kernels
tag=$1
A="--steps 5 --warmup 3 --cpu-sample 500"
python bench.py $A > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/${tag}_launches.csv python bench.py $A > gpurun_out/${tag}_ncu_list.log 2>&1 && \
ncu --set full --import-source on --clock-control none -k regex:tc_scan_kernel --launch-skip 20 -c 2 -o gpurun_out/${tag}_tc_scan -f python bench.py $A > gpurun_out/${tag}_ncu_full.log 2>&1
echo "rc=$?"; tail -c 600 gpurun_out/${tag}_bench.json
