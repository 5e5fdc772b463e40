O=gpurun_out
SHORT="python bench.py --steps 10 --warmup 3 --no-ramp --no-cpu --no-ransac --no-hamming --no-cfg5"
$SHORT > $O/plain_k3.log 2>&1 || exit 1
ncu --set full --clock-control none --import-source on -k regex:l2_finish_kernel -s 6 -c 1 -o $O/prof_l2_finish_kernel_k2j -f $SHORT > $O/ncu_k3.log 2>&1
echo "ncu exit $?"
