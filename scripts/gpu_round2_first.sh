#!/bin/bash
# First GPU pass of the next round: (1) the tests that have not run on hardware yet, (2) the regular GPU suite,
# (3) the three bench workloads on one GPU, (4) launch lists + one --set full capture of the two new roofline kernels.
#   gpurun --timeout 2400 -- 'bash scripts/gpu_round2_first.sh'
#   gpurun --gpus 2 --timeout 1500 -- 'bash scripts/gpu_round2_first.sh mgpu'
set -u
mkdir -p gpurun_out
t0=$SECONDS
export UG4B200_PENDING_GPU_TESTS=1
if [ "${1:-}" = "mgpu" ]; then
  timeout 1200 python -m pytest tests/test_multi_gpu.py -m gpu -x -q 2>&1 | tail -15 | tee gpurun_out/pending_mgpu.log
  NG=$(nvidia-smi -L | wc -l)
  for w in convdiff elasticity; do
    timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29611 \
      bench.py --gpus $NG --workload $w --refs 6 --steps 5 --warmup 3 > gpurun_out/bench_${w}_n${NG}.json 2> gpurun_out/bench_${w}_n${NG}.err
    cut -c1-600 gpurun_out/bench_${w}_n${NG}.json; tail -2 gpurun_out/bench_${w}_n${NG}.err
  done
  echo "total: $((SECONDS-t0)) s"; exit 0
fi
# every GPU test, the opt-in ones included, without -x: the complete list of what is red on hardware
timeout 2400 python -m pytest tests -m gpu -q -rf 2>&1 | tail -40 | tee gpurun_out/pending_1gpu.log
unset UG4B200_PENDING_GPU_TESTS
for w in poisson convdiff elasticity; do
  extra=""; [ $w = elasticity ] && extra="--refs 6"
  timeout 900 python bench.py --workload $w $extra --steps 5 --warmup 3 > gpurun_out/bench_$w.json 2> gpurun_out/bench_$w.err
  cut -c1-900 gpurun_out/bench_$w.json; tail -2 gpurun_out/bench_$w.err
done
# device-resident BiCGStab (flag 32) against the host loop on the same workload
timeout 900 python bench.py --workload convdiff --flags 32 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_convdiff_devbicgstab.json 2> gpurun_out/bench_convdiff_devbicgstab.err
cut -c1-400 gpurun_out/bench_convdiff_devbicgstab.json; tail -2 gpurun_out/bench_convdiff_devbicgstab.err
for w in convdiff elasticity; do
  extra=""; [ $w = elasticity ] && extra="--refs 6"
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_$w.csv \
    python bench.py --workload $w $extra --steps 1 --warmup 1 --no-cpu-baseline > /dev/null 2>&1
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gs_color_kernel -c 2 -o gpurun_out/gs_color \
  python bench.py --workload convdiff --steps 1 --warmup 1 --no-cpu-baseline > /dev/null 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:spmvB_kernel -c 2 -o gpurun_out/spmvB \
  python bench.py --workload elasticity --refs 6 --steps 1 --warmup 1 --no-cpu-baseline > /dev/null 2>&1
echo "total: $((SECONDS-t0)) s"
