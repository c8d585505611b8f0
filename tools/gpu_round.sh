#!/bin/bash
# One GPU-box visit: parity tests, smoke, default bench + reference arm, launch list, and full ncu
# captures of the top kernels (each only after the plain command has exited 0).  Outputs under
# gpurun_out/<tag>_*; tools/ncu_summary.py / tools/make_traffic.py / tools/launch_share.py turn them into
# the committed profiles/ files.
#   gpurun --timeout 1800 -- 'bash tools/gpu_round.sh r02z "k_region_stats k_guided_ab ..." "k_slic_assign k_slic_features"'
tag=${1:-r02}
kernels=${2:-}
slic_kernels=${3:-}
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/${tag}_pytest.log
tail -3 gpurun_out/${tag}_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${tag}_smoke.log 2>&1
echo "smoke rc=$?"
python bench.py --impl reference > gpurun_out/${tag}_bench_reference_arm.json 2> gpurun_out/${tag}_bench_reference_arm.err
echo "reference arm rc=$?"
python bench.py > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err
rc=$?
echo "bench rc=$rc"
python tools/kernel_times.py --tag ${tag} --top 40 > gpurun_out/${tag}_kernel_times.log 2>&1
python tools/kernel_times.py --tag ${tag}_slic --slic --top 16 >> gpurun_out/${tag}_kernel_times.log 2>&1
python tools/latency_single.py > gpurun_out/${tag}_latency_single_image.txt 2>&1
head -30 gpurun_out/${tag}_kernel_times.log
if [ $rc -eq 0 ]; then
  ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv \
      --log-file gpurun_out/${tag}_launches.csv env GG_BENCH_NO_EXTRAS=1 GG_BENCH_NO_SWEEP=1 python bench.py --steps 2 --warmup 1 --no-cpu-baseline \
      > gpurun_out/${tag}_ncu_list.log 2>&1
  for k in $kernels; do
    GG_SUBBATCH=1 ncu --set full --clock-control none --import-source on -k regex:$k -s 3 -c 1 -f \
        -o gpurun_out/${tag}_$k python tools/kernel_times.py --steps 2 > gpurun_out/${tag}_ncu_$k.log 2>&1
  done
  for k in $slic_kernels; do
    GG_SUBBATCH=1 ncu --set full --clock-control none --import-source on -k regex:$k -s 1 -c 1 -f \
        -o gpurun_out/${tag}_$k python tools/kernel_times.py --steps 2 --slic > gpurun_out/${tag}_ncu_$k.log 2>&1
  done
fi
