#!/bin/bash
# Round-2 GPU session.  usage (under gpurun): bash tools/gpu_r02.sh <tag> <stages...>
#   stages: quick (small parity tests) | pytest (all -m gpu) | smoke | bench:<workload>[:steps] | benchall |
#           launches:<workload> | ncu:<workload>:<kernel-regex>:<skip> | knock
TAG=$1; shift
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv > gpurun_out/${TAG}_gpu.txt 2>&1
nproc >> gpurun_out/${TAG}_gpu.txt; free -g | head -2 >> gpurun_out/${TAG}_gpu.txt
for ST in "$@"; do
  IFS=: read -r KIND A B C <<< "$ST"
  case $KIND in
    quick)
      timeout 900 python -m pytest tests -m gpu -x -q -k "not large_config and not taxol_full and not two_rank" > gpurun_out/${TAG}_quick.log 2>&1
      echo "quick exit $?" >> gpurun_out/${TAG}_quick.log; tail -15 gpurun_out/${TAG}_quick.log ;;
    pytest)
      timeout 2400 python -m pytest tests -m gpu -q --durations=8 > gpurun_out/${TAG}_pytest.log 2>&1
      echo "pytest exit $?" >> gpurun_out/${TAG}_pytest.log; tail -25 gpurun_out/${TAG}_pytest.log ;;
    smoke)
      timeout 300 python __graft_entry__.py smoke > gpurun_out/${TAG}_smoke.log 2>&1; tail -3 gpurun_out/${TAG}_smoke.log ;;
    bench)
      timeout 1500 python bench.py --workload $A --steps ${B:-5} --warmup 3 --others "" > gpurun_out/${TAG}_bench_${A}.json 2> gpurun_out/${TAG}_bench_${A}.err
      echo "bench $A exit $?"; python tools/bench_brief.py gpurun_out/${TAG}_bench_${A}.json; tail -3 gpurun_out/${TAG}_bench_${A}.err ;;
    benchall)
      timeout 2400 python bench.py > gpurun_out/${TAG}_bench_default.json 2> gpurun_out/${TAG}_bench_default.err
      echo "bench default exit $?"; python tools/bench_brief.py gpurun_out/${TAG}_bench_default.json; tail -3 gpurun_out/${TAG}_bench_default.err ;;
    launches)
      timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/${TAG}_launches_${A}.csv \
        python bench.py --workload $A --steps 1 --warmup 1 --no-cpu-baseline --others "" --parity-seconds 0.5 > gpurun_out/${TAG}_ncu_launch_${A}.log 2>&1
      echo "launch list $A exit $?" ;;
    ncu)
      timeout 1500 ncu --set full --clock-control none --import-source on -k regex:$B -s ${C:-3} -c 1 -f -o gpurun_out/${TAG}_${A}_${B} \
        python bench.py --workload $A --steps 1 --warmup 1 --no-cpu-baseline --others "" --parity-seconds 0.5 > gpurun_out/${TAG}_ncu_${A}_${B}.log 2>&1
      echo "ncu $A $B exit $?" ;;
    tpi)   # bench:<workload> with GAUXC_B200_TILES_PER_ITEM=<B>
      GAUXC_B200_TILES_PER_ITEM=$B timeout 900 python bench.py --workload $A --steps 3 --warmup 2 --others "" --no-cpu-baseline --parity-seconds 1 > gpurun_out/${TAG}_tpi${B}_${A}.json 2> gpurun_out/${TAG}_tpi${B}_${A}.err
      echo "tpi $B $A exit $?"; python tools/bench_brief.py gpurun_out/${TAG}_tpi${B}_${A}.json | head -3 ;;
    lib)   # bench:<workload> with the diagnostic library variant <B>
      GAUXC_B200_LIB=$PWD/gauxc_b200/libgauxc_b200_$B.so timeout 900 python bench.py --workload $A --steps 3 --warmup 2 --others "" --no-cpu-baseline --parity-seconds 1 > gpurun_out/${TAG}_lib${B}_${A}.json 2> gpurun_out/${TAG}_lib${B}_${A}.err
      echo "lib $B $A exit $?"; python tools/bench_brief.py gpurun_out/${TAG}_lib${B}_${A}.json | head -3 ;;
    fast)  # bench:<workload>[:steps] without the CPU baseline and with a 2 s parity sample
      timeout 900 python bench.py --workload $A --steps ${B:-3} --warmup 3 --others "" --no-cpu-baseline --parity-seconds 2 > gpurun_out/${TAG}_fast_${A}.json 2> gpurun_out/${TAG}_fast_${A}.err
      echo "fast $A exit $?"; python tools/bench_brief.py gpurun_out/${TAG}_fast_${A}.json | head -4; tail -3 gpurun_out/${TAG}_fast_${A}.err ;;
    gradlaunches)
      timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/${TAG}_launches_grad_${A}.csv \
        python tools/grad_once.py $A > gpurun_out/${TAG}_ncu_launch_grad_${A}.log 2>&1
      echo "grad launch list $A exit $?" ;;
    grad)
      timeout 900 python tools/grad_report.py > gpurun_out/${TAG}_exc_grad.txt 2>&1; echo "grad exit $?"; cat gpurun_out/${TAG}_exc_grad.txt ;;
    *) echo "unknown stage $ST" ;;
  esac
done
