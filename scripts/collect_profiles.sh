#!/bin/bash
# Collects the measured evidence of a round into gpurun_out/<tag>_*: run on the GPU box through gpurun, e.g.
#   gpurun --timeout 3000 -- 'bash scripts/collect_profiles.sh r02'
tag=${1:-r02}
out=gpurun_out
set -x
python bench.py --impl reference --steps 5 --warmup 1 > $out/${tag}_bench_reference_cpu.json 2> $out/${tag}_bench_reference_cpu.err
python bench.py --steps 20 --warmup 5 > $out/${tag}_bench_n1.json 2> $out/${tag}_bench_n1.err
python bench.py --config cfg3 --steps 10 --warmup 3 --no-cpu-baseline > $out/${tag}_bench_cfg3.json 2> $out/${tag}_bench_cfg3.err
python bench.py --config cfg4 --steps 5 --warmup 3 --no-cpu-baseline > $out/${tag}_bench_cfg4_n1.json 2> $out/${tag}_bench_cfg4_n1.err
python bench.py --config cfg5 --steps 3 --warmup 2 > $out/${tag}_bench_cfg5.json 2> $out/${tag}_bench_cfg5.err
python scripts/bench_layers.py > $out/${tag}_bench_layers.txt 2>&1
python scripts/bench_layers.py --config cfg4 > $out/${tag}_bench_layers_cfg4.txt 2>&1
python scripts/bench_elementwise.py > $out/${tag}_bench_elementwise.txt 2>&1
ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file $out/${tag}_launches_train_step.csv python scripts/profile_step.py > $out/${tag}_prof_step.log 2>&1
python scripts/summarize_launches.py $out/${tag}_launches_train_step.csv > $out/${tag}_launches_train_step_summary.txt
ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file $out/${tag}_launches_predict.csv python scripts/profile_predict.py > $out/${tag}_prof_predict.log 2>&1
python scripts/summarize_launches.py $out/${tag}_launches_predict.csv > $out/${tag}_launches_predict_summary.txt
B200EM_CONFIG=cfg3 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file $out/${tag}_launches_cfg3.csv python scripts/profile_step.py > $out/${tag}_prof_cfg3.log 2>&1
python scripts/summarize_launches.py $out/${tag}_launches_cfg3.csv > $out/${tag}_launches_cfg3_summary.txt
python scripts/bench_head.py > $out/${tag}_bench_head.txt 2>&1
B200EM_CONFIG=cfg4 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file $out/${tag}_launches_cfg4.csv python scripts/profile_step.py > $out/${tag}_prof_cfg4.log 2>&1
python scripts/summarize_launches.py $out/${tag}_launches_cfg4.csv > $out/${tag}_launches_cfg4_summary.txt
ncu --set full --import-source on --clock-control none --profile-from-start off -f -o $out/${tag}_kernels python scripts/profile_kernels.py > $out/${tag}_prof_kernels.log 2>&1
ncu -i $out/${tag}_kernels.ncu-rep --page raw --csv > $out/${tag}_kernels_raw.csv 2> /dev/null
python scripts/reduce_ncu_raw.py $out/${tag}_kernels_raw.csv $out/${tag}_kernels_ncu_full.csv
rm -f $out/${tag}_kernels.ncu-rep
# compute-sanitizer over the tcgen05 kernel tests (memcheck + racecheck): small cases, bounded
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 1 python -m pytest tests/test_gpu_umma.py -x -q -k "(test_umma_conv_forward_and_dgrad or test_umma_wgrad) and case2 or (test_plain_conv_split_k and 3-case1)" > $out/${tag}_sanitizer_memcheck.log 2>&1
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 1 python -m pytest tests/test_gpu_umma.py -x -q -k "(test_umma_conv_forward_and_dgrad or test_umma_wgrad) and case2" > $out/${tag}_sanitizer_racecheck.log 2>&1
ls -la $out | tail -30
