set -x
python bench.py > gpurun_out/r02f_bench_n1.json 2> gpurun_out/r02f_bench_n1.err || exit 1
python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-secondary > gpurun_out/r02f_plain_bench.json 2> gpurun_out/r02f_plain_bench.err || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/r02f_launches_bench.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-secondary > gpurun_out/r02f_ncu_launch.log 2>&1
python scripts/prof_case.py 8192 1 > gpurun_out/r02f_plain_case.log 2>&1 || exit 1
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:"gemm_nt|trsm_tile|chain_step" --csv --log-file gpurun_out/r02f_gemm_dram_one_eval.csv python scripts/prof_case.py 8192 1 > gpurun_out/r02f_ncu_dram.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"chain_step_kernel" -s 30 -c 2 -f -o gpurun_out/r02f_chain_step python scripts/prof_case.py 8192 1 > gpurun_out/r02f_ncu_chain.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"gemm_nt_persist" -s 8 -c 1 -f -o gpurun_out/r02f_gemm_persist python scripts/prof_case.py 8192 1 > gpurun_out/r02f_ncu_persist.log 2>&1
