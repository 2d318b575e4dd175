set -x
python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-secondary > gpurun_out/r02_plain_bench.json 2> gpurun_out/r02_plain_bench.err || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/r02_launches_bench.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-secondary > gpurun_out/r02_ncu_launch.log 2>&1
python scripts/prof_case.py 8192 1 > gpurun_out/r02_plain_case.log 2>&1 || exit 1
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:"gemm_nt|trsm_tile" --csv --log-file gpurun_out/r02_gemm_dram_one_eval.csv python scripts/prof_case.py 8192 1 > gpurun_out/r02_ncu_dram.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"matern_cov_kernel|contract_kernel" -c 2 -f -o gpurun_out/r02_k1_k4_final python scripts/prof_case.py 8192 1 > gpurun_out/r02_ncu_k1k4.log 2>&1
python scripts/prof_gemm.py > gpurun_out/r02_plain_gemm.log 2>&1 || exit 1
ncu --set full --clock-control none --import-source on -k regex:gemm_nt_tma -s 1 -c 1 -f -o gpurun_out/r02_gemm_tma python scripts/prof_gemm.py > gpurun_out/r02_ncu_gemm.log 2>&1
