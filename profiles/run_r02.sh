#!/bin/bash
# Round-2 evidence pass (one B200): everything profiles/README.md cites under r02/.  Run through gpurun from the repo root:
#   gpurun --timeout 2400 -- 'bash profiles/run_r02.sh'
# Writes into gpurun_out/r02/ (copied to profiles/r02/ on the CPU box afterwards).
set -u
O=gpurun_out/r02
mkdir -p $O
SHORT="--steps 10 --warmup 3 --no-cpu-baseline --no-other-kernels --no-parity --file-images 0 --e2e-images 8 --e2e-steps 1"
timeout 900 python -m pytest tests -m gpu -q > $O/gpu_tests.txt 2>&1
timeout 900 python bench.py > $O/bench_n1.json 2> $O/bench_n1.err
for m in 0 2; do MJX_K2_TC=$m timeout 300 python bench.py $SHORT > $O/bench_short_tc$m.json 2> /dev/null; done
MJX_GPU_HUFFMAN=0 MJ_BATCH_TRACE=1 timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-other-kernels --no-parity --file-images 1024 --e2e-images 8 --e2e-steps 1 > $O/bench_files_host_huffman.json 2> $O/bench_files_host_huffman.err
MJX_GPU_DECODE=0 MJ_BATCH_TRACE=1 timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-other-kernels --no-parity --file-images 1024 --e2e-images 8 --e2e-steps 1 > $O/bench_files_device_huffman.json 2> $O/bench_files_device_huffman.err
MJ_BATCH_TRACE=1 timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-other-kernels --no-parity --file-images 1024 --e2e-images 8 --e2e-steps 1 > $O/bench_files_device_decode_and_encode.json 2> $O/bench_files_device_decode_and_encode.err
# launch list of the bench command (serialised, cold: shares only)
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:'k[1-5]_|list_|generic_prepare|count_classes' -c 300 --csv \
    --log-file $O/launches_bench.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-parity --file-images 0 --e2e-images 8 --e2e-steps 1 > /dev/null 2>&1
# full captures of the dominant kernel and of K4's block kernel
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k2_generic_op_kernel -s 4 -c 1 -o $O/k2_generic_op_kernel -f python bench.py $SHORT > /dev/null 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k4_block_kernel -s 2 -c 2 -o $O/k4_block_kernel -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-parity --file-images 0 --e2e-images 8 --e2e-steps 1 > /dev/null 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k5_decode_kernel -s 1 -c 1 -o $O/k5_decode_kernel -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-parity --file-images 0 --e2e-images 8 --e2e-steps 1 > /dev/null 2>&1
timeout 600 python profiles/parity_report.py > $O/parity_report.json 2> $O/parity_report.err
timeout 600 python profiles/fuzz_parity.py 300 2027 > $O/fuzz_parity.json 2> $O/fuzz_parity.err
timeout 300 python profiles/api_latency.py > $O/api_latency.json 2> $O/api_latency.err
timeout 300 python profiles/configs_perf.py > $O/configs_perf.json 2> $O/configs_perf.err
ls -la $O
