#!/bin/bash
# One GPU session: tests, smoke, bench, ncu launch list + full capture of the fused kernel, sanitizers, phase profile.
# usage: bash tools/gpu_round.sh TAG        (outputs under gpurun_out/TAG/)
cd "$(dirname "$0")/.."
TAG=${1:-r01}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > $OUT/gpu.txt
lscpu | grep -E "Model name|^CPU\(s\)|Thread|Socket" >> $OUT/gpu.txt
echo "=== pytest -m gpu"; timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 | tee $OUT/pytest_gpu.txt
echo "=== smoke"; timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 | tee $OUT/smoke.txt
echo "=== bench"; timeout 600 python bench.py 2>&1 | tail -2 | tee $OUT/bench.json
echo "=== bench tf32 variant"; timeout 600 python bench.py --precision tf32 --no-cpu-baseline 2>&1 | tail -1 | tee $OUT/bench_tf32.json
echo "=== bench fp32-exact variant"; timeout 600 python bench.py --precision fp32 --no-cpu-baseline 2>&1 | tail -1 | tee $OUT/bench_fp32.json
echo "=== bench reference arm"; timeout 600 python bench.py --impl reference --steps 2 --warmup 1 2>&1 | tail -1 | tee $OUT/bench_reference.json
echo "=== phase profile"
export FE_PRECISION=f16      # the bench default; the tf32 / fp32 profiles are taken explicitly
timeout 120 python tools/gpu_diag.py --prof 16k_b 256 50 2>&1 | tee $OUT/phase_profile_16k_b.txt
timeout 120 python tools/gpu_diag.py --prof 16k_t 256 50 2>&1 | tee $OUT/phase_profile_16k_t.txt
timeout 120 python tools/gpu_diag.py --prof 16k_l 148 20 2>&1 | tee $OUT/phase_profile_16k_l.txt
FE_PRECISION=tf32 timeout 120 python tools/gpu_diag.py --prof 16k_b 256 50 2>&1 | tee $OUT/phase_profile_16k_b_tf32.txt
FE_PRECISION=fp32 timeout 120 python tools/gpu_diag.py --prof 16k_b 256 50 2>&1 | tee $OUT/phase_profile_16k_b_fp32.txt
echo "=== timings"
for a in "16k_t 256 200" "16k_b 256 200" "16k_s 256 100" "16k_m 256 60" "16k_l 256 30" "16k_b 1 200" "16k_b 4096 40" "16k_m 512 40" "48k_l 256 20" "16k_t 4096 50"; do
  timeout 120 python tools/gpu_diag.py --time $a 2>&1 | grep TIME
done | tee $OUT/timings.txt
echo "=== ncu launch list"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $OUT/launches.csv \
   python bench.py --steps 2 --warmup 3 --seconds 2 --no-cpu-baseline > $OUT/ncu_launch_bench.log 2>&1
echo "=== ncu full"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:fe_fused -s 1 -c 1 -o $OUT/prof_fused \
   python tools/gpu_diag.py --time 16k_b 256 60 > $OUT/ncu_full.log 2>&1
ncu -i $OUT/prof_fused.ncu-rep --page raw --csv > $OUT/prof_fused_raw.csv 2>/dev/null
echo "=== dram traffic of the bench launch (roofline.traffic)"
timeout 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:fe_fused -s 3 -c 1 --csv \
   --log-file $OUT/bench_traffic.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > $OUT/bench_traffic_run.log 2>&1
echo "=== batch-1 latency sweep"
timeout 600 python tools/latency_sweep.py 2000 f16 2>&1 | tee $OUT/latency_sweep.jsonl
echo "=== sanitizers"
timeout 300 compute-sanitizer --tool memcheck --log-file $OUT/memcheck.log python tools/gpu_diag.py 16k_t 2 3 2 > $OUT/memcheck_run.log 2>&1; tail -3 $OUT/memcheck.log
timeout 300 compute-sanitizer --tool memcheck --log-file $OUT/memcheck_l.log python tools/gpu_diag.py 16k_l 1 2 2 > $OUT/memcheck_l_run.log 2>&1; tail -3 $OUT/memcheck_l.log
timeout 600 compute-sanitizer --tool racecheck --log-file $OUT/racecheck.log python tools/gpu_diag.py 16k_t 2 2 2 > $OUT/racecheck_run.log 2>&1; grep -E "Error:|SUMMARY" $OUT/racecheck.log | head
ls -la $OUT
