#!/bin/bash
# Round-2 evidence session on one B200: tests, smoke, bench lines of every BASELINE config, reference arm, phase profiles, ncu launch list +
# full captures, DRAM traffic, sanitizers.   usage: bash tools/gpu_round2.sh TAG   (outputs under gpurun_out/TAG/)
cd "$(dirname "$0")/.."
TAG=${1:-r02}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > $OUT/gpu.txt
lscpu | grep -E "Model name|^CPU\(s\)|Thread|Socket" >> $OUT/gpu.txt
echo "=== pytest -m gpu"; timeout 1500 python -m pytest tests -m gpu -q --maxfail=10 2>&1 | tail -8 | tee $OUT/pytest_gpu.txt
echo "=== smoke"; timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5 | tee $OUT/smoke.txt
echo "=== bench (default = config 2, fp32-accurate)"; timeout 900 python bench.py 2>&1 | tail -1 | tee $OUT/bench.json
echo "=== bench reference arm"; timeout 900 python bench.py --impl reference --steps 3 --warmup 1 2>&1 | tail -1 | tee $OUT/bench_reference.json
echo "=== bench config 2, f16 family"; timeout 600 python bench.py --precision f16 --no-extras --no-cpu-baseline 2>&1 | tail -1 | tee $OUT/bench_f16.json
echo "=== bench config 1 (offline Model.forward, both arms)"; timeout 600 python bench.py --config 1 --no-cpu-baseline 2>&1 | tail -1 | tee $OUT/bench_config1.json
timeout 600 python bench.py --config 1 --impl reference --steps 3 --warmup 1 2>&1 | tail -1 | tee $OUT/bench_config1_reference.json
echo "=== bench config 3 (one GPU's share: M, 512 streams, bf16)"; timeout 900 python bench.py --config 3 --steps 5 2>&1 | tail -1 | tee $OUT/bench_config3.json
echo "=== bench config 3, f16 family"; timeout 900 python bench.py --config 3 --precision f16 --steps 5 --no-extras --no-cpu-baseline 2>&1 | tail -1 | tee $OUT/bench_config3_f16.json
echo "=== bench config 4 (one GPU's share: 48 kHz L, 256 streams)"; timeout 900 python bench.py --config 4 --steps 3 --no-cpu-baseline 2>&1 | tail -1 | tee $OUT/bench_config4.json
echo "=== bench config 4, f16 family"; timeout 900 python bench.py --config 4 --precision f16 --steps 5 --no-extras --no-cpu-baseline 2>&1 | tail -1 | tee $OUT/bench_config4_f16.json
echo "=== bench config 5 (batch-1 latency sweep)"; timeout 900 python bench.py --config 5 --no-cpu-baseline 2>&1 | tail -1 | tee $OUT/bench_config5.json
echo "=== phase profiles"
FE_PRECISION=fp32x3 timeout 120 python tools/gpu_diag.py --prof 16k_b 256 50 2>&1 | tee $OUT/phase_profile_16k_b_fp32x3.txt
FE_PRECISION=f16 timeout 120 python tools/gpu_diag.py --prof 16k_b 256 50 2>&1 | tee $OUT/phase_profile_16k_b_f16.txt
FE_PRECISION=fp32 timeout 120 python tools/gpu_diag.py --prof 16k_b 256 50 2>&1 | tee $OUT/phase_profile_16k_b_fp32.txt
FE_HOP_SLICING=0 FE_PRECISION=bf16 timeout 120 python tools/gpu_diag.py --prof 16k_m 512 20 2>&1 | tee $OUT/phase_profile_16k_m_bf16.txt
FE_HOP_SLICING=0 FE_PRECISION=f16 timeout 120 python tools/gpu_diag.py --prof 48k_l 256 10 2>&1 | tee $OUT/phase_profile_48k_l_f16.txt
echo "=== timings"
for p in fp32x3 f16 tf32 bf16 fp32; do
for a in "16k_t 256 200" "16k_b 256 200" "16k_b 1 200" "16k_b 4096 40"; do
  FE_PRECISION=$p timeout 120 python tools/gpu_diag.py --time $a 2>&1 | grep TIME
done; done | tee $OUT/timings.txt
for a in "16k_s 256 100" "16k_m 256 60" "16k_m 512 64" "16k_l 256 64" "48k_l 256 64" "48k_m 256 64" "16k_t 4096 50"; do
  FE_PRECISION=f16 timeout 120 python tools/gpu_diag.py --time $a 2>&1 | grep TIME
done | tee -a $OUT/timings.txt
echo "=== offline schedules / tensor-core STFT"
timeout 600 python tools/offline_timing.py 2>&1 | grep -E "OFFLINE|rror" | tee $OUT/offline_timings.txt
for c in 16k_t:1:10 16k_b:1:10 16k_m:1:10 48k_l:1:10; do FE_TP_TIMING=1 timeout 100 python tools/offline_timing.py $c 2>&1 | grep "per launch" | tail -1 | sed "s/^/$c /" | tee -a $OUT/offline_timings.txt; done
timeout 300 python tools/stft_timing.py 2>&1 | grep -E "STFT|rror" | tee $OUT/stft_timings.txt
echo "=== ncu launch list"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $OUT/launches.csv \
   python bench.py --steps 2 --warmup 3 --seconds 2 --no-cpu-baseline --no-extras > $OUT/ncu_launch_bench.log 2>&1
echo "=== ncu full (fp32x3, then f16)"
for p in fp32x3 f16; do
FE_PRECISION=$p timeout 900 ncu --set full --clock-control none --import-source on -k regex:fe_fused -s 1 -c 1 -o $OUT/prof_fused_$p \
   python tools/gpu_diag.py --time 16k_b 256 60 > $OUT/ncu_full_$p.log 2>&1
ncu -i $OUT/prof_fused_$p.ncu-rep --page raw --csv > $OUT/prof_fused_${p}_raw.csv 2>/dev/null
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:fe_stft_gemm -s 2 -c 1 -o $OUT/prof_stft_gemm python tools/stft_timing.py 16k_b:256:626 > $OUT/ncu_stft_gemm.log 2>&1
ncu -i $OUT/prof_stft_gemm.ncu-rep --page raw --csv > $OUT/prof_stft_gemm_raw.csv 2>/dev/null
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gru_scan -s 4 -c 1 -o $OUT/prof_gru_scan python tools/offline_timing.py 16k_b:1:10 > $OUT/ncu_gru_scan.log 2>&1
ncu -i $OUT/prof_gru_scan.ncu-rep --page raw --csv > $OUT/prof_gru_scan_raw.csv 2>/dev/null
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $OUT/offline_launches_16k_b.csv python tools/offline_timing.py 16k_b:1:10 > /dev/null 2>&1
echo "=== dram traffic of the bench launch (roofline.traffic)"
timeout 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:fe_fused -s 3 -c 1 --csv \
   --log-file $OUT/bench_traffic.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-extras > $OUT/bench_traffic_run.log 2>&1
echo "=== sanitizers"
FE_PRECISION=fp32x3 timeout 300 compute-sanitizer --tool memcheck --log-file $OUT/memcheck.log python tools/gpu_diag.py 16k_b 2 3 3 > $OUT/memcheck_run.log 2>&1; tail -2 $OUT/memcheck.log
FE_PRECISION=f16 timeout 300 compute-sanitizer --tool memcheck --log-file $OUT/memcheck_l.log python tools/gpu_diag.py 16k_l 1 2 2 > $OUT/memcheck_l_run.log 2>&1; tail -2 $OUT/memcheck_l.log
FE_PRECISION=f16 timeout 600 compute-sanitizer --tool racecheck --log-file $OUT/racecheck.log python tools/gpu_diag.py 16k_t 2 2 2 > $OUT/racecheck_run.log 2>&1; grep -E "Error:|SUMMARY" $OUT/racecheck.log | sort | uniq -c | head
timeout 300 compute-sanitizer --tool memcheck --log-file $OUT/memcheck_offline_stft.log python -m pytest tests/test_gpu_parity.py tests/test_stft_gemm.py -m gpu -q -x -k "frame_parallel_long_utterance and 16k_b-3 or rfft and 16k_m-2" > $OUT/memcheck_offline_stft_run.log 2>&1; tail -2 $OUT/memcheck_offline_stft.log
FE_PRECISION=bf16 timeout 300 compute-sanitizer --tool memcheck --log-file $OUT/memcheck_sliced.log python tools/gpu_diag.py --time 16k_m 300 8 > $OUT/memcheck_sliced_run.log 2>&1; tail -2 $OUT/memcheck_sliced.log
rm -f $OUT/*.ncu-rep.tmp
ls -la $OUT
