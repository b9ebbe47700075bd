#!/bin/bash
# round 2 evidence: full GPU suite, smoke, default bench line, ncu launch list of the bench command, ncu --set full captures of the dominant
# kernels from the bench process (B = 8 and B = 32 launch shapes), summarised ON THE BOX (the reports exceed the 64 MiB copy-back limit)
mkdir -p gpurun_out
rm -f gpurun_out/*.ncu-rep
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r02_pytest_gpu.log 2>&1; tail -3 gpurun_out/r02_pytest_gpu.log
timeout 300 python __graft_entry__.py --smoke > gpurun_out/r02_smoke.log 2>&1; tail -1 gpurun_out/r02_smoke.log
timeout 900 python bench.py > gpurun_out/r02_bench_default_line.json 2> gpurun_out/r02_bench_default_line.err; tail -c 300 gpurun_out/r02_bench_default_line.json
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 4000 -c 1600 --csv --log-file gpurun_out/r02_launches_bench.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-gpu-baseline --no-config3 --diffusion-steps 20 > gpurun_out/r02_ncu_launches.log 2>&1
python scripts/summarize_launches.py gpurun_out/r02_launches_bench.csv > gpurun_out/r02_launches_bench_summary.txt; head -12 gpurun_out/r02_launches_bench_summary.txt
cap() {  # name regex skip count extra-bench-args
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:$2 -s $3 -c $4 -o /tmp/$1 python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-gpu-baseline --no-config3 --diffusion-steps 10 $5 > gpurun_out/$1.log 2>&1
  for k in $(seq 0 $(($4 - 1))); do python scripts/ncu_source_summary.py /tmp/$1.ncu-rep 24 $k > gpurun_out/${1}_launch$k.txt 2>&1; done
  ncu -i /tmp/$1.ncu-rep --page raw --csv 2>/dev/null | python -c "
import csv,sys
rows=list(csv.reader(sys.stdin)); hdr=rows[0]
keep=['Kernel Name','gpu__time_duration.sum','launch__grid_size','sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active','dram__bytes_read.sum','dram__bytes_write.sum','sm__warps_active.avg.pct_of_peak_sustained_active','smsp__issue_active.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active','lts__t_bytes.sum','sm__cycles_active.avg','launch__registers_per_thread','gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed']
idx=[hdr.index(k) for k in keep if k in hdr]
print(' | '.join(hdr[i] for i in idx)); print(' | '.join(rows[1][i] for i in idx))
for r in rows[2:]: print(' | '.join(r[i][:60] for i in idx))
" > gpurun_out/${1}_raw_metrics.txt
}
cap r02_ncu_attn2_b8 umma_attn2 96 3 ""
cap r02_ncu_chain_b8 umma_chain 100 5 ""
cap r02_ncu_attn2_b32 umma_attn2 48 3 "--batch 32"
cap r02_ncu_chain_b32 umma_chain 50 5 "--batch 32"
cp /tmp/r02_ncu_attn2_b32.ncu-rep gpurun_out/ 2>/dev/null
ls -la gpurun_out/ | tail -30; du -sh gpurun_out
