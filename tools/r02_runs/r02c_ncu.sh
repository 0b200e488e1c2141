#!/bin/bash
# round 2, third session: ncu evidence for the final headline kernel - launch list of the bench command + one full capture
set -u
O=gpurun_out; mkdir -p $O
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r02c_launches_bench.csv python bench.py --steps 5 --warmup 3 --no-cpu-baseline --e2e-steps 1 > $O/r02c_launches_bench.out 2>&1
MET="gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,dram__bytes_read.sum.per_second,dram__bytes_write.sum.per_second,lts__t_sector_hit_rate.pct,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed,sm__throughput.avg.pct_of_peak_sustained_elapsed,gpu__compute_memory_throughput.avg.pct_of_peak_sustained_elapsed,launch__registers_per_thread,launch__shared_mem_per_block_dynamic,launch__grid_size,launch__block_size,sm__cycles_elapsed.max,smsp__inst_executed.sum,sm__warps_active.avg.pct_of_peak_sustained_active,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum,dram__throughput.avg.pct_of_peak_sustained_elapsed,lts__t_bytes.sum,lts__throughput.avg.pct_of_peak_sustained_elapsed"
for c in headline; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:"blockmix_kernel" -s 2 -c 2 -o $O/r02c_${c}_full -f env NCU_CALLS=4 python tools/ncu_cases.py $c > $O/r02c_ncu_$c.out 2>&1
  ncu -i $O/r02c_${c}_full.ncu-rep --page raw --csv --metrics $MET > $O/r02c_${c}_ncu_raw.csv 2>/dev/null
  ncu -i $O/r02c_${c}_full.ncu-rep --page details > $O/r02c_${c}_ncu_details.txt 2>/dev/null
  echo "== $c"; tail -2 $O/r02c_ncu_$c.out
done
grep -c blockmix $O/r02c_launches_bench.csv; tail -3 $O/r02c_launches_bench.csv | cut -c1-200
cut -c1-400 $O/r02c_headline_ncu_raw.csv | tail -2
