"""Key metrics per launch from an ncu report: python scripts/ncu_metrics.py rep.ncu-rep"""
import csv, subprocess, sys
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = rows[0]
want = {'Kernel Name':'name','launch__grid_size':'grid','gpu__time_duration.sum':'us','dram__bytes_read.sum':'rdMB','dram__bytes_write.sum':'wrMB',
 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active':'tensor%','sm__warps_active.avg.pct_of_peak_sustained_active':'warps%',
 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed':'dram%','sm__throughput.avg.pct_of_peak_sustained_elapsed':'sm%',
 'launch__registers_per_thread':'regs','lts__t_sector_hit_rate.pct':'l2hit%','smsp__cycles_active.avg':'smsp_cyc','sm__cycles_elapsed.max':'cyc',
 'launch__occupancy_limit_shared_mem':'occ_smem','launch__occupancy_limit_registers':'occ_reg','l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum':'bankconf'}
idx = [(hdr.index(k), v) for k, v in want.items() if k in hdr]
units = rows[1]
for r in rows[2:]:
    if len(r) < len(hdr): continue
    print(' '.join(f"{v}={r[i][:44]}{'' if v in ('name','grid') else units[i][:6]}" for i, v in idx))
