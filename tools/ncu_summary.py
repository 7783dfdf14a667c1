import csv,sys,subprocess
rep=sys.argv[1]
out=subprocess.run(['ncu','-i',rep,'--page','raw','--csv'],capture_output=True,text=True).stdout
rows=list(csv.reader(out.splitlines()))
h=rows[0]; r=rows[2]
d=dict(zip(h,r))
keys=['gpu__time_duration.sum','sm__cycles_elapsed.avg','smsp__cycles_active.avg','sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed',
'sm__inst_executed_pipe_tensor_subpipe_hmma.avg.pct_of_peak_sustained_active','dram__bytes_read.sum','dram__bytes_write.sum','dram__throughput.avg.pct_of_peak_sustained_elapsed',
'lts__t_bytes.sum','lts__throughput.avg.pct_of_peak_sustained_elapsed','l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed','l1tex__throughput.avg.pct_of_peak_sustained_elapsed',
'sm__throughput.avg.pct_of_peak_sustained_elapsed','launch__grid_size','launch__registers_per_thread','sm__cycles_active.avg','gpc__cycles_elapsed.max','sm__clock_rate' if False else 'gpc__cycles_elapsed.avg.per_second',
'smsp__inst_executed.sum','l1tex__m_xbar2l1tex_read_bytes.sum','lts__t_sector_hit_rate.pct','smsp__sass_inst_executed_op_tmem_ldt.sum']
for k in keys:
    for kk in d:
        if kk.endswith(k) or kk==k: print(f'{kk} = {d[kk]}')
for kk in d:
    if 'stall' in kk and 'pct' not in kk and 'ratio' in kk: print(kk,'=',d[kk])
