import csv, subprocess, sys
rep=sys.argv[1]
out=subprocess.run(['ncu','-i',rep,'--page','raw','--csv'],capture_output=True,text=True).stdout
rows=list(csv.reader(out.splitlines()))
hdr=rows[0]; units=rows[1]; data=rows[2:]
pat=sys.argv[2:] or ['gpu__time_duration.sum','dram__bytes_read.sum','dram__bytes_write.sum','gpu__dram_throughput.avg.pct','sm__warps_active.avg.pct','launch__registers','launch__occupancy_limit','lts__t_sector_hit_rate.pct','l1tex__t_sector_hit_rate.pct','sm__throughput.avg.pct','launch__grid_size','launch__waves','smsp__issue_active.avg.pct','sm__inst_executed.sum','smsp__average_warp','smsp__warp_issue_stalled','lts__t_bytes.sum','launch__shared_mem','sm__ctas_launched','achieved_occupancy','sm__maximum_warps']
for r in data:
    print('---', r[hdr.index('Kernel Name')][:80])
    for i,h in enumerate(hdr):
        if any(p in h for p in pat):
            v=r[i]
            if v in ('0','0.000000','') : continue
            print(f"  {h:90s} {v:>18s} {units[i]}")
