"""Stand-alone simulator step kernel at growing env counts (bench.py's simulator_sweep), with the (pair, chunk, action)
outcome table (default) and with every step gathering (MANSY_NO_OUTCOME_TABLE=1)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from mansy_immersivevideostreaming_b200.simulator import ViewportTiler
tables = bench.workload_tables(ViewportTiler(device=0).chunk_masks, 4096)
for flag in ("0", "1"):
    os.environ["MANSY_NO_OUTCOME_TABLE"] = flag
    print("outcome table:", "off (gather every step)" if flag == "1" else "on")
    for r in bench.simulator_sweep(tables, 0, sizes=(4096, 65536, 262144, 1048576)):
        print(f"{r['env']:10s} {r['envs_per_gpu']:8d} envs  {r['ms_per_step']*1e3:9.1f} us  {r['chunk_steps_per_s']/1e6:8.1f} M steps/s  "
              f"{r['achieved_GBps']:7.1f} GB/s  {100*r['frac_of_hbm_peak']:5.1f}% of HBM peak")
