import sys, time, os
sys.path.insert(0,'.'); sys.path.insert(0,'tests')
import torch, bench
for th in (8, 16, 32, 64):
    t0=time.perf_counter(); dt, rec = bench.cpu_run(1024, th)[:2]; print(th, 'B=1024', round(dt,2), flush=True)
