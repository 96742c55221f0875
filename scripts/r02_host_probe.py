import sys,os,time,json
sys.path.insert(0,"scripts"); sys.path.insert(0,"simple-vector-db_b200"); sys.path.insert(0,".")
import numpy as np, torch
from svdb import binding as B
from bench_extra import fill, DEV
torch.cuda.set_stream(torch.cuda.Stream(device=DEV))
n,D=1_000_000,128
with B.Engine(D,D,reserve_rows=n) as e:
    e.set_stream(torch.cuda.current_stream().cuda_stream)
    fill(e,n,D,seed=3)
    q=np.random.rand(1,D)
    ts=[]
    for i in range(14):
        t0=time.perf_counter(); e.nearest(q,1); ts.append(round((time.perf_counter()-t0)*1e6,1))
    print("own-stream-set-to-torch, per-call us:", ts, e.stats()["fp64_reruns"], e.stats()["exact_reruns"], flush=True)
    e.set_option("profile.scan_events", 1); e.set_option("profile.scan_events", 0)
    ts=[]
    for i in range(8):
        t0=time.perf_counter(); e.nearest(q,1); ts.append(round((time.perf_counter()-t0)*1e6,1))
    print("after option toggle:", ts, flush=True)
    q8=np.random.rand(8,D)
    ts=[]
    for i in range(8):
        t0=time.perf_counter(); e.nearest(q8,1); ts.append(round((time.perf_counter()-t0)*1e6,1))
    print("nq=8:", ts, flush=True)
    st=e.stats(); print({k:st[k] for k in ("kernels_launched","fp64_reruns","exact_reruns","coalesced_calls")})
