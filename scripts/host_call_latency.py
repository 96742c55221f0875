import sys,os,time,json
sys.path.insert(0,"scripts"); sys.path.insert(0,"simple-vector-db_b200"); sys.path.insert(0,".")
import numpy as np, torch, ctypes as C
from svdb import binding as B
from bench_extra import fill, DEV
torch.cuda.set_stream(torch.cuda.Stream(device=DEV))
def run(n,D,K,k):
    with B.Engine(D,K,reserve_rows=n) as e:
        fill(e,n,D,seed=3)
        q=np.random.rand(1,D)
        L=e.L; idx=np.empty((1,k),dtype=np.uint64); qp=q.ctypes.data_as(B._dp); ip=idx.ctypes.data_as(B._zp)
        for graphs in (0,1,0,1):
            e.set_option("host.graphs",graphs)
            for _ in range(20): L.svdb_nearest_batch(e.h,qp,1,D,k,ip,None,None)
            t0=time.perf_counter()
            for _ in range(2000): L.svdb_nearest_batch(e.h,qp,1,D,k,ip,None,None)
            dt=(time.perf_counter()-t0)/2000*1e6
            print(json.dumps({"rows":n,"dim":D,"kd_dim":K,"k":k,"graphs":graphs,"us_per_host_call":round(dt,2)}),flush=True)
run(10000,128,3,1)
run(100000,128,3,1)
run(4096,768,768,1)
run(4096,128,128,10)
run(1000000,128,128,1)
