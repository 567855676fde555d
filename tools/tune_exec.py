import sys, ctypes, json, time
sys.path.insert(0,'/root/repo')
import numpy as np, torch
import __graft_entry__ as ge
from bench import make_workload, CHUNK
pkg = ge.import_package(); lib = pkg.load_library()
n=16384
data, blob, offs, sizes = make_workload(0, n, 8)
codec = pkg.ZstdBatchCodec(level=3)
dev=torch.device('cuda')
d_comp=torch.from_numpy(blob).to(dev); d_out=torch.empty(n*CHUNK,dtype=torch.uint8,device=dev)
idx=np.arange(n,dtype=np.uint64)
t_in=torch.from_numpy((np.uint64(d_comp.data_ptr())+offs).astype(np.int64)).to(dev)
t_is=torch.from_numpy(sizes.astype(np.int64)).to(dev)
t_out=torch.from_numpy((np.uint64(d_out.data_ptr())+idx*np.uint64(CHUNK)).astype(np.int64)).to(dev)
caps=torch.full((n,),CHUNK,dtype=torch.int64,device=dev); osz=caps.clone(); st=torch.zeros(n,dtype=torch.int32,device=dev)
ws=torch.empty(codec.decompress_temp_size(n,sizes),dtype=torch.uint8,device=dev)
import itertools
res = {}
for rep in range(3):
    for k, l in ((8, 8), (7, 8), (6, 8), (8, 9), (7, 9), (6, 9), (8, 12), (7, 12)):
        lib.cuda_zstd_b200_tune_exec_ctas(k); lib.cuda_zstd_b200_tune_exec_ctas_last(l)
        for _ in range(2):
            osz.copy_(caps); codec.decompress_nosync(t_in,t_is,n,t_out,osz,st,ws)
        torch.cuda.synchronize()
        e0=torch.cuda.Event(enable_timing=True); e1=torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(6):
            osz.copy_(caps); codec.decompress_nosync(t_in,t_is,n,t_out,osz,st,ws)
        e1.record(); torch.cuda.synchronize()
        res.setdefault((k, l), []).append(e0.elapsed_time(e1)/6)
for kl, v in res.items():
    print(kl, [round(x, 3) for x in v], 'min', round(min(v), 3), 'ms', int(st.max()))
