import os, sys, torch
ROOT=os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, ROOT+'/tools')
import pose_b200 as pb
from _inputs import spm_inputs
from extra_workloads import graph_time
dev=torch.device('cuda',0)
for n in (256,1024):
    c,j,cnt,t,x=spm_inputs(n,dev)
    ms=graph_time(lambda: pb.spm_loss_fused(x,t),10)
    ms2=graph_time(lambda: pb.spm_loss_fused(x,t,want_grad=False),10)
    print(f"N={n}: dense loss+grad {ms*1e3:.1f} us ({n*35*65536*3/ms/1e6/6550.7*100:.1f} %), loss only {ms2*1e3:.1f} us ({n*35*65536*2/ms2/1e6/6550.7*100:.1f} %)", flush=True)
