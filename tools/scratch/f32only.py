import sys
sys.path[:0]=['/root/repo','/root/repo/warmup-fir-filter_b200']
import torch, numpy as np, firgpu
x=torch.randint(0,256,(1<<26,),dtype=torch.uint8,device='cuda')
y=torch.empty(1<<26,dtype=torch.float32,device='cuda')
for _ in range(3): firgpu.fir1d_ideal(x,[1/16,4/16,6/16,4/16,1/16],out_dtype=np.float32,out=y,validate=False)
torch.cuda.synchronize()
