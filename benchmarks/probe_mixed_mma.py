import torch, sys
sys.path.insert(0,'/root/repo')
from dfno_b200.ops import build
from dfno_b200.ops.gemm import pad_operator
C=build.load()
M,K,N=1000,128,48
g=torch.Generator(device='cuda').manual_seed(0)
A=torch.randn(M,K,device='cuda',generator=g)
B=(torch.randn(N,K,device='cuda',generator=g)/K**0.5)
A16=A.to(torch.float16)
out=torch.zeros(M,N,device='cuda',dtype=torch.float32)
epi=[0,1,N,0, 0,0,0,0, 0,0,0,0, 1,1,0,0, 0,0,1,0]
C.dft_gemm(A16,M,K,K,pad_operator(B),N,epi,[out.data_ptr()],None,0,0,None,None,0.0,True)
torch.cuda.synchronize()
ref=A16.float()@B.to(torch.bfloat16).float().t()
print("MIXED f16xbf16 max err", float((out-ref).abs().max()), "ref max", float(ref.abs().max()))
