import sys, os
sys.path.insert(0, "/root/repo")
import numpy as np, torch
from gan_lib_tensorflow_b200 import framework, functional as F
from gan_lib_tensorflow_b200.framework import Var
store = framework.reset_default_graph("cuda")
def rel(a,b): return float(np.linalg.norm(a-b)/(np.linalg.norm(b)+1e-30))
for (shape, mean, std) in [((1,256,256,8),0.0,1.0),((1,256,256,8),3.0,0.2),((1,256,256,8),10.0,0.1),((2,4,4,64),1.0,0.3),((1,512,512,16),0.5,1.0)]:
    rs=np.random.RandomState(1)
    x=(rs.standard_normal(shape)*std+mean).astype("float32")
    cot=rs.standard_normal(shape).astype("float32")
    c=shape[-1]
    with store.variable_scope("T%d"%rs.randint(1e9)):
        g_v=store.get_variable("gamma", initializer=np.ones(c,"float32")); b_v=store.get_variable("beta", initializer=np.zeros(c,"float32"))
    g_v.grad=torch.zeros_like(g_v.data); b_v.grad=torch.zeros_like(b_v.data)
    xv=Var(torch.from_numpy(x).cuda(), requires_grad=True)
    with store.gradient_tape() as tape:
        out,_=F.norm_act(xv, stats="instance", eps=1e-5, gamma=g_v, beta=b_v, out_dtype=torch.float32)
        tape.backward(out, grad=torch.from_numpy(cot).cuda())
    torch.cuda.synchronize()
    xt=torch.from_numpy(x).double().requires_grad_(True)
    m=xt.mean(dim=(1,2),keepdim=True); v=xt.var(dim=(1,2),unbiased=False,keepdim=True)
    yo=(xt-m)*torch.rsqrt(v+1e-5)
    dx,=torch.autograd.grad(yo,xt,torch.from_numpy(cot).double())
    print(shape,mean,std,"out",rel(out.data.cpu().numpy(),yo.detach().numpy()),"dx",rel(xv.grad.cpu().numpy(),dx.numpy()),
          "dgamma",rel(g_v.grad.cpu().numpy(),(torch.from_numpy(cot).double()*yo.detach()).sum((0,1,2)).numpy()),
          "dbeta",rel(b_v.grad.cpu().numpy(),cot.astype("float64").sum((0,1,2))))
