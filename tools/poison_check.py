"""Runs one training step per shape with the workspace poisoned (0xFF = NaN patterns) on both tcgen05
paths: any read of scratch that was never written shows up as NaN in the gradients."""
import os, sys
os.environ["NCF_UMMA_MIN_B"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from ncf_b200 import ops
from ncf_b200.models import NCF
dev=torch.device("cuda:0")
for (U,I,f,L,B) in ((24,18,64,3,40),(24,18,64,3,128),(200,100,64,3,40),(24,18,32,3,40)):
    torch.manual_seed(0)
    model=NCF(U,I,f,L,0.0,"NeuMF-end").to(dev)
    g=ops.GradBuffers.allocate(model.abi_type(),f,L,U,I,B,dev); m=model.abi_struct()
    ws=torch.empty(ops.train_workspace_bytes(m,B),dtype=torch.uint8,device=dev); ws.fill_(0xFF)
    u=torch.randint(0,U,(B,),device=dev); i=torch.randint(0,I,(B,),device=dev); y=(torch.rand(B,device=dev)<0.3).float()
    loss=torch.zeros(1,dtype=torch.float64,device=dev); lg=torch.empty(B,device=dev)
    ops.mark_rows(m,g.struct(),u,i); ops.train_step_grads(m,g.struct(),u,i,y,None,1.0,loss,ws,lg); torch.cuda.synchronize()
    print((U,I,f,L,B), "loss", float(loss), {n:int(torch.isnan(getattr(g,n)).sum()) for n in ("g_user_gmf","g_item_gmf","g_user_mlp","g_item_mlp","g_tower")})
