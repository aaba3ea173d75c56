import sys, torch
sys.path.insert(0, ".")
import kp_b200
from kp_b200 import conv, tapconv as tc
dev = torch.device("cuda:0")
for (N,H,W,C,cout,stats) in [(32,128,128,32,32,True),(32,128,128,16,16,True),(32,128,128,64,64,False),(32,64,64,128,128,True)]:
    x = torch.randn((N,H,W,C), device=dev).to(torch.bfloat16)
    w = torch.randn((3,3,C,cout), device=dev) / (9*C)**0.5
    plan,(n,ho,wo) = tc.plan_conv_fwd([tuple(x.shape)],3,1,0,cout)
    wp = conv.pack_weights(plan, w)
    out = torch.empty((n,ho,wo,cout), device=dev, dtype=torch.bfloat16)
    st = (torch.zeros(plan.rows_pad, device=dev), torch.zeros(plan.rows_pad, device=dev)) if stats else None
    for i in range(2):
        conv.run_plan(plan,[x],wp,None,out,stats=st)
    torch.cuda.synchronize()
