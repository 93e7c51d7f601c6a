"""Companion of tools/numeric_sweep.py over the SHAPE parameters: model widths that are not multiples of 8, 1-5 input channels, 1-9
output classes, depths 1-4, rectangular inputs, dense_loop, alpha, q, t, 1D kernel sizes 1-7 — 592 graphs, same float64 emulator-vs-oracle
comparison.  Run like numeric_sweep.py (one torch thread per shard).  Last run: 582 agree; the rest are MultiRes widths whose
int(alpha*W*0.167) is 0 (now a ValueError like Keras') and depth-4 / recurrent Self-ONN graphs whose cubes of cubes leave float32 range
(test conditioning, not lowering)."""
import sys, itertools, time, os
root=os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0,root+'/tests'); sys.path.insert(0,root); sys.path.insert(0,root+'/tf-1d-2d-segmentation-end2endpipelines_b200')
import numpy as np, torch
torch.set_num_threads(1)
from b2seg.models2d import unet_model_builder, IN_SCOPE_DECODERS
from b2seg.models1d import UNet
from b2seg.planner import PlanError
from oracle.ref_models import Ref1D, Ref2D
from test_plan_cpu import _run
from test_plan_families_cpu import _targets
shard, nshard = int(sys.argv[1]), int(sys.argv[2])
jobs=[]
for dec in IN_SCOPE_DECODERS:
    for W in (4,12,20): jobs.append(("2d",dec,(16,16),2,dict(ds=1),W))
    for nc in (1,3,5): jobs.append(("2d",dec,(16,16),2,dict(num_channels=nc),8))
    for on,fa in ((2,"softmax"),(5,"softmax"),(9,"softmax"),(3,"sigmoid"),(1,"linear")): jobs.append(("2d",dec,(16,16),2,dict(output_nums=on,final_activation=fa),8))
    for d in (1,3,4): jobs.append(("2d",dec,(2**(d+1),2**(d+1)),d,dict(ds=1),8))
    jobs.append(("2d",dec,(16,32),2,dict(ds=1,ag=1),8))
    for dl in (0,3): jobs.append(("2d",dec,(16,16),2,dict(dense_loop=dl),8))
    for al in (0.5,1.67,2.5): jobs.append(("2d",dec,(16,16),2,dict(alpha=al),16))
    for q in (1,2,4): jobs.append(("2d",dec,(16,16),2,dict(q=q,ds=1),8))
V1=["UNet","UNetE","UNetP","UNetPP","UNet3P","UNet4P","MultiResUNet","MultiResUNet3P","RUNet","R2UNet","R2UNetPP","R2UNet3P","SelfUNetPP","SelfR2UNetPP","SelfUNet3P"]
for var in V1:
    for ks in (1,2,4,5,7): jobs.append(("1d",var,32,2,dict(),8,ks))
    for W in (4,12,20): jobs.append(("1d",var,32,2,dict(),W,3))
    for on,pt in ((3,"Regression"),(2,"Classification"),(9,"Classification")): jobs.append(("1d",var,32,2,dict(problem_type=pt,output_nums=on),8,3))
    for d in (1,3,4): jobs.append(("1d",var,2**(d+2),d,dict(),8,3))
    for t in (1,3): jobs.append(("1d",var,32,2,dict(t=t),8,3))
    for al in (0.5,1.67): jobs.append(("1d",var,32,2,dict(alpha=al),16,3))
jobs=jobs[shard::nshard]
ok=bad=0; t0=time.time()
for j in jobs:
    rng=np.random.default_rng(9)
    selfonn="Self" in j[1]
    tol=dict(act_atol=2e-7, act_rtol=2e-6, grad_rtol=2e-5, adam_atol=2e-4, loss_rtol=1.0 if selfonn else 1e-6)
    try:
        if j[0]=="2d":
            _,dec,(H,Wd),d,kw,W=j
            kw=dict(dict(num_channels=2),**kw)
            g=unet_model_builder(dec,H,Wd,W,d,train_mode="from_scratch",**kw).build_graph()
            x=torch.from_numpy((0.5 if selfonn else 1.0)*rng.random((2,H,Wd,kw["num_channels"]),dtype=np.float32))
            ts,losses=_targets(g,2,rng,2)
            try: _run(g,Ref2D(dec,H,Wd,W,d,**kw),x,ts,losses,2,strict=True,**tol)
            except KeyError as e:
                if "oracle: weight" not in str(e): raise
                _run(g,Ref2D(dec,H,Wd,W,d,**kw),x,ts,losses,2,strict=False,**tol)
        else:
            _,var,Ln,d,kw,W,ks=j
            g=getattr(UNet(Ln,d,2,W,ks,**kw),var)().graph
            x=torch.from_numpy(((0.3 if selfonn else 1.0)*(rng.random((2,Ln,2))-0.5)).astype(np.float32))
            ts,losses=_targets(g,2,rng,1)
            try: _run(g,Ref1D(var,Ln,d,2,W,ks,**kw),x,ts,losses,1,strict=True,**tol)
            except KeyError as e:
                if "oracle: weight" not in str(e): raise
                _run(g,Ref1D(var,Ln,d,2,W,ks,**kw),x,ts,losses,1,strict=False,**tol)
        ok+=1; print("OK",j,flush=True)
    except PlanError as e:
        bad+=1; print("PLANERR",j,str(e)[:160],flush=True)
    except AssertionError as e:
        bad+=1; print("ASSERT",j,str(e)[:200].replace("\n"," "),flush=True)
    except Exception as e:
        bad+=1; print("ERR",j,type(e).__name__,str(e)[:160].replace("\n"," "),flush=True)
print(f"shard {shard}: ok {ok} bad {bad} in {time.time()-t0:.0f}s",flush=True)
