"""Fingerprint (op count + SHA-256 over op codes, descriptor bytes and notes) of the kernel programs the planner emits for the
BASELINE configs (reduced sizes) and a few other families, on the CPU.  Run it on two checkouts (`python tools/plan_fingerprint.py
<repo root>`) to prove that a planner / builder change leaves the programs of the measured configurations byte-identical."""
import sys, hashlib, ctypes
root=sys.argv[1] if len(sys.argv) > 1 else __import__('os').path.dirname(__import__('os').path.dirname(__import__('os').path.abspath(__file__)))
sys.path.insert(0,root+'/tests'); sys.path.insert(0,root); sys.path.insert(0,root+'/tf-1d-2d-segmentation-end2endpipelines_b200')
from b2seg.models2d import unet_model_builder
from b2seg.models1d import UNet, BCDUNet
from b2seg.planner import Planner
from desc_emulator import PlanMem
adam=dict(lr=2e-4,beta1=.9,beta2=.999,eps=1e-7)
def dump(name,g,N,losses):
    mem=PlanMem()
    pl=Planner(g,N,mem.alloc_bytes,training=True,losses=losses,adam=adam).build()
    h=hashlib.sha256()
    n=0
    for ph in (0,1,2):
        for (op,d,note) in pl.ops[ph]:
            h.update(bytes([op])); h.update(bytes(ctypes.string_at(ctypes.addressof(d), ctypes.sizeof(d)))); h.update(note.encode()); n+=1
    print(name, n, h.hexdigest()[:16])
dump("cfg2", unet_model_builder('UNet',64,64,64,5,train_mode='from_scratch').build_graph(), 2, ["bce"])
g=unet_model_builder('UNetPP',64,64,16,3,output_nums=4,ds=1,ag=1,final_activation='softmax',train_mode='from_scratch').build_graph()
dump("cfg3", g, 2, ["cce"]+["mse"]*(len(g.outputs)-1))
dump("cfg4", unet_model_builder('MultiResUNet',64,64,32,3,num_channels=1,train_mode='from_scratch').build_graph(), 2, ["bce"])
dump("cfg5", unet_model_builder('UNet',64,64,16,3,lstm=1,dense_loop=3,train_mode='from_scratch').build_graph(), 2, ["bce"])
dump("cfg1", UNet(256,5,1,16,3,problem_type='Classification',output_nums=2,ds=0).UNet().graph, 2, ["cce"])
g=UNet(256,3,2,16,3,ds=1,t=2).R2UNet().graph; dump("r2unet", g, 2, ["mse"]*len(g.outputs))
g=unet_model_builder('UNet3P',64,64,16,3,ds=1,train_mode='from_scratch').build_graph(); dump("unet3p", g, 2, ["bce"]+["mse"]*(len(g.outputs)-1))
g=unet_model_builder('KSSNet',64,64,32,3,ag=1,train_mode='from_scratch').build_graph(); dump("kssnet", g, 2, ["bce"])
g=unet_model_builder('UNet4P',64,64,16,3,ds=1,train_mode='from_scratch').build_graph(); dump("unet4p", g, 2, ["bce"]+["mse"]*(len(g.outputs)-1))


# ---- the models of the GPU test-suite (tests/test_gpu_model.py), at their test sizes
def dump2(name, g, N=4):
    losses = []
    for n in g.outputs:
        a = n.attrs.get("activation")
        losses.append("cce" if a == "softmax" else ("bce" if a == "sigmoid" else "mse"))
    dump(name, g, N, losses)


from b2seg.models2d import fpn_model_builder  # noqa: E402
FAMILY_CASES = [
    ("UNetPP", dict(ds=1, ag=1, output_nums=4, final_activation="softmax"), 64, 16, 3),
    ("UNet", dict(lstm=1, dense_loop=3), 64, 16, 3),
    ("UNet3P", dict(ds=1), 64, 16, 3),
    ("UNetE", dict(is_transconv=False, ag=1, ds=1), 32, 16, 2),
    ("MultiResUNet", dict(), 64, 32, 3),
    ("MultiResUNet", dict(is_transconv=False, ds=1), 32, 16, 2),
    ("UNet", dict(ae=1, feature_number=64), 64, 16, 3),
    ("UNet4P", dict(ds=1), 64, 16, 3),
    ("AHNet", dict(), 64, 16, 3),
    ("MultiResUNet3P", dict(ds=1), 64, 32, 3),
    ("KSSNet", dict(ag=1), 64, 32, 3),
]
for dec,kw,size,width,depth in FAMILY_CASES:
    dump2("fam "+dec+str(kw), unet_model_builder(dec,size,size,width,depth,num_channels=3,train_mode="from_scratch",**kw).build_graph())
for kw in (dict(), dict(ds=1, ag=1)):
    dump2("fpn"+str(kw), fpn_model_builder("FPN",64,64,16,3,num_channels=3,train_mode="from_scratch",**kw).build_graph())
for var,kw in [("RUNet", dict(ds=1, t=2)), ("R2UNet", dict(ds=1, ag=1, t=2)), ("R2UNetPP", dict(ds=1, t=1)),("R2UNet3P", dict(ds=1, t=1)), ("UNet4P", dict(ds=1, ag=1)), ("MultiResUNet3P", dict(ds=1))]:
    dump2("1d "+var, getattr(UNet(256,3,2,16,3,problem_type="Regression",output_nums=1,**kw),var)().graph)
dump2("1d bcd", BCDUNet(256,3,2,16,3,ds=1,ag=1,lstm=1,dense_loop=2).BCDUNet().graph)
dump2("ae2d", unet_model_builder("UNet",32,32,8,2,num_channels=1,ae=1,feature_number=32,train_mode="from_scratch").build_graph())
dump2("ae1d", UNet(64,2,1,8,3,ae=1,ds=0,feature_number=16).UNet().graph)
dump2("cfg2 16", unet_model_builder("UNet",64,64,16,2,num_channels=3,train_mode="from_scratch").build_graph(), 8)
dump2("1d unet", UNet(128,2,1,16,3,problem_type="Regression",output_nums=1,ds=1).UNet().graph, 8)
