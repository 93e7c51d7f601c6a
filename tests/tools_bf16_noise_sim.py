"""Report tool (CPU, not a test): how far does bf16 *storage* alone move a free-running comparison?  The float64 oracle is run
twice on the same shallow 2D UNet — exactly, and with every stored tensor (conv outputs, activations, pooled tensors, their
gradients) rounded to bf16 — and per-layer rel-L2 deviations are printed.  Result (depth 2, width 16, 64x64, batch 8):
activations 0.2-1.3 %, gradients of raw conv outputs 7-20 %, weight gradients 2-20 % — the same figures the B200 path shows
against the exact oracle, i.e. the end-to-end gradient deviation is a property of bf16 storage + ReLU/BN, not of the kernels."""
import sys
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tf-1d-2d-segmentation-end2endpipelines_b200')
import numpy as np, torch
from oracle.keras_ref import KerasRef, keras_loss
from oracle.ref_models import Ref2D
from b2seg.models2d import unet_model_builder
from b2seg.graph import init_params

class Q(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x): return x.to(torch.bfloat16).to(x.dtype)
    @staticmethod
    def backward(ctx, g): return g.to(torch.bfloat16).to(g.dtype)   # gradients are stored in bf16 too

class KQ(KerasRef):
    def _rec(self, name, y):
        if name.startswith('conv2d') or name.startswith('activation') or name.startswith('add') or name.startswith('max_pool'):
            y = Q.apply(y)
        return super()._rec(name, y)

kw = dict(num_channels=3, output_nums=1, dense_loop=1, is_transconv=True)
g = unet_model_builder("UNet", 64, 64, 16, 2, train_mode="from_scratch", **kw).build_graph()
params = init_params(g)
params = {k: torch.from_numpy(v).to(torch.bfloat16).double() if k.endswith('/kernel') else torch.from_numpy(v).double() for k, v in params.items()}
rng = np.random.default_rng(0)
x = torch.from_numpy(rng.random((8, 64, 64, 3), dtype=np.float32)).double()
y = torch.from_numpy((rng.random((8, 64, 64, 1)) > 0.6).astype(np.float32)).double()
ref = Ref2D("UNet", 64, 64, 16, 2, **kw)
res = {}
for cls in (KerasRef, KQ):
    tp = {k: v.clone() for k, v in params.items()}
    k = cls(2, params=tp, dtype=torch.float64, training=True, strict=True)
    out = ref(k, x)[0]
    keras_loss("bce", out, y, logits=k.logits["out"]).backward()
    res[cls.__name__] = (k, tp)
k0, t0 = res['KerasRef']; k1, t1 = res['KQ']
rel = lambda a, b: float((a-b).norm()/b.norm())
for name in k0.acts:
    if name.startswith('conv2d') and k0.acts[name].grad is not None:
        print(name, 'act', f"{rel(k1.acts[name].detach(), k0.acts[name].detach()):.2e}", 'grad', f"{rel(k1.acts[name].grad, k0.acts[name].grad):.2e}")
for key in t0:
    if key.endswith('kernel') and t0[key].grad is not None:
        print(key, f"{rel(t1[key].grad, t0[key].grad):.2e}")
