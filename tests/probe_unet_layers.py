"""Layer-by-layer comparison of the Pix2Pix U-Net (product vs bf16-operand oracle). Not a pytest file."""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np, torch
from gan_lib_tensorflow_b200 import framework, functional as F
from gan_lib_tensorflow_b200.Pix2Pix import networks as P
from oracle import ops as O_ops, pix2pix as OP, tfshim

def rel(a, b): return float(np.linalg.norm(a - b) / (np.linalg.norm(b) + 1e-30))
size, ngf, n = int(sys.argv[1]) if len(sys.argv) > 1 else 512, 8, 1
rs = np.random.RandomState(51)
x = rs.uniform(-1, 1, size=(n, size, size, 3)).astype("float32")
store = framework.reset_default_graph("cuda", u_seed=2)
np.random.seed(0)
out = P.unet_g(F.Var(torch.from_numpy(x).cuda()), 3, ngf)
pl = [l.data.float().cpu().numpy() for l in P.unet_g.last_layers]
for mode in (True, False):
    O_ops.BF16_OPERANDS = mode
    np.random.seed(0)
    g = tfshim.Graph(dtype=torch.float32, u_seed=2)
    with torch.no_grad():
        OP.unet_g(g, torch.from_numpy(x), 3, ngf)
    ol = [l.numpy() for l in OP.unet_g.last_layers]
    print("bf16-operand oracle" if mode else "fp32 oracle")
    for i, (a, b) in enumerate(zip(pl, ol)):
        print(f"  layer {i:2d} {str(a.shape):>22} rel {rel(a, b):.3e}  |ref| mean {np.abs(b).mean():.3e} std {b.std():.3e}")
O_ops.BF16_OPERANDS = False
