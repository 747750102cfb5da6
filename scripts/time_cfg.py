"""time forward / backward separately for a few configs (device-resident, CUDA events, warm)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from neurallaplace_b200 import laplace_reconstruct
from neurallaplace_b200.rep_func import LaplaceRepresentationFunc
dev = torch.device("cuda", 0)
B, T = 4096, 1024
for terms, D in [(33, 1), (33, 2), (33, 4), (65, 4), (129, 2), (33, 16), (129, 16)]:
    f = LaplaceRepresentationFunc(terms, D, 2).to(dev)
    p = torch.randn(B, 2, device=dev, requires_grad=True)
    t = torch.linspace(20.0 / T, 20.0, T, device=dev)
    g = torch.randn(B, T, D, device=dev)
    for _ in range(3):
        x = laplace_reconstruct(f, p, t, recon_dim=D, ilt_reconstruction_terms=terms); x.backward(g)
    torch.cuda.synchronize()
    fs, bs = [], []
    for _ in range(5):
        e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
        e[0].record(); x = laplace_reconstruct(f, p, t, recon_dim=D, ilt_reconstruction_terms=terms)
        e[1].record(); x.backward(g); e[2].record(); torch.cuda.synchronize()
        fs.append(e[0].elapsed_time(e[1])); bs.append(e[1].elapsed_time(e[2]))
    print(f"terms={terms} D={D} streams={os.environ.get('NL_B200_TC_STREAMS','0')}: fwd {min(fs):.3f} ms (max {max(fs):.3f})  bwd {min(bs):.3f} ms (max {max(bs):.3f})", flush=True)
