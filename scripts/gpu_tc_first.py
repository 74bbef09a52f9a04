"""First-light check of the tcgen05 GPTQ kernel on one small shape (run under `timeout`)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np, torch
from oracle import gptq as ogptq
from sparsebit_b200 import _lib
from sparsebit_b200.gptq import cuda_kernel
M, K, N, GS = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3]), 128
rng = np.random.default_rng(0)
w = (rng.standard_normal((N, K)) / np.sqrt(K)).astype(np.float32)
s, z = ogptq.find_params_int4(w, GS)
qw, scales, zeros = ogptq.pack_int4(ogptq.quantize_weight(w, s, z, GS), s, z)
x = rng.standard_normal((M, K)).astype(np.float32)
dev = torch.device("cuda:0")
lib = _lib.load(); lib.sb200_gptq4_set_impl(2)
y = torch.zeros(M, N, device=dev)
cuda_kernel.vecgroupquant4matmul(torch.from_numpy(x).to(dev), torch.from_numpy(qw).to(dev), y, torch.from_numpy(scales).to(dev), torch.from_numpy(zeros).to(dev), GS)
torch.cuda.synchronize()
exp = ogptq.dequant_matmul(x, qw, np.zeros((M, N)), scales, zeros, GS)
err = np.abs(y.cpu().numpy() - exp)
print("shape", M, K, N, "max abs err", err.max(), "max |exp|", np.abs(exp).max(), "argmax", np.unravel_index(err.argmax(), err.shape))
print(y.cpu().numpy()[:2, :6]); print(exp[:2, :6])
