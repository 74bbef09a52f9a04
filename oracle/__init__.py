"""CPU oracle for the fake-quant / observer / sparser / GPTQ-int4 hot path.

TEST INFRASTRUCTURE ONLY.  This package is a numpy restatement of the reference's algorithm
(megvii-research/Sparsebit @ f473aef); every function cites the reference file:line it follows.
Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl
reference`` leg may import it -- as the checker / the timed CPU baseline, never as part of the
product path (``sparsebit_b200/`` never imports ``oracle`` and fails loudly without its CUDA
library).

Pinning: the reference's own tests hold NO golden vectors for the QDQ / observer / sparser ops
(SURVEY.md section 4), so this oracle is pinned against outputs of the unmodified reference Python
imported in the build container (``tests/golden/make_golden.py`` -> ``tests/golden/*.npz``), and
for GPTQ additionally against the reference's known-answer test
(large_language_models/llama/quantization/test_cuda_kernel.py: QuantLinear(x) == Linear(dequantised
W)(x) at rtol = atol = 1e-5).
"""
from . import gptq, observers, qdq, sparse  # noqa: F401
