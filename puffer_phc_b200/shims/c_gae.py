"""Top-level ``c_gae`` module: with this directory on ``sys.path`` (``puffer_phc_b200.install_c_gae_shim()``) the reference's

    pyximport.install(...); from c_gae import compute_gae            (reference puffer_phc/clean_pufferl/core.py:33-36)

resolves here instead of compiling ``c_gae.pyx``: same name, same signature, numpy float32 arrays in, a NEW numpy float32 array out
(reference puffer_phc/c_gae.pyx:11-32), computed by the CUDA kernel ``phc_gae`` (H2D, kernel, D2H); CUDA tensors are accepted too and
stay on the device."""
from puffer_phc_b200.c_gae import compute_gae, compute_gae_cuda  # noqa: F401
