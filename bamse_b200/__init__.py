"""bamse_b200 -- B200-native accelerator for BAMSE's tree marginal-likelihood hot path.

Only what the path needs: csrc/ (CUDA kernels + the C ABI, built into
libbamse_cuda.so), cuda_shim (ctypes binding) and the host-side mirror of the
reference's `clustering` interface.  No CPU fallback exists anywhere in this package.
"""
__version__ = "0.1.0"
