"""lidal_b200 -- B200-native hot path of hzykent/LiDAL behind the reference's own interfaces.

* ``lidal_b200.compat``   torchsparse-1.4.0 drop-in (SparseTensor, spnn.Conv3d, F.sphash ... ) on sm_100a kernels
* ``lidal_b200.network``  MinkUNet / SPVCNN mirrors (same state_dict) for tests and benchmarks
* ``lidal_b200.engine``   fused inference plan (BN/ReLU/residual folded into the conv epilogue, bf16 activations)
* ``lidal_b200.score``    prob_inference tail + inter-frame scoring + region selection (score.sv_level.LiDAL)
* ``lidal_b200.synth``    seeded synthetic scans / sequences

The compute path is the C-ABI library ``liblidal_b200.so`` (include/lidal_b200.h); there is no CPU fallback.
"""
__version__ = "0.1.0"
