"""movfem_b200 -- B200-native element assembly for MoVFEM_3DMT (one hot path, nothing else).

``movfem_b200.host``  ctypes mirror of the reference interface (global_vfem + zero strip)
``movfem_b200.mesh``  synthetic stand-in for geometry.f90's outputs (test / bench inputs)
``movfem_b200.csrc``  CUDA kernels + the C-ABI shared library (include/movfem_b200.h)
"""
__version__ = "0.1.0"
