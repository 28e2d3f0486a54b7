"""Drop-in mirror of the reference's ``modules`` package for the VQA forward path
(same class names, constructor signatures, parameter names and dict-in/dict-out
conventions; forward runs on the C-ABI CUDA kernels)."""
