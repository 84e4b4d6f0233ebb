"""Probe: theoretical residency of the library kernels via the CUDA occupancy API."""
import ctypes as C, sys
sys.path.insert(0, '.')
import torch
torch.cuda.init()
