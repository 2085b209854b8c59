"""Batched device classes mirroring spin_torque_gym.devices (reference: devices/base_device.py, stt_mram.py, sot_mram.py,
vcma_mram.py, device_factory.py). Same constructors (a `device_params` dict) and method names; `magnetization` may be one
vector [3] or a batch [N,3], NumPy or torch. Field / torque / resistance / K_eff(V) run as CUDA kernels (K4, csrc/
device_kernels.cu) through the C-ABI; there is no CPU fallback for them."""
from .devices import (BaseSpintronicDevice, DeviceFactory, SOTMRAMDevice, STTMRAMDevice, VCMAMRAMDevice,  # noqa: F401
                      create_device)
