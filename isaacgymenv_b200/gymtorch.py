"""``gymtorch`` shim (Isaac Gym's ``isaacgym.gymtorch``): zero-copy exchange between sim-owned device
buffers and torch tensors (reference: ``tasks/anymal.py:121-126,229,289-297``)."""
from __future__ import annotations

from . import _lib
from .gymapi import GymTensor


def wrap_tensor(gym_tensor: GymTensor, counts=None, offsets=None):
    """Non-owning torch view of a sim tensor (DLPack, kDLCUDA).  The sim keeps ownership; the view stays valid for
    the sim's lifetime, and in-place writes are seen by the next ``simulate`` (root/DOF state are live sim state)."""
    if gym_tensor.tensor is not None:
        return gym_tensor.tensor
    return _lib.desc_to_torch(gym_tensor.desc)


def unwrap_tensor(tensor) -> GymTensor:
    """Descriptor of a torch tensor for the ``gym.set_*`` calls.  The tensor must be contiguous and live on the sim
    device; float32 for state/targets, int32 for index tensors (``tasks/anymal.py:289``)."""
    if not tensor.is_cuda:
        raise _lib.B2GError("gymtorch.unwrap_tensor: tensor must live on the simulation (CUDA) device")
    if not tensor.is_contiguous():
        tensor = tensor.contiguous()
    return GymTensor(None, tensor)
