"""Device selector — identical to the reference's enum (nanograd/device.py:4-11).

The enum carries no device ordinal: ``Device.GPU`` means "the GPU of this process".  Data
parallel runs use one process per GPU (see nanograd_b200/dist.py), so the enum stays as is.
"""
from enum import Enum


class Device(Enum):
    CPU = 1
    GPU = 2
