import numpy as _np


def create_device_mesh(shape, devices=None):
    return _np.array(devices if devices is not None else ["cpu:0"], dtype=object).reshape(shape)
