def check_shape(tensor, pattern, **kwargs):  # stubbed exactly like gaussian_diffusion_test.py:52-56 patches it
    return tensor
