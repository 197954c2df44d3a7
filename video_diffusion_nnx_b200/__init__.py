"""B200-native (sm_100a) Unet3D / GaussianDiffusion hot path."""
