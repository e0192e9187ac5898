import torch


def ifft(x):
    return torch.fft.ifft(x)
