import torch

_ACT = {"tanh": torch.tanh, "relu": torch.relu, None: None}


class Dense:
    """y = activation(x @ kernel + bias); kernel [in, units] built on first call, bias zeros."""

    def __init__(self, units, activation=None, kernel_initializer=None):
        self.units = int(units)
        self.activation = _ACT[activation] if (activation is None or isinstance(activation, str)) else activation
        self.init = kernel_initializer
        self.kernel = None
        self.bias = None

    def variables(self):
        return [] if self.kernel is None else [self.kernel, self.bias]

    def __call__(self, x):
        if self.kernel is None:
            self.kernel = self.init([x.shape[-1], self.units]).clone().requires_grad_(True)
            self.bias = torch.zeros(self.units, requires_grad=True)
        y = x @ self.kernel + self.bias
        return y if self.activation is None else self.activation(y)
