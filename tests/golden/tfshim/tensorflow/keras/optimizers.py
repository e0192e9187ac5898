import math

import torch


class Adam:
    """Keras OptimizerV2 Adam (TF 2.10): lr_t = lr*sqrt(1-b2^t)/(1-b1^t); m += (g-m)(1-b1); v += (g^2-v)(1-b2);
    var -= lr_t*m/(sqrt(v)+eps), eps = 1e-7; one step counter per optimizer object, slots per variable."""

    def __init__(self, learning_rate=0.001, beta_1=0.9, beta_2=0.999, epsilon=1e-7):
        self.lr, self.b1, self.b2, self.eps = float(learning_rate), beta_1, beta_2, epsilon
        self.t = 0
        self.slots = {}

    def apply_gradients(self, grads_and_vars):
        self.t += 1
        lr_t = self.lr * math.sqrt(1 - self.b2 ** self.t) / (1 - self.b1 ** self.t)
        with torch.no_grad():
            for g, v in grads_and_vars:
                if g is None:
                    continue
                m, s = self.slots.setdefault(id(v), (torch.zeros_like(v), torch.zeros_like(v)))
                m += (g - m) * (1 - self.b1)
                s += (g * g - s) * (1 - self.b2)
                v -= lr_t * m / (torch.sqrt(s) + self.eps)
