import torch

from . import initializers, layers, optimizers  # noqa: F401


class Model:
    """Minimal tf.keras.Model: __call__ -> call; trainable_variables = Dense kernels/biases in layer order, then
    scalar Variables held as attributes (the order Keras tracks them in for the reference's two Net classes)."""

    def __init__(self, *a, **k):
        pass

    def __call__(self, *args, **kwargs):
        return self.call(*args, **kwargs)

    @property
    def trainable_variables(self):
        out = []
        for v in self.__dict__.values():
            if isinstance(v, (list, tuple)):
                for lyr in v:
                    if isinstance(lyr, layers.Dense):
                        out += lyr.variables()
            elif isinstance(v, layers.Dense):
                out += v.variables()
        for v in self.__dict__.values():
            if isinstance(v, torch.Tensor) and v.requires_grad:
                out.append(v)
        return out
