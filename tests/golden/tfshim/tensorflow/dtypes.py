import numpy as np
import torch


def complex(re, im):  # noqa: A001
    re = re if isinstance(re, torch.Tensor) else torch.as_tensor(np.asarray(re, dtype=np.float32))
    im = im if isinstance(im, torch.Tensor) else torch.as_tensor(np.asarray(im, dtype=np.float32))
    re, im = torch.broadcast_tensors(re.to(torch.float32), im.to(torch.float32))
    return torch.complex(re.contiguous(), im.contiguous())
