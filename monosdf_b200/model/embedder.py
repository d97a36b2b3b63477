"""NeRF positional encoding (reference code/model/embedder.py).  The field kernels fuse this encoding
(csrc/mlp.cu k_encode / k_color_input); this module only reports the widths and offers the stand-alone call."""
import torch


class Embedder:
    def __init__(self, **kwargs):
        self.kwargs = kwargs
        d = kwargs["input_dims"]
        n = kwargs["num_freqs"]
        self.freq_bands = 2.0 ** torch.linspace(0.0, kwargs["max_freq_log2"], n) if kwargs["log_sampling"] else \
            torch.linspace(2.0 ** 0.0, 2.0 ** kwargs["max_freq_log2"], n)
        self.out_dim = (d if kwargs["include_input"] else 0) + d * 2 * n

    def embed(self, inputs):
        parts = [inputs] if self.kwargs["include_input"] else []
        for f in self.freq_bands:
            parts.append(torch.sin(inputs * f))
            parts.append(torch.cos(inputs * f))
        return torch.cat(parts, -1)


def get_embedder(multires, input_dims=3):
    e = Embedder(include_input=True, input_dims=input_dims, max_freq_log2=multires - 1, num_freqs=multires,
                 log_sampling=True, periodic_fns=[torch.sin, torch.cos])
    return e.embed, e.out_dim
