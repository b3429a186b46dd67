"""PositionalEncoding with the reference's constructor, buffers and state_dict names
(``src/model/code.py``); ``forward`` runs the sm_100a kernel (``pnr_positional_encoding``)."""
import numpy as np
import torch

from .. import _lib


class PositionalEncoding(torch.nn.Module):
    def __init__(self, num_freqs=6, d_in=3, freq_factor=np.pi, include_input=True):
        super().__init__()
        self.num_freqs = num_freqs
        self.d_in = d_in
        self.freq_factor = float(freq_factor)
        self.freqs = freq_factor * 2.0 ** torch.arange(0, num_freqs)
        self.d_out = self.num_freqs * 2 * d_in
        self.include_input = include_input
        if include_input:
            self.d_out += d_in
        # kept for checkpoint compatibility (code._freqs / code._phases are in the reference state_dict)
        self.register_buffer("_freqs", torch.repeat_interleave(self.freqs, 2).view(1, -1, 1))
        phases = torch.zeros(2 * self.num_freqs)
        phases[1::2] = np.pi * 0.5
        self.register_buffer("_phases", phases.view(1, -1, 1))

    def forward(self, x):
        """x (batch, d_in) -> (batch, d_out) = [x, sin(f0 x), cos(f0 x), ...]."""
        _lib.require_cuda(x, "PositionalEncoding input")
        _lib.require_device(x.device)
        x = x.contiguous().float()
        out = torch.empty(x.shape[0], self.d_out, device=x.device, dtype=torch.float32)
        with torch.cuda.device(x.device):
            rc = _lib.load().pnr_positional_encoding(x.data_ptr(), out.data_ptr(), x.shape[0], self.d_in,
                                                     self.num_freqs, self.freq_factor, int(self.include_input),
                                                     _lib.stream_ptr(x.device))
        _lib.check(rc, "pnr_positional_encoding")
        return out

    @classmethod
    def from_conf(cls, conf, d_in=3):
        return cls(conf.get_int("num_freqs", 6), d_in, conf.get_float("freq_factor", np.pi),
                   conf.get_bool("include_input", True))
