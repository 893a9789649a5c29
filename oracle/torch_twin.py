"""TEST / BENCH INFRASTRUCTURE ONLY.  The reference architecture assembled from STOCK torch.nn modules (nn.Conv2d,
nn.InstanceNorm2d, nn.TransformerDecoder, nn.Embedding, nn.Conv1d, CrossEntropyLoss, Adam), i.e. what the reference
repository executes -- cuDNN / cuBLAS / SDPA library kernels on a GPU.  It cannot be the imported reference itself
(that tree does not travel to the GPU box and needs lightning/librosa), so it is restated here from the same
description as oracle/restate.py (reference src/transformer/{encoder,decoder,model}.py) and pinned
against restate.py by tests/test_oracle_golden.py::test_torch_twin_matches_restatement.

bench.py times it on the same B200 as the "library kernels to beat" leg (`library_baseline`): same shapes, bf16
autocast, forward + backward + Adam.  It is never on the product path.
"""
from __future__ import annotations

import math
from typing import Optional

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import restate


class _DepthSep(nn.Module):
    def __init__(self, cin: int, cout: int):
        super().__init__()
        self.depth_conv = nn.Conv2d(cin, cin, 3, padding=1, groups=cin)
        self.point_conv = nn.Conv2d(cin, cout, 1)

    def forward(self, x):
        return self.point_conv(self.depth_conv(x))


class _ConvBlock(nn.Module):
    def __init__(self, cin: int, cout: int, stride):
        super().__init__()
        self.conv1 = nn.Conv2d(cin, cout, 3, padding=1)
        self.conv2 = nn.Conv2d(cout, cout, 3, padding=1)
        self.conv3 = nn.Conv2d(cout, cout, 3, padding=1, stride=stride)
        self.norm = nn.InstanceNorm2d(cout, eps=1e-3, momentum=0.99, track_running_stats=False)
        self.drop = nn.Dropout(0.5)

    def forward(self, x):
        x = F.relu(self.conv1(x))
        x = F.relu(self.conv2(x))
        x = self.norm(x)
        x = F.relu(self.conv3(x))
        return self.drop(x)  # ONE dropout per block, as the reference's MixDropout (position fixed here)


class _DSCBlock(nn.Module):
    def __init__(self, cin: int, cout: int):
        super().__init__()
        self.conv1, self.conv2, self.conv3 = _DepthSep(cin, cout), _DepthSep(cout, cout), _DepthSep(cout, cout)
        self.norm = nn.InstanceNorm2d(cout, eps=1e-3, momentum=0.99, track_running_stats=False)
        self.drop = nn.Dropout(0.5)

    def forward(self, x):
        x = F.relu(self.conv1(x))
        x = F.relu(self.conv2(x))
        x = self.norm(x)
        return self.drop(self.conv3(x))


class TwinEncoder(nn.Module):
    def __init__(self, in_channels: int = 1):
        super().__init__()
        chans = [in_channels, 16, 32, 64, 128, 128]
        self.conv_blocks = nn.ModuleList(_ConvBlock(chans[i], chans[i + 1], restate.CONV_BLOCK_STRIDES[i]) for i in range(5))
        self.dscblocks = nn.ModuleList([_DSCBlock(128, 128), _DSCBlock(128, 128), _DSCBlock(128, 128), _DSCBlock(128, 256)])

    def forward(self, x):
        for b in self.conv_blocks:
            x = b(x)
        for b in self.dscblocks:
            xt = b(x)
            x = x + xt if x.shape == xt.shape else xt
        return x


class TwinMultimodal(nn.Module):
    """concat mixer; forward(xi, xli, xa, xla, y_in) -> logits [B,V,T] with the reference's mask algebra"""

    def __init__(self, vocab: int, max_len: int, img_hw, aud_hw, d: int = 256, nhead: int = 4, layers: int = 8, ff: int = 256):
        super().__init__()
        self.image_encoder, self.audio_encoder = TwinEncoder(1), TwinEncoder(1)
        self.register_buffer("image_pe", restate.pe2d_table(d, math.ceil(img_hw[0] / 16), math.ceil(img_hw[1] / 8)))
        self.register_buffer("audio_pe", restate.pe2d_table(d, math.ceil(aud_hw[0] / 16), math.ceil(aud_hw[1] / 8)))
        self.register_buffer("pe1d", restate.pe1d_table(max_len, d))
        self.embedding = nn.Embedding(vocab, d, padding_idx=0)
        layer = nn.TransformerDecoderLayer(d_model=d, nhead=nhead, dim_feedforward=ff, dropout=0.1, batch_first=True)
        self.transformer_decoder = nn.TransformerDecoder(layer, num_layers=layers)
        self.out_layer = nn.Conv1d(d, vocab, kernel_size=1)
        self.pe_drop = nn.Dropout(0.1)

    def _memory(self, enc, pe, x):
        f = enc(x)
        f = self.pe_drop(f + pe[:, :, : f.shape[2], : f.shape[3]].to(f.dtype))
        return f.flatten(2).permute(0, 2, 1).contiguous()

    def forward(self, xi, xli: Optional[torch.Tensor], xa, xla: Optional[torch.Tensor], y_in):
        mi = self._memory(self.image_encoder, self.image_pe, xi)
        ma = self._memory(self.audio_encoder, self.audio_pe, xa)
        mem = torch.cat([mi, ma], dim=1)
        mem_mask = None
        tgt_pad = None
        if xli is not None and xla is not None:
            pi = torch.arange(mi.shape[1], device=mem.device)[None, :] >= xli[:, None]
            pa = torch.arange(ma.shape[1], device=mem.device)[None, :] >= xla[:, None]
            mem_mask = torch.cat([pi, pa], dim=1)  # bool: true masking (concat mixer)
            tgt_pad = (y_in == 0).to(torch.float32)  # FLOAT: additive +1.0 (reference decoder.py:253)
        t = y_in.shape[1]
        tgt = self.pe_drop(self.embedding(y_in) + self.pe1d[:, :t].to(mem.dtype))
        causal = nn.Transformer.generate_square_subsequent_mask(t, device=mem.device)
        h = self.transformer_decoder(tgt, mem, tgt_mask=causal, tgt_key_padding_mask=tgt_pad, memory_key_padding_mask=mem_mask)
        return self.out_layer(h.permute(0, 2, 1).contiguous())

    def load_reference_state_dict(self, sd) -> None:
        """load a reference-named state dict (the keys of the product model / of oracle/restate.py)"""
        own = self.state_dict()
        mapped = {}
        for k, v in sd.items():
            k2 = k.replace("decoder.transformer_decoder.", "transformer_decoder.").replace("decoder.embedding.", "embedding.")
            k2 = k2.replace("decoder.out_layer.", "out_layer.")
            if k2 in own and own[k2].shape == v.shape:
                mapped[k2] = v
        mapped["image_pe"], mapped["audio_pe"], mapped["pe1d"] = own["image_pe"], own["audio_pe"], own["pe1d"]
        missing = [k for k in own if k not in mapped]
        if missing:
            raise KeyError(f"torch twin: unmapped parameters {missing[:5]} ...")
        self.load_state_dict(mapped)
