"""Drop-in for the reference's `wavenet/faster_audio_data.py`: in-memory dataset of mu-law windows.

Reference: audio_dataset :7-48 (windowing `_make_data_pieces` :24-40), audio_data_loader :51-59, one_hot_encode :62-83.
Host-side data plumbing only (no arithmetic hot path).  Same on-disk format (a pickled list of int arrays,
wavenet/data/wav_to_numpy.py:34-35), same windowing - including its tail behaviour: when fewer than rf + window samples
remain, the loop re-appends the PREVIOUS (piece, target) pair (`piece` / `target` are stale at :34-39).

`encoding` selects what `__getitem__` returns as "audio_piece":
  "reference" (default) the reference's (Q, T) float tensor, which is NOT a one-hot: it builds a (T, Q) one-hot and then
              `reshape(Q, T)` instead of transposing (:77-81).  Feeds the dense-input path.
  "onehot"    the true (Q, T) one-hot (what the reshape was meant to be).
  "index"     the int64 codes (T,): feeds `Trainer.step` / `forward_indices` without materialising Q x T floats
              (one-hot happens on the device, inside the causal-layer gather).
  "codes"     the int64 codes (T,) too, but meant for `one_hot_encode_device`: the training loop builds the reference's
              (Q, T) tensor (reshape quirk included) ON THE DEVICE from the codes, so that 8 bytes per sample cross PCIe
              instead of 1 KB (312 MB per batch of 16 x 19070 samples) and no numpy one-hot runs on the host.
"""
from __future__ import annotations

import pickle

import numpy as np
import torch
from torch.utils.data import DataLoader, Dataset


class audio_dataset(Dataset):

    def __init__(self, audio_path, receptive_field, window_length, cuda_available=False, quantization_channels=256,
                 encoding="reference"):
        self.audio_path = audio_path
        self.receptive_field = receptive_field
        self.window_length = window_length
        self.cuda_available = cuda_available
        self.quantization_channels = quantization_channels
        if encoding not in ("reference", "onehot", "index", "codes"):
            raise ValueError(encoding)
        self.encoding = encoding
        with open(self.audio_path, 'rb') as f:
            data = pickle.load(f)
        self.data = self._make_data_pieces(data)

    def _make_data_pieces(self, data):
        rf, win = self.receptive_field, self.window_length
        data_pieces = []
        piece = target = None
        for item in data:
            item = torch.from_numpy(np.asarray(item))
            while len(item) > rf:
                if len(item) >= rf + win:
                    piece = item[:rf + win - 1]
                    target = item[rf:rf + win]
                    item = item[win:]
                else:                       # short remainder: the previous pair is appended again (reference behaviour)
                    item = item[rf:]
                target = target.long()
                data_pieces.append({'audio_piece': piece, 'audio_target': target})
        return data_pieces

    def __len__(self):
        return len(self.data)

    def __getitem__(self, idx):
        if self.encoding in ("index", "codes"):
            s = self.data[idx]
            return {"audio_piece": s['audio_piece'].long(), "audio_target": s['audio_target']}
        return one_hot_encode(self.data[idx], self.cuda_available, self.quantization_channels,
                              transpose=(self.encoding == "onehot"))


def audio_data_loader(batch_size, shuffle, num_workers, pin_memory, **kwargs):
    audioDataset = audio_dataset(**kwargs)
    print("{} pieces in total".format(len(audioDataset)))
    return DataLoader(audioDataset, batch_size=batch_size, shuffle=shuffle, num_workers=num_workers, pin_memory=pin_memory)


def one_hot_encode(sample_piece, cuda_available=False, quantization_channels=256, transpose=False):
    """{'audio_piece': int (T,), 'audio_target': ...} -> the piece as a (Q, T) float tensor.  transpose=False reproduces
    the reference bit for bit (one-hot (T, Q) RESHAPED to (Q, T)); transpose=True gives the real one-hot."""
    piece, target = sample_piece['audio_piece'], sample_piece['audio_target']
    seq_len = piece.size()[0]
    piece_one_hot = np.zeros((seq_len, quantization_channels))
    piece_one_hot[np.arange(seq_len), piece.numpy()] = 1.0
    piece_one_hot = piece_one_hot.T if transpose else piece_one_hot.reshape(quantization_channels, seq_len)
    return {"audio_piece": torch.FloatTensor(np.ascontiguousarray(piece_one_hot)), "audio_target": target}


def one_hot_encode_device(codes, quantization_channels=256, transpose=False):
    """`one_hot_encode` (:62-83) on the GPU: (B, T) or (T,) integer codes on a CUDA device -> (B, Q, T) float32 (or (Q, T)).
    transpose=False is the reference's tensor bit for bit - a (T, Q) one-hot RESHAPED to (Q, T), i.e. not a one-hot;
    transpose=True the true one-hot.  Runs in libwavenet_b200.so (csrc/codec.cu); no CPU fallback."""
    from .. import _lib as L
    if not codes.is_cuda:
        raise L.WavenetB200Error("one_hot_encode_device: codes must be on the GPU (use one_hot_encode for the host path)")
    single = codes.dim() == 1
    c = codes.reshape(1, -1) if single else codes
    c = c.to(torch.int64).contiguous()
    B, T = c.shape
    lib = L.init(c.device.index if c.device.index is not None else torch.cuda.current_device())
    out = torch.empty(B, quantization_channels, T, dtype=torch.float32, device=c.device)
    L.check(lib.wn_onehot_encode(L.ptr(c), B, T, quantization_channels, int(bool(transpose)), L.ptr(out), L.stream_ptr()))
    return out[0] if single else out
