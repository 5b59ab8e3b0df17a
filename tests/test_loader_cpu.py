"""CPU: the loader mirror (music_b200/wavenet/faster_audio_data.py) against golden outputs of the reference loader."""
import pickle

import numpy as np
import torch

from music_b200.wavenet.faster_audio_data import audio_dataset, one_hot_encode


def test_windowing_and_reshape_onehot_match_reference(golden, tmp_path):
    z = golden("loader")
    items = [z[f"item{i}"] for i in range(int(z["n_items"]))]
    path = tmp_path / "np_audio.pkl"
    with open(path, "wb") as f:
        pickle.dump(items, f)
    ds = audio_dataset(str(path), int(z["rf"]), int(z["window"]))
    assert len(ds) == int(z["n_pieces"])
    for i in range(len(ds)):
        assert np.array_equal(ds.data[i]["audio_piece"].numpy(), z[f"piece{i}"])
        assert np.array_equal(ds.data[i]["audio_target"].numpy(), z[f"target{i}"])
    got = one_hot_encode({"audio_piece": torch.from_numpy(z["onehot_in"]), "audio_target": torch.zeros(1)})
    assert np.array_equal(got["audio_piece"].numpy(), z["onehot_out"])          # the reshape quirk, bit for bit
    real = one_hot_encode({"audio_piece": torch.from_numpy(z["onehot_in"]), "audio_target": torch.zeros(1)}, transpose=True)
    assert np.array_equal(real["audio_piece"].numpy().argmax(0), z["onehot_in"])
    ds_idx = audio_dataset(str(path), int(z["rf"]), int(z["window"]), encoding="index")
    s = ds_idx[0]
    assert s["audio_piece"].dtype == torch.int64 and s["audio_piece"].shape[0] == int(z["rf"]) + int(z["window"]) - 1
    ds_codes = audio_dataset(str(path), int(z["rf"]), int(z["window"]), encoding="codes")
    assert torch.equal(ds_codes[3]["audio_piece"], ds_idx[3]["audio_piece"])
