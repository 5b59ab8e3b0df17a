"""Drop-in for the reference's `wavenet/` script directory (model, audio_func, fast_generate, train,
faster_audio_data), backed by libwavenet_b200.so."""
