"""Drop-in for the reference's `wavenet_autoencoder/` package (model1.py), backed by libwavenet_b200.so."""
