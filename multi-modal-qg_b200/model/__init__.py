"""Drop-in replacement of the reference's `model` package (model/encoder.py, model/decoder.py):
same class names, constructor and forward signatures and state_dict keys, CUDA kernels inside.
Put this directory's parent ahead of the reference on PYTHONPATH and `train.py`'s
`from model.encoder import ...` / `from model.decoder import ...` resolve here."""
