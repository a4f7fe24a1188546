"""diffspectra_b200 — B200-native (sm_100a) implementation of DiffSpectra's reverse-diffusion sampling hot path."""
from ._lib import DiffSpectraError, LIB_PATH  # noqa: F401
