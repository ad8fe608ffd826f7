from .stft_loss import *  # noqa: F401,F403
